// Attention cores on the packed qkv tensor bf16 [B][N][3*heads*32] produced by the to_qkv 1x1
// GEMM (q | k | v, each heads*32 channels, head-major -- the chunk(3)/rearrange of model.py:312-313
// and 349-350).  dim_head is fixed at 32 (config default; one lane per head channel).
//
//  * LinearAttention core (model.py:315-323):
//        q = softmax_d(q) * 32^-1/2 ; k = softmax_n(k) ; ctx[d][e] = sum_n k[d][n] v[e][n] ;
//        out[e][n] = sum_d ctx[d][e] q[d][n]
//    Phase 1 streams k,v once: every block owns a pixel range and keeps, per head, a running
//    column max m[d], normaliser Z[d] and the unnormalised 32x32 context (online softmax over n),
//    then writes the partial triple.  Phase 2 merges the partials (rescale by exp(m_i - m)),
//    divides by Z and folds the 32^-1/2 of q in.  Phase 3 streams q once and applies ctx.
//    HBM-bound: reads 3*128 ch, writes 128 ch per pixel (bf16) = 1 KiB/pixel algorithmic bytes.
//  * Full attention core (Attend, model.py:352): flash-style, one query per thread, K/V tiles in
//    shared memory, online softmax in fp32.  (N = 1024 tokens, 0.2 % of the U-Net's FLOPs.)
#include "common.cuh"

namespace srgd {

constexpr int kDH = 32;                 // dim_head
constexpr float kQScale = 0.17677669529663687f;   // 32^-1/2  (model.py:295, 318; Attend default scale)

// ---------------------------------------------------------------------------------------------
// linear attention, phase 1: partial (m, Z, ctx) per (sample, split, head)
// block = heads warps (warp h <-> head h); lane d <-> k channel d
// ---------------------------------------------------------------------------------------------
constexpr int kLaTile = 64;             // pixels staged per iteration

template <int HEADS>
__global__ void __launch_bounds__(HEADS * 32) la_context_partial_kernel(const bf16* __restrict__ qkv,
                                                                       float* __restrict__ part, int N,
                                                                       int splits) {
  constexpr int HID = HEADS * kDH;                       // 128
  constexpr int ROW = 3 * HID;                           // 384 channels per pixel
  __shared__ __align__(16) bf16 sk[kLaTile][HID];        // k tile, all heads
  __shared__ __align__(16) bf16 sv[kLaTile][HID];        // v tile
  const int b = blockIdx.y, sp = blockIdx.x;
  const int h = threadIdx.x >> 5, d = threadIdx.x & 31;
  const int per = (N + splits - 1) / splits;
  const int n_begin = sp * per;
  const int n_end = min(N, n_begin + per);
  const bf16* base = qkv + (int64_t)b * N * ROW;

  float m = -INFINITY, z = 0.f;
  float ctx[kDH];
#pragma unroll
  for (int e = 0; e < kDH; ++e) ctx[e] = 0.f;

  for (int n0 = n_begin; n0 < n_end; n0 += kLaTile) {
    const int cnt = min(kLaTile, n_end - n0);
    // cooperative load: per pixel 2*HID bf16 (k then v) = 2*HID/8 16-byte vectors
    constexpr int VPP = 2 * HID / 8;
    for (int i = threadIdx.x; i < cnt * VPP; i += blockDim.x) {
      const int pix = i / VPP, v = i % VPP;
      const uint4 val = ld_stream(base + (int64_t)(n0 + pix) * ROW + HID + v * 8);
      if (v < HID / 8) *reinterpret_cast<uint4*>(&sk[pix][v * 8]) = val;
      else *reinterpret_cast<uint4*>(&sv[pix][(v - HID / 8) * 8]) = val;
    }
    __syncthreads();
    // online softmax over n for column (h, d)
    float tmax = -INFINITY;
    for (int n = 0; n < cnt; ++n) tmax = fmaxf(tmax, __bfloat162float(sk[n][h * kDH + d]));
    const float m_new = fmaxf(m, tmax);
    const float resc = __expf(m - m_new);                // 0 on the first tile (m = -inf)
    z *= resc;
#pragma unroll
    for (int e = 0; e < kDH; ++e) ctx[e] *= resc;
    m = m_new;
    for (int n = 0; n < cnt; ++n) {
      const float w = __expf(__bfloat162float(sk[n][h * kDH + d]) - m);
      z += w;
      const uint4* vrow = reinterpret_cast<const uint4*>(&sv[n][h * kDH]);   // broadcast reads
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float f[8];
        unpack8(vrow[t], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) ctx[t * 8 + j] = fmaf(w, f[j], ctx[t * 8 + j]);
      }
    }
    __syncthreads();
  }
  // partial record: [b][split][head][d][34] = m, Z, ctx[32]
  float* dst = part + ((((int64_t)b * splits + sp) * HEADS + h) * kDH + d) * 34;
  dst[0] = m;
  dst[1] = z;
#pragma unroll
  for (int e = 0; e < kDH; ++e) dst[2 + e] = ctx[e];
}

// phase 2: merge partials -> ctx_final[b][h][d][e] = 32^-1/2 * sum_n softmax_n(k)[d][n] v[e][n]
__global__ void __launch_bounds__(32) la_context_merge_kernel(const float* __restrict__ part,
                                                              float* __restrict__ ctx_out, int heads, int splits) {
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int d = threadIdx.x;
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s)
    m = fmaxf(m, part[((((int64_t)b * splits + s) * heads + h) * kDH + d) * 34]);
  float z = 0.f, acc[kDH];
#pragma unroll
  for (int e = 0; e < kDH; ++e) acc[e] = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* src = part + ((((int64_t)b * splits + s) * heads + h) * kDH + d) * 34;
    const float w = __expf(src[0] - m);
    z += w * src[1];
#pragma unroll
    for (int e = 0; e < kDH; ++e) acc[e] = fmaf(w, src[2 + e], acc[e]);
  }
  const float inv = kQScale / z;
  float* dst = ctx_out + (((int64_t)b * heads + h) * kDH + d) * kDH;
#pragma unroll
  for (int e = 0; e < kDH; ++e) dst[e] = acc[e] * inv;
}

// phase 3: out[n][h*32+e] = sum_d ctx[d][e] * softmax_d(q[n][h*32+:])[d]
// thread <-> (pixel, head); a warp covers 32 consecutive pixels of one head (ctx reads broadcast)
template <int HEADS>
__global__ void __launch_bounds__(HEADS * 32) la_apply_kernel(const bf16* __restrict__ qkv,
                                                             const float* __restrict__ ctx, bf16* __restrict__ out,
                                                             int N) {
  constexpr int HID = HEADS * kDH, ROW = 3 * HID;
  __shared__ __align__(16) float sc[HEADS][kDH][kDH];
  const int b = blockIdx.y;
  const int h = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < HEADS * kDH * kDH; i += blockDim.x)
    (&sc[0][0][0])[i] = ctx[(int64_t)b * HEADS * kDH * kDH + i];
  __syncthreads();
  for (int n = blockIdx.x * 32 + (threadIdx.x & 31); n < N; n += gridDim.x * 32) {
    const bf16* qp = qkv + ((int64_t)b * N + n) * ROW + h * kDH;
    float q[kDH];
#pragma unroll
    for (int t = 0; t < 4; ++t) unpack8(ld_stream(qp + t * 8), q + t * 8);
    float mx = q[0];
#pragma unroll
    for (int d = 1; d < kDH; ++d) mx = fmaxf(mx, q[d]);
    float sum = 0.f;
#pragma unroll
    for (int d = 0; d < kDH; ++d) { q[d] = __expf(q[d] - mx); sum += q[d]; }
    const float inv = 1.0f / sum;
    float o[kDH];
#pragma unroll
    for (int e = 0; e < kDH; ++e) o[e] = 0.f;
#pragma unroll 4
    for (int d = 0; d < kDH; ++d) {
      const float w = q[d] * inv;
      const float4* row = reinterpret_cast<const float4*>(&sc[h][d][0]);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float4 c4 = row[t];
        o[t * 4 + 0] = fmaf(w, c4.x, o[t * 4 + 0]);
        o[t * 4 + 1] = fmaf(w, c4.y, o[t * 4 + 1]);
        o[t * 4 + 2] = fmaf(w, c4.z, o[t * 4 + 2]);
        o[t * 4 + 3] = fmaf(w, c4.w, o[t * 4 + 3]);
      }
    }
    bf16* op = out + ((int64_t)b * N + n) * HID + h * kDH;
#pragma unroll
    for (int t = 0; t < 4; ++t) st_stream(op + t * 8, pack8(o + t * 8));
  }
}

// ---------------------------------------------------------------------------------------------
// full attention: block = 64 queries of one (sample, head); K/V staged 64 keys at a time
// ---------------------------------------------------------------------------------------------
constexpr int kFaQ = 64, kFaK = 64;

__global__ void __launch_bounds__(kFaQ) full_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out,
                                                             int N, int heads) {
  const int hid = heads * kDH, row = 3 * hid;
  __shared__ __align__(16) bf16 sk[kFaK][kDH];
  __shared__ __align__(16) bf16 sv[kFaK][kDH];
  const int bh = blockIdx.y;
  const int b = bh / heads, h = bh % heads;
  const int qi = blockIdx.x * kFaQ + threadIdx.x;
  const bool qvalid = qi < N;
  const bf16* base = qkv + (int64_t)b * N * row;
  float q[kDH], o[kDH];
  if (qvalid) {
#pragma unroll
    for (int t = 0; t < 4; ++t) unpack8(ld_stream(base + (int64_t)qi * row + h * kDH + t * 8), q + t * 8);
  } else {
#pragma unroll
    for (int d = 0; d < kDH; ++d) q[d] = 0.f;
  }
#pragma unroll
  for (int d = 0; d < kDH; ++d) { q[d] *= kQScale; o[d] = 0.f; }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < N; k0 += kFaK) {
    const int cnt = min(kFaK, N - k0);
    // 64 keys x (32 k + 32 v) bf16 = 64 x 8 vectors; 64 threads
    for (int i = threadIdx.x; i < cnt * 8; i += blockDim.x) {
      const int kk = i >> 3, v = i & 7;
      const bf16* src = base + (int64_t)(k0 + kk) * row + (v < 4 ? hid : 2 * hid) + h * kDH + (v & 3) * 8;
      const uint4 val = ld_stream(src);
      if (v < 4) *reinterpret_cast<uint4*>(&sk[kk][(v & 3) * 8]) = val;
      else *reinterpret_cast<uint4*>(&sv[kk][(v & 3) * 8]) = val;
    }
    __syncthreads();
    for (int j0 = 0; j0 < cnt; j0 += 8) {
      float s[8];
      float gmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float acc = -INFINITY;
        if (j0 + jj < cnt) {
          acc = 0.f;
          const uint4* kr = reinterpret_cast<const uint4*>(&sk[j0 + jj][0]);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float f[8];
            unpack8(kr[t], f);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = fmaf(q[t * 8 + u], f[u], acc);
          }
        }
        s[jj] = acc;
        gmax = fmaxf(gmax, acc);
      }
      const float m_new = fmaxf(m, gmax);
      const float resc = __expf(m - m_new);
      l *= resc;
#pragma unroll
      for (int d = 0; d < kDH; ++d) o[d] *= resc;
      m = m_new;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        if (j0 + jj < cnt) {
          const float pw = __expf(s[jj] - m);
          l += pw;
          const uint4* vr = reinterpret_cast<const uint4*>(&sv[j0 + jj][0]);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float f[8];
            unpack8(vr[t], f);
#pragma unroll
            for (int u = 0; u < 8; ++u) o[t * 8 + u] = fmaf(pw, f[u], o[t * 8 + u]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (qvalid) {
    const float inv = 1.0f / l;
#pragma unroll
    for (int d = 0; d < kDH; ++d) o[d] *= inv;
    bf16* op = out + ((int64_t)b * N + qi) * hid + h * kDH;
#pragma unroll
    for (int t = 0; t < 4; ++t) st_stream(op + t * 8, pack8(o + t * 8));
  }
}

static int la_splits_mode(int B, int N, bool invariant) {
  if (invariant) B = kInvariantRefBatch;                 // partition independent of the launch's batch size
  int s = (4 * sm_count() + B - 1) / B;                  // ~4 blocks per SM overall
  int max_s = (N + 2 * kLaTile - 1) / (2 * kLaTile);     // at least two tiles per block
  if (s > max_s) s = max_s;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return s;
}
static int la_splits(int B, int N) { return la_splits_mode(B, N, g_batch_invariant != 0); }
static int la_splits_max(int B, int N) {
  const int a = la_splits_mode(B, N, false), b = la_splits_mode(B, N, true);
  return a > b ? a : b;
}

}  // namespace srgd

using namespace srgd;

extern "C" size_t srgd_linear_attention_workspace(int32_t B, int32_t N, int32_t heads) {
  if (B <= 0 || N <= 0 || heads <= 0) return 0;
  const size_t part = (size_t)B * la_splits_max(B, N) * heads * kDH * 34 * sizeof(float);
  const size_t ctx = (size_t)B * heads * kDH * kDH * sizeof(float);
  return part + ctx + 256;
}

extern "C" int srgd_linear_attention(const void* qkv, void* out, int32_t B, int32_t N, int32_t heads,
                                     void* workspace, size_t workspace_bytes, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(qkv && out && workspace && B > 0 && N > 0, "linear_attention: bad arguments");
  SRGD_REQUIRE(heads == 4, "linear_attention: only heads=4, dim_head=32 is built (got heads=%d)", heads);
  SRGD_REQUIRE(B <= 65535, "linear_attention: B too large");
  if (workspace_bytes < srgd_linear_attention_workspace(B, N, heads)) {
    set_error("linear_attention: workspace too small");
    return SRGD_E_WORKSPACE;
  }
  const int splits = la_splits(B, N);
  float* part = reinterpret_cast<float*>(workspace);
  float* ctx = part + (size_t)B * splits * heads * kDH * 34;
  ctx = reinterpret_cast<float*>(((uintptr_t)ctx + 127) & ~(uintptr_t)127);
  cudaStream_t st = as_stream(stream);
  const bf16* in = reinterpret_cast<const bf16*>(qkv);
  ProfScope prof(SRGD_PK_LINEAR_ATTN, 4.0 * (double)B * N * heads * kDH * kDH,
                 2.0 * (double)B * N * heads * kDH * 4, st);
  la_context_partial_kernel<4><<<dim3(splits, B), 128, 0, st>>>(in, part, N, splits);
  SRGD_LAUNCH_OK("la_context_partial_kernel");
  la_context_merge_kernel<<<B * heads, 32, 0, st>>>(part, ctx, heads, splits);
  SRGD_LAUNCH_OK("la_context_merge_kernel");
  int gx = (N + 31) / 32;
  const int cap = (8 * sm_count() + B - 1) / B;
  if (gx > cap) gx = cap;
  la_apply_kernel<4><<<dim3(gx, B), 128, 0, st>>>(in, ctx, reinterpret_cast<bf16*>(out), N);
  SRGD_LAUNCH_OK("la_apply_kernel");
  count_launch(3);
  return SRGD_OK;
}

extern "C" int srgd_attention(const void* qkv, void* out, int32_t B, int32_t N, int32_t heads,
                              srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(qkv && out && B > 0 && N > 0 && heads > 0 && heads <= 16, "attention: bad arguments");
  SRGD_REQUIRE((int64_t)B * heads <= 65535, "attention: B*heads too large");
  ProfScope prof(SRGD_PK_FULL_ATTN, 4.0 * (double)B * heads * (double)N * N * kDH,
                 2.0 * (double)B * N * heads * kDH * 4, as_stream(stream));
  full_attention_kernel<<<dim3((N + kFaQ - 1) / kFaQ, B * heads), kFaQ, 0, as_stream(stream)>>>(
      reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), N, heads);
  SRGD_LAUNCH_OK("full_attention_kernel");
  count_launch();
  return SRGD_OK;
}
