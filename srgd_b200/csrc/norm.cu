// GroupNorm + scale/shift + SiLU (+ residual) and RMSNorm kernels on bf16 NHWC activations.
// Reference: Block.forward model.py:250-259 (nn.GroupNorm(8, C), eps 1e-5, affine), ResnetBlock's
// `h + res_conv(x)` (model.py:285), RMSNorm model.py:201-207.
//
// All of these are HBM-bound streaming passes: 16-byte vector accesses (8 bf16 channels), fp32
// math, per-(sample, channel) affine folded into one FMA:  y = silu(x * A[c] + B[c]) with
//   A = rstd*gamma*(scale+1),  B = (beta - mean*rstd*gamma)*(scale+1) + shift.
// Algorithmic bytes: 2 B read + 2 B written per element (+2 B when a residual is added).
#include <stdlib.h>

#include "common.cuh"
#include "tiles.h"

namespace srgd {

constexpr float kGnEps = 1e-5f;        // nn.GroupNorm default (model.py:247)
constexpr int kGroups = 8;

// ---- statistics ------------------------------------------------------------------------------
// Fold the conv-epilogue partial sums (conv_igemm.cu) of sample n in fp64, by a 256-thread block.  A record is the
// 64 bytes [8 groups][sum, sumsq] of one (M tile, sample slot); thread t reads float4 number (t & 3) of every 64th
// record, so a warp touches 8 whole records per load instruction.  A 256x256 sample has 512 records (32 KB, L2
// resident: the conv wrote them last).  The summation order is fixed, so every block that folds the same sample
// gets bit-identical statistics -- gn_apply_kernel relies on that to fold them redundantly in its prologue instead
// of paying a separate launch.  out16 (shared): [8 groups][mean, rstd]; ends with a block barrier.
__device__ __forceinline__ void gn_sample_stats(const float* __restrict__ partials, int n, const TileGeom& g,
                                                double cnt, float* out16, double (*red)[4][4]) {
  const int tb = n >> g.tn_log2, slot = n & ((1 << g.tn_log2) - 1);
  const int tiles_per_b = g.tiles_x * g.tiles_y;
  const int part = threadIdx.x & 3;                      // groups 2*part, 2*part+1
  const float4* base = reinterpret_cast<const float4*>(partials) +
                       ((((int64_t)tb * tiles_per_b) << g.tn_log2) + slot) * 4 + part;
  const int64_t rstride = (int64_t)4 << g.tn_log2;       // float4s between the records of consecutive tiles
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  constexpr int kU = 4;                                  // independent 16-byte loads in flight per thread
  for (int t0 = threadIdx.x >> 2; t0 < tiles_per_b; t0 += 64 * kU) {
    float4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int t = t0 + u * 64;
      v[u] = t < tiles_per_b ? __ldg(base + t * rstride) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) { a0 += v[u].x; a1 += v[u].y; a2 += v[u].z; a3 += v[u].w; }
  }
  // reduce over the threads with the same `part`: lanes xor 4, 8, 16, then the 8 warps through shared memory
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < 4) { red[warp][lane][0] = a0; red[warp][lane][1] = a1; red[warp][lane][2] = a2; red[warp][lane][3] = a3; }
  __syncthreads();
  if (threadIdx.x < kGroups) {
    const int grp = threadIdx.x, pt = grp >> 1, o = (grp & 1) * 2;
    double sum = 0.0, q = 0.0;
    for (int w = 0; w < 8; ++w) { sum += red[w][pt][o]; q += red[w][pt][o + 1]; }
    const double mean = sum / cnt;
    double var = q / cnt - mean * mean;                  // biased variance, as nn.GroupNorm
    if (var < 0.0) var = 0.0;
    out16[grp * 2] = (float)mean;
    out16[grp * 2 + 1] = (float)(1.0 / sqrt(var + (double)kGnEps));
  }
  __syncthreads();
}

// Stand-alone form (tests, callers that want the statistics): one block per sample -> stats[B][8][2].
__global__ void __launch_bounds__(256) gn_finalize_kernel(const float* __restrict__ partials,
                                                         float* __restrict__ stats, int H, int W, int C,
                                                         TileGeom g) {
  __shared__ double red[8][4][4];
  __shared__ float st16[16];
  pdl_wait();
  gn_sample_stats(partials, blockIdx.x, g, (double)H * W * (C / kGroups), st16, red);
  if (threadIdx.x < 16) stats[blockIdx.x * 16 + threadIdx.x] = st16[threadIdx.x];
}

// Stand-alone statistics over a bf16 NHWC tensor: one block per (sample, group).
__global__ void __launch_bounds__(512) gn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ stats,
                                                       int HW, int C) {
  const int n = blockIdx.x / kGroups, grp = blockIdx.x % kGroups;
  const int G = C / kGroups;                             // multiple of 8
  const int vec_per_pix = G / 8;
  const int64_t total = (int64_t)HW * vec_per_pix;
  const bf16* base = x + (int64_t)n * HW * C + grp * G;
  float s = 0.f, q = 0.f;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int64_t pix = i / vec_per_pix;
    const int v = (int)(i % vec_per_pix);
    float f[8];
    unpack8(ld_stream(base + pix * C + v * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += f[j]; q += f[j] * f[j]; }
  }
  __shared__ double sh[2][16];
  double ds = warp_sum(s), dq = warp_sum(q);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ds; sh[1][threadIdx.x >> 5] = dq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ds = 0; dq = 0;
    for (int i = 0; i < 16; ++i) { ds += sh[0][i]; dq += sh[1][i]; }
    const double cnt = (double)HW * G;
    const double mean = ds / cnt;
    double var = dq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[(n * kGroups + grp) * 2] = (float)mean;
    stats[(n * kGroups + grp) * 2 + 1] = (float)(1.0 / sqrt(var + (double)kGnEps));
  }
}

// ---- apply -----------------------------------------------------------------------------------
// INV_LANES > 0: additionally emit inv_out[pixel] = 1 / max(||y[pixel,:]||_2, 1e-12) of the bf16-rounded output row
// (the RMSNorm of the attention block that consumes y, model.py:207); the C/8 = INV_LANES threads of a pixel are
// consecutive lanes of one warp.
// FINAL: y is not stored; instead the U-Net's final 1x1 conv C -> 3 (model.py:675, 725) is evaluated on the fp32
// values and written as fp32 NCHW eps (fin_w [3][C], fin_b [3]); C = 128, the 16 threads of a pixel reduce by shuffle.
// REG: C/8 is a power of two <= 256, so a thread always meets the same 8 channels (every stride is a multiple of
// 256 vectors) and keeps its affine coefficients -- and the final-conv weights -- in registers: no shared-memory
// traffic and no block barrier.
// (FINAL keeps 24 final-conv weights in registers next to the 16 affine coefficients and spills 72 B per thread under
// the 64-register cap of 4 CTAs per SM; 3 CTAs per SM without spills was measured SLOWER, 0.30 vs 0.24 ms at 16
// tiles -- the pass is bound by loads in flight per SM, not by the spill traffic; profiles/r02_experiments.txt.)
template <bool HAS_RES, int INV_LANES, bool FINAL, bool REG, int U>
__global__ void __launch_bounds__(256, 4) gn_apply_kernel(const bf16* x, int Bx, const float* __restrict__ stats,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta,
                                                       const float* __restrict__ scale_shift, int64_t ss_stride,
                                                       const bf16* residual, bf16* y, float* __restrict__ inv_out,
                                                       const float* __restrict__ fin_w, const float* __restrict__ fin_b,
                                                       float* __restrict__ eps, int HW, int C, int reverse,
                                                       const float* __restrict__ partials, TileGeom geom, int Bres) {
  extern __shared__ float sm[];                          // !REG: A[C] | B[C] | (FINAL: fin_w[3][C])
  __shared__ double red[8][4][4];
  __shared__ float st16[16];
  pdl_wait();                                            // partials / stats come from the preceding launch
  float* sA = sm;
  float* sB = sm + C;
  float* sW = sm + 2 * C;
  // reverse: walk the tensor from its end.  The producer conv wrote x front to back, so its tail is what the
  // 126 MB L2 still holds; reading it first turns ~1/3 of a 268 MB pass into L2 hits, and leaves the head of y
  // in L2 for the consumer conv, which starts at the front again.
  const int b = reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int bs = b % Bx;
  const int G = C / kGroups;
  const int vec_per_pix = C / 8;
  const int64_t total = (int64_t)HW * vec_per_pix;
  const bf16* xs = x + (int64_t)bs * HW * C;
  const bf16* rs = HAS_RES ? residual + (int64_t)(b % Bres) * HW * C : nullptr;   // broadcast residual: Bres rows
  bf16* ys = y + (int64_t)b * HW * C;
  // U independent 16-byte vectors (2U with a residual) in flight per thread per iteration; the first batch is
  // issued BEFORE the affine coefficients are loaded, so a CTA pays one memory latency, not two, before its
  // first store.  Four loads in flight per thread is the measured optimum: U = 2 with a residual (U = 4 there costs
  // occupancy, 90 registers, and was 11 % slower), U = 4 without one (the plain pass at U = 2 moved 4.6 TB/s where
  // the residual pass moved 5.7 TB/s).
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 xv[U], rv[U];
  auto issue = [&](int64_t i0) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t iu = i0 + u * stride;
      if (iu < total) {
        const int64_t ju = reverse ? total - 1 - iu : iu;
        xv[u] = ld_stream(xs + ju * 8);
        if (HAS_RES) rv[u] = ld_stream(rs + ju * 8);
      }
    }
  };
  if (i < total) issue(i);

  // statistics: either given (stats[B][8][2]) or folded here from the producer conv's partial records
  if (partials != nullptr) {
    gn_sample_stats(partials, bs, geom, (double)HW * G, st16, red);
  } else {
    if (threadIdx.x < 16) st16[threadIdx.x] = stats[bs * 16 + threadIdx.x];
    __syncthreads();
  }
  float ra[8], rb[8], rw[24];
  auto coef = [&](int c, float& a, float& bb) {
    const float mean = st16[(c / G) * 2];
    const float rstd = st16[(c / G) * 2 + 1];
    a = rstd * gamma[c];
    bb = beta[c] - mean * a;
    if (scale_shift != nullptr) {
      const float sc = scale_shift[(int64_t)b * ss_stride + c] + 1.0f;     // model.py:256
      const float sh = scale_shift[(int64_t)b * ss_stride + C + c];
      a *= sc;
      bb = bb * sc + sh;
    }
  };
  if (REG) {
    const int cv = (int)(threadIdx.x % vec_per_pix);
    const int c0 = (reverse ? vec_per_pix - 1 - cv : cv) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) coef(c0 + j, ra[j], rb[j]);
    if (FINAL) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        rw[j] = fin_w[c0 + j];
        rw[8 + j] = fin_w[C + c0 + j];
        rw[16 + j] = fin_w[2 * C + c0 + j];
      }
    }
  } else {
    if (FINAL)
      for (int c = threadIdx.x; c < 3 * C; c += blockDim.x) sW[c] = fin_w[c];
    for (int c = threadIdx.x; c < C; c += blockDim.x) coef(c, sA[c], sB[c]);
    __syncthreads();
  }
  while (i < total) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i_u = i + u * stride;
      if (i_u >= total) break;
      const int64_t iu = reverse ? total - 1 - i_u : i_u;
      const int c0 = (int)(iu % vec_per_pix) * 8;
      float f[8];
      unpack8(xv[u], f);
      if (REG) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], ra[j], rb[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sA[c0 + j], sB[c0 + j]);
      }
      silu2(f[0], f[1]); silu2(f[2], f[3]); silu2(f[4], f[5]); silu2(f[6], f[7]);
      if (HAS_RES) {
        float r[8];
        unpack8(rv[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += r[j];
      }
      if (FINAL) {
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          d0 = fmaf(f[j], REG ? rw[j] : sW[c0 + j], d0);
          d1 = fmaf(f[j], REG ? rw[8 + j] : sW[C + c0 + j], d1);
          d2 = fmaf(f[j], REG ? rw[16 + j] : sW[2 * C + c0 + j], d2);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          d0 += __shfl_xor_sync(0xffffffffu, d0, o);
          d1 += __shfl_xor_sync(0xffffffffu, d1, o);
          d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        if ((threadIdx.x & 15) == 0) {
          const int64_t pix = iu / vec_per_pix;
          float* e0 = eps + (int64_t)b * 3 * HW + pix;
          e0[0] = d0 + fin_b[0];
          e0[HW] = d1 + fin_b[1];
          e0[2 * (int64_t)HW] = d2 + fin_b[2];
        }
        continue;
      }
      const uint4 packed = pack8(f);
      st_stream(ys + iu * 8, packed);
      if (INV_LANES > 0) {
        float fr[8];
        unpack8(packed, fr);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) ss = fmaf(fr[j], fr[j], ss);
#pragma unroll
        for (int o = INV_LANES / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if ((threadIdx.x & (INV_LANES - 1)) == 0)
          inv_out[(int64_t)b * HW + iu / vec_per_pix] = fminf(rsqrtf(ss), 1e12f);   // = 1 / max(sqrt(ss), 1e-12)
      }
    }
    i += U * stride;
    if (i < total) issue(i);
  }
}

// ---- RMSNorm ---------------------------------------------------------------------------------
// LP lanes cooperate on one pixel (LP = min(32, C/8)); each lane keeps its share of the row in
// registers (<= 4 vectors of 8 channels: C <= 1024).
template <int LP>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LP>
__global__ void __launch_bounds__(256) pixel_inv_norm_kernel(const bf16* __restrict__ x, float* __restrict__ inv,
                                                             int64_t M, int C) {
  pdl_wait();
  const int vec_per_pix = C / 8;
  const int sub = threadIdx.x % LP;
  const int64_t pix_per_block = blockDim.x / LP;
  for (int64_t m = (int64_t)blockIdx.x * pix_per_block + threadIdx.x / LP; m < M;
       m += (int64_t)gridDim.x * pix_per_block) {
    float ss = 0.f;
    for (int v = sub; v < vec_per_pix; v += LP) {
      float f[8];
      unpack8(ld_stream(x + m * C + v * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
    }
    ss = group_sum<LP>(ss);
    if (sub == 0) inv[m] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);                 // F.normalize eps (model.py:207)
  }
}

template <int LP, bool HAS_RES>
__global__ void __launch_bounds__(256) rmsnorm_residual_kernel(const bf16* x, const float* __restrict__ g,
                                                               const bf16* residual, bf16* y, int64_t M, int C,
                                                               float sqrt_c) {
  pdl_wait();
  const int vec_per_pix = C / 8;
  const int sub = threadIdx.x % LP;
  const int64_t pix_per_block = blockDim.x / LP;
  for (int64_t m = (int64_t)blockIdx.x * pix_per_block + threadIdx.x / LP; m < M;
       m += (int64_t)gridDim.x * pix_per_block) {
    float f[4][8];
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int v = sub + t * LP;
      if (v < vec_per_pix) {
        unpack8(ld_stream(x + m * C + v * 8), f[t]);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += f[t][j] * f[t][j];
      }
    }
    ss = group_sum<LP>(ss);
    const float scale = sqrt_c / fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int v = sub + t * LP;
      if (v < vec_per_pix) {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + v * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + v * 8 + 4));
        float o[8] = {f[t][0] * scale * g0.x, f[t][1] * scale * g0.y, f[t][2] * scale * g0.z, f[t][3] * scale * g0.w,
                      f[t][4] * scale * g1.x, f[t][5] * scale * g1.y, f[t][6] * scale * g1.z, f[t][7] * scale * g1.w};
        if (HAS_RES) {
          float r[8];
          unpack8(ld_stream(residual + m * C + v * 8), r);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += r[j];
        }
        st_stream(y + m * C + v * 8, pack8(o));
      }
    }
  }
}

static int gn_reverse() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("SRGD_GN_REVERSE");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached;
}

// CTAs per SM the gn_apply grid is capped at.  Every CTA pays the statistics fold + coefficient set-up once, so with
// fused statistics the grid is one resident wave (4 CTAs of 256 threads and <= 64 registers per SM) and each thread
// streams many vectors; SRGD_GN_CTAS_PER_SM overrides (A/B knob).
static int gn_ctas_per_sm() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("SRGD_GN_CTAS_PER_SM");
    cached = (e != nullptr && atoi(e) > 0) ? atoi(e) : 4;
  }
  return cached;
}

static int stream_grid(int64_t items_per_block_total, int ctas_per_sm) {
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (items_per_block_total < 1) items_per_block_total = 1;
  return (int)(items_per_block_total < cap ? items_per_block_total : cap);
}

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_groupnorm_finalize(const float* gn_partials, float* stats, int32_t B, int32_t H, int32_t W,
                                       int32_t C, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(gn_partials && stats && B > 0 && H > 0 && W > 0 && C > 0 && C % 64 == 0,
               "groupnorm_finalize: bad arguments");
  const TileGeom g = tile_geom(B, H, W);
  SRGD_REQUIRE(g.tn_log2 <= 2, "groupnorm_finalize: needs H*W >= 32");
  ProfScope prof(SRGD_PK_NORM_MISC, 0.0, (double)(g.m_tiles << g.tn_log2) * 8 * 2 * 4, as_stream(stream));
  SRGD_CUDA_OK(launch_k(gn_finalize_kernel, dim3(B), dim3(256), 0, as_stream(stream), gn_partials, stats, H, W, C, g));
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_groupnorm_stats(const void* x, float* stats, int32_t B, int32_t H, int32_t W, int32_t C,
                                    srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x && stats && B > 0 && H > 0 && W > 0 && C > 0 && C % 64 == 0, "groupnorm_stats: bad arguments");
  ProfScope prof(SRGD_PK_NORM_MISC, 0.0, (double)B * H * W * C * 2.0, as_stream(stream));
  gn_stats_kernel<<<B * kGroups, 512, 0, as_stream(stream)>>>(reinterpret_cast<const bf16*>(x), stats, H * W, C);
  SRGD_LAUNCH_OK("gn_stats_kernel");
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_groupnorm_apply(const void* x, int32_t Bx, const float* stats, const float* gn_partials,
                                    const float* gamma, const float* beta, const float* scale_shift, int64_t ss_stride,
                                    const void* residual, void* y, float* inv_out, int32_t B, int32_t H, int32_t W,
                                    int32_t C, srgd_stream_t stream) {
  return srgd_groupnorm_apply_ex(x, Bx, stats, gn_partials, gamma, beta, scale_shift, ss_stride, residual, B, y, inv_out,
                                 B, H, W, C, stream);
}

extern "C" int srgd_groupnorm_apply_ex(const void* x, int32_t Bx, const float* stats, const float* gn_partials,
                                       const float* gamma, const float* beta, const float* scale_shift,
                                       int64_t ss_stride, const void* residual, int32_t res_rows, void* y,
                                       float* inv_out, int32_t B, int32_t H, int32_t W, int32_t C,
                                       srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(residual == nullptr || (res_rows > 0 && res_rows <= B && B % res_rows == 0),
               "groupnorm_apply: B=%d must be a multiple of the residual's rows (%d)", B, res_rows);
  if (residual == nullptr) res_rows = 1;
  SRGD_REQUIRE(x && gamma && beta && y, "groupnorm_apply: null argument");
  SRGD_REQUIRE((stats != nullptr) != (gn_partials != nullptr),
               "groupnorm_apply: pass exactly one of stats / gn_partials");
  const TileGeom geom = tile_geom(Bx, H, W);             // geometry of the conv launch that produced the partials
  SRGD_REQUIRE(gn_partials == nullptr || geom.tn_log2 <= 2, "groupnorm_apply: gn_partials need H*W >= 32");
  SRGD_REQUIRE(B > 0 && Bx > 0 && Bx <= B && H > 0 && W > 0 && C > 0 && C % 64 == 0 && C <= 4096,
               "groupnorm_apply: bad shape B=%d Bx=%d H=%d W=%d C=%d", B, Bx, H, W, C);
  SRGD_REQUIRE(B <= 65535, "groupnorm_apply: B too large");
  const int64_t total = (int64_t)H * W * (C / 8);
  int gx = stream_grid((total + 256 * 4 - 1) / (256 * 4), 8);
  const int cap = sm_count() * gn_ctas_per_sm();
  if ((int64_t)gx * B > cap) gx = (cap + B - 1) / B;
  dim3 grid(gx, B);
  const size_t smem = (size_t)C * 2 * sizeof(float);
  const bf16* xr = reinterpret_cast<const bf16*>(x);
  const bf16* rr = reinterpret_cast<const bf16*>(residual);
  bf16* yr = reinterpret_cast<bf16*>(y);
  ProfScope prof(SRGD_PK_GN_APPLY, 0.0, (double)B * H * W * C * (residual ? 6.0 : 4.0), as_stream(stream));
  SRGD_REQUIRE(inv_out == nullptr || (residual != nullptr && (C == 128 || C == 256) && (H * W) % 4 == 0),
               "groupnorm_apply: inv_out needs the residual variant, C in {128, 256} and H*W %% 4 == 0");
  cudaStream_t cst = as_stream(stream);
  const int vpp = C / 8;
  const bool reg = (vpp & (vpp - 1)) == 0 && vpp <= 256;
#define SRGD_GN_LAUNCH(RES, INV, R)                                                                              \
  SRGD_CUDA_OK(launch_k(gn_apply_kernel<RES, INV, false, R, RES ? 2 : 4>, grid, dim3(256), R ? 0 : smem, cst, xr, Bx, \
                        stats, gamma, beta, scale_shift, ss_stride, rr, yr, inv_out, nullptr, nullptr, nullptr,   \
                        H * W, C, gn_reverse(), gn_partials, geom, res_rows))
  if (inv_out != nullptr && C == 128) SRGD_GN_LAUNCH(true, 16, true);
  else if (inv_out != nullptr) SRGD_GN_LAUNCH(true, 32, true);
  else if (residual && reg) SRGD_GN_LAUNCH(true, 0, true);
  else if (residual) SRGD_GN_LAUNCH(true, 0, false);
  else if (reg) SRGD_GN_LAUNCH(false, 0, true);
  else SRGD_GN_LAUNCH(false, 0, false);
#undef SRGD_GN_LAUNCH
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_groupnorm_apply_final(const void* x, const float* stats, const float* gn_partials,
                                          const float* gamma, const float* beta,
                                          const void* residual, const float* final_w, const float* final_b,
                                          float* eps, int32_t B, int32_t H, int32_t W, int32_t C,
                                          srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x && gamma && beta && residual && final_w && final_b && eps, "groupnorm_apply_final: null argument");
  SRGD_REQUIRE((stats != nullptr) != (gn_partials != nullptr),
               "groupnorm_apply_final: pass exactly one of stats / gn_partials");
  const TileGeom geom = tile_geom(B, H, W);
  SRGD_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && C == 128 && (H * W) % 2 == 0,
               "groupnorm_apply_final: needs C == 128 and an even pixel count (B=%d H=%d W=%d C=%d)", B, H, W, C);
  const int64_t total = (int64_t)H * W * (C / 8);
  int gx = stream_grid((total + 256 * 4 - 1) / (256 * 4), 8);
  const int cap = sm_count() * gn_ctas_per_sm();
  if ((int64_t)gx * B > cap) gx = (cap + B - 1) / B;
  dim3 grid(gx, B);
  const size_t smem = (size_t)C * 5 * sizeof(float);
  ProfScope prof(SRGD_PK_GN_APPLY, 2.0 * B * H * W * C * 3, (double)B * H * W * (C * 4.0 + 12.0), as_stream(stream));
  (void)smem;
  SRGD_CUDA_OK(launch_k(gn_apply_kernel<true, 0, true, true, 2>, grid, dim3(256), 0, as_stream(stream),
                        reinterpret_cast<const bf16*>(x), B, stats, gamma, beta, nullptr, 0,
                        reinterpret_cast<const bf16*>(residual), nullptr, nullptr, final_w, final_b, eps, H * W, C,
                        gn_reverse(), gn_partials, geom, B));
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_pixel_inv_norm(const void* x, float* inv, int64_t M, int32_t C, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x && inv && M > 0 && C >= 64 && C % 64 == 0, "pixel_inv_norm: bad arguments");
  const bf16* xr = reinterpret_cast<const bf16*>(x);
  ProfScope prof(SRGD_PK_NORM_MISC, 0.0, (double)M * C * 2.0 + (double)M * 4.0, as_stream(stream));
  if (C / 8 >= 32) {
    const int grid = stream_grid((M + 7) / 8, 8);
    SRGD_CUDA_OK(launch_k(pixel_inv_norm_kernel<32>, dim3(grid), dim3(256), 0, as_stream(stream), xr, inv, M, C));
  } else if (C / 8 == 16) {
    const int grid = stream_grid((M + 15) / 16, 8);
    SRGD_CUDA_OK(launch_k(pixel_inv_norm_kernel<16>, dim3(grid), dim3(256), 0, as_stream(stream), xr, inv, M, C));
  } else {
    const int grid = stream_grid((M + 31) / 32, 8);
    SRGD_CUDA_OK(launch_k(pixel_inv_norm_kernel<8>, dim3(grid), dim3(256), 0, as_stream(stream), xr, inv, M, C));
  }
  SRGD_LAUNCH_OK("pixel_inv_norm_kernel");
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_rmsnorm_residual(const void* x, const float* g, const void* residual, void* y, int64_t M,
                                     int32_t C, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x && g && y && M > 0 && C >= 64 && C % 64 == 0 && C <= 1024, "rmsnorm_residual: bad arguments (C=%d)", C);
  const bf16* xr = reinterpret_cast<const bf16*>(x);
  const bf16* rr = reinterpret_cast<const bf16*>(residual);
  bf16* yr = reinterpret_cast<bf16*>(y);
  const float sc = sqrtf((float)C);
  cudaStream_t st = as_stream(stream);
  ProfScope prof(SRGD_PK_NORM_MISC, 0.0, (double)M * C * (residual ? 6.0 : 4.0), st);
#define SRGD_RMS(LP)                                                                                     \
  do {                                                                                                   \
    const int grid = stream_grid((M + (256 / LP) - 1) / (256 / LP), 8);                                  \
    if (residual)                                                                                        \
      SRGD_CUDA_OK(launch_k(rmsnorm_residual_kernel<LP, true>, dim3(grid), dim3(256), 0, st, xr, g, rr, yr, M, C, sc)); \
    else                                                                                                 \
      SRGD_CUDA_OK(launch_k(rmsnorm_residual_kernel<LP, false>, dim3(grid), dim3(256), 0, st, xr, g, rr, yr, M, C, sc)); \
  } while (0)
  if (C / 8 >= 32) SRGD_RMS(32);
  else if (C / 8 == 16) SRGD_RMS(16);
  else SRGD_RMS(8);
#undef SRGD_RMS
  SRGD_LAUNCH_OK("rmsnorm_residual_kernel");
  count_launch();
  return SRGD_OK;
}
