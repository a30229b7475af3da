// Fused LinearAttention block on tcgen05 tensor cores (reference: LinearAttention.forward,
// model.py:307-324, plus the caller's residual `attn(x) + x`, model.py:703/718):
//
//     xn = RMSNorm(x)                               (model.py:310; gain folded into the qkv weight)
//     q,k,v = to_qkv(xn)                            1x1 conv, no bias (model.py:311-313)
//     q = softmax_d(q) * 32^-1/2 ; k = softmax_n(k) (model.py:315-318)
//     ctx[d][e] = sum_n k[d][n] v[e][n]             (model.py:320)
//     o[e][n]   = sum_d ctx[d][e] q[d][n]           (model.py:322)
//     y = RMSNorm(to_out(o)) + x                    (model.py:303-304, 324, 703)
//
// The unfused path (conv_igemm.cu + attention.cu + norm.cu) moves ~3.8 KB per pixel through HBM at
// C = 128 (qkv tensor written and re-read, o, y before the norm ...).  Here the block is two
// persistent tcgen05 kernels that keep q/k/v/o entirely on chip:
//
//   la_ctx_kernel  : per 128-pixel tile  K^T,V^T[128 ch][128 px] = W_{k,v} x^T  (channels as the MMA M
//                    dimension, so softmax over n is thread-local: TMEM lane = channel), P = exp(k - max)
//                    and V^T go to shared memory as bf16 K-major operands, ctx_tile = P V via a second
//                    MMA, merged into per-thread running (max, Z, ctx[32]) -- the online softmax.
//                    Reads x once (2C B/pixel), writes 17 KB per CTA.
//   la_merge_kernel: merges the per-CTA partials into bf16 block-diagonal matrices [B][128][128].
//   la_out_kernel  : per 128-pixel tile  Q = x Wq^T (pixels as M: softmax over d is thread-local),
//                    O = softmax(Q) * blockdiag(ctx), Y = O Wout^T, then bias + RMSNorm + residual in
//                    the epilogue.  Reads x (2C B/pixel + an L2-hot re-read for the residual), writes y.
//
// Supported: heads*dim_head = 128, C in {128, 256}, N % 128 == 0; everything else takes the unfused path.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace srgd {

constexpr int kLfHid = 128;             // heads * dim_head
constexpr float kLfQScale = 0.17677669529663687f;   // 32^-1/2 (model.py:295, 318)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint8_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}
__device__ __forceinline__ uint4 ld_shared_v4(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

// ---------------------------------------------------------------------------------------------
// kernel A: context partials
// ---------------------------------------------------------------------------------------------
constexpr int kCtxStages = 3;
struct alignas(64) LaCtxParams {
  CUtensorMap x_map;                    // bf16 [B*N][C], box {64, 128}
  CUtensorMap w_map;                    // bf16 [384][C] (q | k | v rows), box {64, 128}
  const float* inv;                     // [B*N] 1/||x||
  float* part;                          // [B][splits][34][128] = m, Z, ctx[32] per channel
  int32_t C, splits, tiles_per_sample;
};
struct LaCtxSmem {
  static constexpr int kStageBytes = 3 * 16384;              // Wk | Wv | x, one 64-wide k-block each
  static constexpr int kPOffset = kCtxStages * kStageBytes;
  static constexpr int kVtOffset = kPOffset + 32768;
  static constexpr int kInvOffset = kVtOffset + 32768;       // float [2][128]
  static constexpr int kBarOffset = kInvOffset + 1024;
  static constexpr int kTotal = kBarOffset + 128 + 1024;
};

__global__ void __launch_bounds__(320, 1) la_ctx_kernel(const __grid_constant__ LaCtxParams p) {
  using L = LaCtxSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kCtxStages;
  uint64_t* kv_full = empty_bar + kCtxStages;
  uint64_t* kv_empty = kv_full + 1;
  uint64_t* pv_ready = kv_empty + 1;
  uint64_t* ctx_full = pv_ready + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ctx_full + 1);
  float* inv_s = reinterpret_cast<float*>(smem + L::kInvOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, sp = blockIdx.x;
  const int per = (p.tiles_per_sample + p.splits - 1) / p.splits;
  const int t0 = min(sp * per, p.tiles_per_sample), t1 = min(t0 + per, p.tiles_per_sample);
  const int ntiles = t1 - t0;
  const int kblocks = p.C / 64;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kCtxStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(kv_full, 1);
    ptx::mbar_init(kv_empty, 8);
    ptx::mbar_init(pv_ready, 8);
    ptx::mbar_init(ctx_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&p.x_map);
      ptx::prefetch_tmap(&p.w_map);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr_smem, 512);                  // K^T [0,128) | V^T [128,256) | ctx tile [256,384)
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < ntiles; ++it) {
      const int row0 = (b * p.tiles_per_sample + t0 + it) * 128;
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* dst = smem + stage * L::kStageBytes;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        ptx::tma_load_2d(dst, &p.w_map, &full_bar[stage], kb * 64, 128);           // W_k rows
        ptx::tma_load_2d(dst + 16384, &p.w_map, &full_bar[stage], kb * 64, 256);   // W_v rows
        ptx::tma_load_2d(dst + 32768, &p.x_map, &full_bar[stage], kb * 64, row0);
        if (++stage == kCtxStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ====================================== MMA issuer ======================================
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t par = it & 1;
      ptx::mbar_wait(kv_empty, par ^ 1);                  // epilogue has read the previous K^T / V^T
      ptx::tc_fence_after();
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t base = ptx::smem_u32(smem + stage * L::kStageBytes);
        const uint64_t wk = ptx::make_sw128_kmajor_desc(base);
        const uint64_t wv = ptx::make_sw128_kmajor_desc(base + 16384);
        const uint64_t xd = ptx::make_sw128_kmajor_desc(base + 32768);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
          ptx::umma_bf16_ss(tmem_base, wk + 2 * k, xd + 2 * k, idesc, accum);
          ptx::umma_bf16_ss(tmem_base + 128, wv + 2 * k, xd + 2 * k, idesc, accum);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (kb == kblocks - 1) ptx::umma_commit(kv_full);
        if (++stage == kCtxStages) { stage = 0; phase ^= 1; }
      }
      // ctx_tile[(h,d)][(h',e)] = sum_px P[(h,d)][px] * Vt[(h',e)][px]
      ptx::mbar_wait(pv_ready, par);
      ptx::tc_fence_after();
      const uint32_t pb = ptx::smem_u32(smem + L::kPOffset), vb = ptx::smem_u32(smem + L::kVtOffset);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t pd = ptx::make_sw128_kmajor_desc(pb + (ks >> 2) * 16384) + 2 * (ks & 3);
        const uint64_t vd = ptx::make_sw128_kmajor_desc(vb + (ks >> 2) * 16384) + 2 * (ks & 3);
        ptx::umma_bf16_ss(tmem_base + 256, pd, vd, idesc, ks != 0 ? 1u : 0u);
      }
      ptx::umma_commit(ctx_full);
    }
  } else if (warp >= 2) {
    // ================================== softmax / operand warps ==================================
    const bool is_k = warp < 6;                            // warps 2-5: K^T rows, warps 6-9: V^T rows
    const int q = warp & 3;                                // TMEM lane quarter (= head)
    const int row = q * 32 + lane;                         // channel (h,d) resp. (h,e)
    const int et = threadIdx.x - 64;                       // 0..255
    const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
    uint8_t* tile_smem = smem + (is_k ? L::kPOffset : L::kVtOffset);
    float m_run = -INFINITY, z_run = 0.f;
    float ctx[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) ctx[e] = 0.f;

    for (int it = 0; it < ntiles; ++it) {
      const uint32_t par = it & 1;
      const int row0 = (b * p.tiles_per_sample + t0 + it) * 128;
      float* invb = inv_s + par * 128;
      if (et < 128) invb[et] = p.inv[row0 + et];
      named_bar_sync(1, 256);
      ptx::mbar_wait(kv_full, par);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + lane_bits + (is_k ? 0u : 128u);
      float m_t = -INFINITY, z_t = 0.f;
      if (is_k) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(taddr + c * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) m_t = fmaxf(m_t, __uint_as_float(v[j]) * invb[c * 32 + j]);
        }
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(taddr + c * 32, v);
        ptx::tmem_ld_wait();
        uint32_t w[16];
        if (is_k) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = __expf(__uint_as_float(v[j]) * invb[c * 32 + j] - m_t);
            const float p1 = __expf(__uint_as_float(v[j + 1]) * invb[c * 32 + j + 1] - m_t);
            z_t += p0 + p1;
            w[j >> 1] = pack_bf16(p0, p1);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            w[j >> 1] = pack_bf16(__uint_as_float(v[j]) * invb[c * 32 + j],
                                  __uint_as_float(v[j + 1]) * invb[c * 32 + j + 1]);
        }
        uint8_t* kbase = tile_smem + (c >> 1) * 16384;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          st_shared_v4(kbase + ptx::sw128_offset(row, (c & 1) * 4 + jj), w[4 * jj], w[4 * jj + 1], w[4 * jj + 2],
                       w[4 * jj + 3]);
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive(kv_empty);
        ptx::mbar_arrive(pv_ready);
      }
      if (is_k) {
        ptx::mbar_wait(ctx_full, par);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + lane_bits + 256 + q * 32, v);    // diagonal 32x32 block of head q
        ptx::tmem_ld_wait();
        const float m_new = fmaxf(m_run, m_t);
        const float so = __expf(m_run - m_new), sn = __expf(m_t - m_new);
        z_run = z_run * so + z_t * sn;
#pragma unroll
        for (int e = 0; e < 32; ++e) ctx[e] = ctx[e] * so + __uint_as_float(v[e]) * sn;
        m_run = m_new;
        ptx::tc_fence_before();
      }
    }
    if (is_k) {
      float* dst = p.part + ((int64_t)b * p.splits + sp) * (34 * kLfHid) + row;   // record layout [34][128]
      dst[0] = m_run;
      dst[kLfHid] = z_run;
#pragma unroll
      for (int e = 0; e < 32; ++e) dst[(2 + e) * kLfHid] = ctx[e];
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// merge: partials -> per-sample output matrix
//     Mb[b][c][(h,d)] = sum_e Wout[c][(h,e)] * 32^-1/2 * ctx_h[d][e] / Z_h[d]          (bf16 [B][C][128])
// i.e. to_out applied to the normalised context once per sample, so that the output kernel needs a single
// GEMM  y = softmax_d(q) Mb^T  instead of  o = softmax_d(q) ctx ; y = o Wout^T  (model.py:322-324).
// block = (b, h), 128 threads: d = t & 31, quarter t >> 5 of the C output channels.
// ---------------------------------------------------------------------------------------------
constexpr int kMergeWarps = 16;
__global__ void __launch_bounds__(kMergeWarps * 32) la_merge_mb_kernel(const float* __restrict__ part,
                                                                      const bf16* __restrict__ wout,
                                                                      bf16* __restrict__ mb, int splits, int C) {
  extern __shared__ uint8_t merge_smem[];                  // Wout[:, h*32:(h+1)*32] as bf16 [C][32] | float [16][34][32]
  const int b = blockIdx.x >> 2, h = blockIdx.x & 3;
  const int d = threadIdx.x & 31, cq = threadIdx.x >> 5;
  uint4* wsm = reinterpret_cast<uint4*>(merge_smem);
  float* red = reinterpret_cast<float*>(merge_smem + (size_t)C * 64);
  for (int i = threadIdx.x; i < C * 4; i += kMergeWarps * 32)   // 4 x 16 B per output channel
    wsm[i] = __ldg(reinterpret_cast<const uint4*>(wout + (int64_t)(i >> 2) * kLfHid + h * 32) + (i & 3));
  pdl_wait();                                             // weights above are never written by a kernel
  // Each of the 16 warps folds every sixteenth split record (a one-tile launch has 2 x 148 of them per sample: with
  // four warps this serial chain of L2 round trips was the longest kernel of the block); the sixteen partial results
  // are combined through shared memory in a fixed order, so the result does not depend on timing.
  const float* base = part + (int64_t)b * splits * (34 * kLfHid) + h * 32 + d;
  const int64_t sstride = 34 * kLfHid;
  float m = -INFINITY;
  for (int s = cq; s < splits; s += kMergeWarps) m = fmaxf(m, __ldg(base + s * sstride));
  float z = 0.f, acc[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) acc[e] = 0.f;
  for (int s = cq; s < splits; s += kMergeWarps) {
    const float* src = base + s * sstride;
    const float ms = __ldg(src);
    const float w = (ms == -INFINITY) ? 0.f : __expf(ms - m);   // empty records (CTA without tiles) carry m = -inf
    z = fmaf(w, __ldg(src + kLfHid), z);
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = fmaf(w, __ldg(src + (2 + e) * kLfHid), acc[e]);
  }
  float* mine = red + (cq * 34) * 32 + d;
  mine[0] = m;
  mine[32] = z;
#pragma unroll
  for (int e = 0; e < 32; ++e) mine[(2 + e) * 32] = acc[e];
  __syncthreads();
  float mm = -INFINITY;
#pragma unroll
  for (int k = 0; k < kMergeWarps; ++k) mm = fmaxf(mm, red[(k * 34) * 32 + d]);
  z = 0.f;
#pragma unroll
  for (int e = 0; e < 32; ++e) acc[e] = 0.f;
#pragma unroll 4
  for (int k = 0; k < kMergeWarps; ++k) {
    const float* r = red + (k * 34) * 32 + d;
    const float mk = r[0];
    const float w = (mk == -INFINITY) ? 0.f : __expf(mk - mm);
    z = fmaf(w, r[32], z);
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = fmaf(w, r[(2 + e) * 32], acc[e]);
  }
  const float inv = kLfQScale / z;
#pragma unroll
  for (int e = 0; e < 32; ++e) acc[e] *= inv;
  const int cper = C / kMergeWarps;
#pragma unroll 2
  for (int c = cq * cper; c < (cq + 1) * cper; ++c) {
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float f[8];
      unpack8(wsm[c * 4 + t], f);                          // same address across the warp: broadcast
#pragma unroll
      for (int j = 0; j < 8; ++j) sum = fmaf(f[j], acc[t * 8 + j], sum);
    }
    mb[((int64_t)b * C + c) * kLfHid + h * 32 + d] = __float2bfloat16(sum);
  }
}

// host launcher shared with linattn_pp.cu: > 48 KB of dynamic shared memory needs the per-device opt-in
int launch_la_merge(const float* part, const bf16* wout, bf16* mb, int B, int splits, int C, cudaStream_t st) {
  const size_t smem = (size_t)C * 64 + (size_t)kMergeWarps * 34 * 32 * 4;
  static uint64_t configured = 0;
  if (first_launch_on_device(configured)) {
    SRGD_CUDA_OK(cudaFuncSetAttribute(la_merge_mb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      512 * 64 + kMergeWarps * 34 * 32 * 4));
  }
  SRGD_CUDA_OK(launch_k(la_merge_mb_kernel, dim3(B * 4), dim3(kMergeWarps * 32), smem, st, part, wout, mb, splits, C));
  return SRGD_OK;
}

// ---------------------------------------------------------------------------------------------
// kernel B: q softmax, context apply, to_out, RMSNorm, residual
// ---------------------------------------------------------------------------------------------
struct alignas(64) LaOutParams {
  CUtensorMap x_map;                    // bf16 [B*N][C], box {64, 128}
  CUtensorMap w_map;                    // bf16 [384][C], box {64, 128} (rows 0..127 = W_q)
  CUtensorMap mb_map;                   // bf16 [B*C][128] (la_merge_mb_kernel), box {64, C}
  const float* inv;
  const float* bias;
  const float* g;
  const bf16* x;
  bf16* out;
  int32_t chunks, tiles_per_sample;
};
template <int C>
struct LaOutSmem {
  static constexpr int kStages = (C == 128) ? 3 : 2;
  static constexpr int kStageBytes = 2 * 16384;              // x | Wq, one k-block each
  static constexpr int kQsOffset = kStages * kStageBytes;    // softmax(q) as bf16 A operand
  static constexpr int kMbOffset = kQsOffset + 32768;        // 2 k-blocks of [C][64]
  static constexpr int kSsqOffset = kMbOffset + 2 * C * 128;     // float [2][2][128]
  static constexpr int kBarOffset = kSsqOffset + 2048;
  static constexpr int kTotal = kBarOffset + 128 + 1024;
};

template <int C>
__global__ void __launch_bounds__(320, 1) la_out_kernel(const __grid_constant__ LaOutParams p) {
  using L = LaOutSmem<C>;
  constexpr int kStages = L::kStages;
  constexpr int kblocks = C / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* const_full = empty_bar + kStages;
  uint64_t* q_full = const_full + 1;
  uint64_t* qs_ready = q_full + 1;
  uint64_t* y_full = qs_ready + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(y_full + 1);
  float* ssq_s = reinterpret_cast<float*>(smem + L::kSsqOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, ck = blockIdx.x;
  const int per = (p.tiles_per_sample + p.chunks - 1) / p.chunks;
  const int t0 = min(ck * per, p.tiles_per_sample), t1 = min(t0 + per, p.tiles_per_sample);
  const int ntiles = t1 - t0;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(const_full, 1);
    ptx::mbar_init(q_full, 1);
    ptx::mbar_init(qs_ready, 8);
    ptx::mbar_init(y_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&p.x_map);
      ptx::prefetch_tmap(&p.w_map);
      ptx::prefetch_tmap(&p.mb_map);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr_smem, 512);                  // Q [0,128) | Y [256,256+C)
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    if (ntiles > 0) {
      ptx::mbar_arrive_expect_tx(const_full, 2 * C * 128);
      ptx::tma_load_2d(smem + L::kMbOffset, &p.mb_map, const_full, 0, b * C);
      ptx::tma_load_2d(smem + L::kMbOffset + C * 128, &p.mb_map, const_full, 64, b * C);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < ntiles; ++it) {
      const int row0 = (b * p.tiles_per_sample + t0 + it) * 128;
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* dst = smem + stage * L::kStageBytes;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
        ptx::tma_load_2d(dst, &p.x_map, &full_bar[stage], kb * 64, row0);
        ptx::tma_load_2d(dst + 16384, &p.w_map, &full_bar[stage], kb * 64, 0);     // W_q rows
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ====================================== MMA issuer ======================================
    constexpr uint32_t idesc128 = ptx::make_idesc_bf16_f32(128, 128);
    constexpr uint32_t idescC = ptx::make_idesc_bf16_f32(128, C);
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t qs = ptx::smem_u32(smem + L::kQsOffset), wo = ptx::smem_u32(smem + L::kMbOffset);
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t par = it & 1;
      // Q[px][(h,d)] = x Wq^T.  (The previous tile's Q/O/Y accumulators were drained before the epilogue
      // signalled qs_ready for this tile, which this thread observes below in program order.)
      for (int kb = 0; kb < kblocks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t base = ptx::smem_u32(smem + stage * L::kStageBytes);
        const uint64_t xd = ptx::make_sw128_kmajor_desc(base);
        const uint64_t wq = ptx::make_sw128_kmajor_desc(base + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss(tmem_base, xd + 2 * k, wq + 2 * k, idesc128, (kb | k) != 0 ? 1u : 0u);
        ptx::umma_commit(&empty_bar[stage]);
        if (kb == kblocks - 1) ptx::umma_commit(q_full);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (it == 0) ptx::mbar_wait(const_full, 0);
      // Y[px][c] = softmax(q)[px][(h,d)] * Mb[c][(h,d)]^T
      ptx::mbar_wait(qs_ready, par);
      ptx::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t ad = ptx::make_sw128_kmajor_desc(qs + (ks >> 2) * 16384) + 2 * (ks & 3);
        const uint64_t bd = ptx::make_sw128_kmajor_desc(wo + (ks >> 2) * (C * 128)) + 2 * (ks & 3);
        ptx::umma_bf16_ss(tmem_base + 256, ad, bd, idescC, ks != 0 ? 1u : 0u);
      }
      ptx::umma_commit(y_full);
    }
  } else if (warp >= 2) {
    // ======================================= epilogue =======================================
    // 8 warps: lane quarter q = warp & 3 selects the 32 pixels, `half` the column half this warp owns.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
    uint8_t* qs_smem = smem + L::kQsOffset + half * 16384;   // k-block `half` of the A operand tile
    const float sqrt_c = sqrtf((float)C);
    constexpr int kYChunks = C / 64;                         // 32-column chunks per half

    for (int it = 0; it < ntiles; ++it) {
      const uint32_t par = it & 1;
      const int64_t px = (int64_t)(b * p.tiles_per_sample + t0 + it) * 128 + row;
      const float my_inv = p.inv[px];

      // ---- q: softmax over the 32 channels of each head (model.py:315) ----
      ptx::mbar_wait(q_full, par);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + lane_bits + (half * 2 + i) * 32, v);
        ptx::tmem_ld_wait();
        float f[32];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __uint_as_float(v[j]) * my_inv;
          mx = fmaxf(mx, f[j]);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __expf(f[j] - mx);
          sum += f[j];
        }
        const float rs = 1.0f / sum;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          st_shared_v4(qs_smem + ptx::sw128_offset(row, i * 4 + jj), pack_bf16(f[8 * jj] * rs, f[8 * jj + 1] * rs),
                       pack_bf16(f[8 * jj + 2] * rs, f[8 * jj + 3] * rs),
                       pack_bf16(f[8 * jj + 4] * rs, f[8 * jj + 5] * rs),
                       pack_bf16(f[8 * jj + 6] * rs, f[8 * jj + 7] * rs));
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(qs_ready);

      // ---- y: + bias, RMSNorm over all C channels of the pixel (model.py:207), * g, + x ----
      ptx::mbar_wait(y_full, par);
      ptx::tc_fence_after();
      const uint32_t ycol0 = 256 + half * (C / 2);
      const int c_first = half * (C / 2);
      float ssq = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < kYChunks; ++cc) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + lane_bits + ycol0 + cc * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + c_first + cc * 32 + j));
          const float y0 = __uint_as_float(v[j]) + bv.x, y1 = __uint_as_float(v[j + 1]) + bv.y;
          const float y2 = __uint_as_float(v[j + 2]) + bv.z, y3 = __uint_as_float(v[j + 3]) + bv.w;
          ssq += y0 * y0 + y1 * y1 + y2 * y2 + y3 * y3;
        }
      }
      float* ssq_t = ssq_s + par * 256;
      ssq_t[half * 128 + row] = ssq;
      named_bar_sync(1, 256);
      const float tot = ssq_t[row] + ssq_t[128 + row];
      const float scale = sqrt_c / fmaxf(sqrtf(tot), 1e-12f);
      // Residual rows in and output rows out go through this warp's own 4 KB of the softmax(q) operand tile (free once
      // the Y MMA has completed), 64 columns at a time: global accesses are made with eight lanes per row (128 bytes),
      // so an instruction touches 4 full lines instead of 32 half sectors, and the tile transposes between that
      // layout and one row per thread.
      const int64_t px0 = (int64_t)(b * p.tiles_per_sample + t0 + it) * 128 + q * 32;
      const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll 1
      for (int cp = 0; cp < kYChunks / 2; ++cp) {
        uint4 xr[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          xr[k] = *reinterpret_cast<const uint4*>(p.x + (px0 + k * 4 + sub) * C + c_first + cp * 64 + chunk * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          st_shared_v4(qs_smem + ptx::sw128_offset(q * 32 + k * 4 + sub, chunk), xr[k].x, xr[k].y, xr[k].z, xr[k].w);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t) xr[t] = ld_shared_v4(qs_smem + ptx::sw128_offset(row, t));
        __syncwarp();
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const int cc = cp * 2 + c2;
          uint32_t v[32];
          ptx::tmem_ld_32x32(tmem_base + lane_bits + ycol0 + cc * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float r[8], o[8];
            unpack8(xr[c2 * 4 + jj], r);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c_first + cc * 32 + jj * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c_first + cc * 32 + jj * 8 + 4));
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.g + c_first + cc * 32 + jj * 8));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.g + c_first + cc * 32 + jj * 8 + 4));
            o[0] = (__uint_as_float(v[jj * 8 + 0]) + b0.x) * scale * g0.x + r[0];
            o[1] = (__uint_as_float(v[jj * 8 + 1]) + b0.y) * scale * g0.y + r[1];
            o[2] = (__uint_as_float(v[jj * 8 + 2]) + b0.z) * scale * g0.z + r[2];
            o[3] = (__uint_as_float(v[jj * 8 + 3]) + b0.w) * scale * g0.w + r[3];
            o[4] = (__uint_as_float(v[jj * 8 + 4]) + b1.x) * scale * g1.x + r[4];
            o[5] = (__uint_as_float(v[jj * 8 + 5]) + b1.y) * scale * g1.y + r[5];
            o[6] = (__uint_as_float(v[jj * 8 + 6]) + b1.z) * scale * g1.z + r[6];
            o[7] = (__uint_as_float(v[jj * 8 + 7]) + b1.w) * scale * g1.w + r[7];
            const uint4 pk = pack8(o);
            st_shared_v4(qs_smem + ptx::sw128_offset(row, c2 * 4 + jj), pk.x, pk.y, pk.z, pk.w);
          }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rl = k * 4 + sub;
          const uint4 pk = ld_shared_v4(qs_smem + ptx::sw128_offset(q * 32 + rl, chunk));
          st_stream(p.out + (px0 + rl) * C + c_first + cp * 64 + chunk * 8, pk);
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

int launch_la_block_pp(const void* x, const void* qkv_w, const void* out_w, const float* out_b, const float* out_g,
                       void* out, int B, int N, const float* inv, float* part, bf16* bd, int splits, cudaStream_t st);

static int la_splits_mode(int B, int tiles_per_sample, bool invariant) {
  int s = sm_count() / (invariant ? kInvariantRefBatch : B);
  if (s < 1) s = 1;
  if (s > tiles_per_sample) s = tiles_per_sample;
  return s;
}
static int la_splits_for(int B, int tiles_per_sample) { return la_splits_mode(B, tiles_per_sample, g_batch_invariant != 0); }
// workspace sizing covers both modes: the flag may change between srgd_unet_workspace_bytes and the forward
static int la_splits_max(int B, int tiles_per_sample) {
  const int a = la_splits_mode(B, tiles_per_sample, false), b = la_splits_mode(B, tiles_per_sample, true);
  return a > b ? a : b;
}

template <int C>
static int launch_la_out(const LaOutParams& kp, int chunks, int B, cudaStream_t st) {
  using L = LaOutSmem<C>;
  static uint64_t configured = 0;
  if (first_launch_on_device(configured)) {
    SRGD_CUDA_OK(cudaFuncSetAttribute(la_out_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  }
  SRGD_CUDA_OK(launch_k(la_out_kernel<C>, dim3(chunks, B), dim3(320), L::kTotal, st, kp));
  SRGD_LAUNCH_OK("la_out_kernel");
  count_launch();
  return SRGD_OK;
}

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_linear_attention_block_supported(int32_t N, int32_t C, int32_t heads) {
  return (heads == 4 && (C == 128 || C == 256) && N > 0 && N % 128 == 0) ? 1 : 0;
}

extern "C" size_t srgd_linear_attention_block_workspace(int32_t B, int32_t N, int32_t C, int32_t heads) {
  if (!srgd_linear_attention_block_supported(N, C, heads) || B <= 0) return 0;
  const int splits = la_splits_max(B, N / 128);
  // C = 128 runs two pipelines per CTA (linattn_pp.cu): two partial records per split
  return (size_t)B * N * sizeof(float) + (size_t)B * 2 * splits * kLfHid * 34 * sizeof(float) +
         (size_t)B * C * kLfHid * 2 + 1024;
}

extern "C" int srgd_linear_attention_block(const void* x, const float* inv_norm, const void* qkv_w, const void* out_w,
                                           const float* out_b, const float* out_g, void* out, int32_t B, int32_t N,
                                           int32_t C, int32_t heads, void* workspace, size_t workspace_bytes,
                                           srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x && qkv_w && out_w && out_b && out_g && out && workspace && B > 0, "linear_attention_block: null argument");
  SRGD_REQUIRE(srgd_linear_attention_block_supported(N, C, heads),
               "linear_attention_block: unsupported shape N=%d C=%d heads=%d (need heads=4, C in {128,256}, N %% 128 == 0)",
               N, C, heads);
  SRGD_REQUIRE(B <= 65535, "linear_attention_block: B too large");
  SRGD_REQUIRE(((uintptr_t)x | (uintptr_t)qkv_w | (uintptr_t)out_w | (uintptr_t)out | (uintptr_t)workspace) % 16 == 0,
               "linear_attention_block: pointers must be 16-byte aligned");
  if (workspace_bytes < srgd_linear_attention_block_workspace(B, N, C, heads)) {
    set_error("linear_attention_block: workspace too small");
    return SRGD_E_WORKSPACE;
  }
  const int tiles = N / 128;
  const int splits = la_splits_for(B, tiles);
  const int64_t M = (int64_t)B * N;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* inv = reinterpret_cast<float*>(ws);
  size_t off = ((size_t)M * sizeof(float) + 255) & ~(size_t)255;
  float* part = reinterpret_cast<float*>(ws + off);
  off += ((size_t)B * 2 * splits * kLfHid * 34 * sizeof(float) + 255) & ~(size_t)255;
  bf16* bd = reinterpret_cast<bf16*>(ws + off);
  cudaStream_t st = as_stream(stream);

  if (inv_norm != nullptr) {
    inv = const_cast<float*>(inv_norm);
  } else {
    rc = srgd_pixel_inv_norm(x, inv, M, C, stream);
    if (rc) return rc;
  }

  ProfScope prof(SRGD_PK_LINEAR_ATTN, 2.0 * (double)M * ((double)C * 384 + 128.0 * 128 * 2 + 128.0 * C),
                 (double)M * C * 2.0 * 2.0, st);
  if (C == 128 && getenv("SRGD_LA_SERIAL") == nullptr)     // test knob: force the single-pipeline kernels
    return launch_la_block_pp(x, qkv_w, out_w, out_b, out_g, out, B, N, inv, part, bd, splits, st);
  LaCtxParams ap;
  memset(&ap, 0, sizeof(ap));
  rc = make_tmap_2d_bf16(&ap.x_map, x, C, M, (uint64_t)C * 2, 64, 128, "linear_attention_block(x)");
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&ap.w_map, qkv_w, C, 384, (uint64_t)C * 2, 64, 128, "linear_attention_block(qkv_w)");
  if (rc) return rc;
  ap.inv = inv; ap.part = part; ap.C = C; ap.splits = splits; ap.tiles_per_sample = tiles;
  static uint64_t configured = 0;
  if (first_launch_on_device(configured)) {
    SRGD_CUDA_OK(cudaFuncSetAttribute(la_ctx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LaCtxSmem::kTotal));
  }
  SRGD_CUDA_OK(launch_k(la_ctx_kernel, dim3(splits, B), dim3(320), LaCtxSmem::kTotal, st, ap));
  SRGD_LAUNCH_OK("la_ctx_kernel");
  rc = launch_la_merge(part, reinterpret_cast<const bf16*>(out_w), bd, B, splits, C, st);
  if (rc) return rc;
  SRGD_LAUNCH_OK("la_merge_mb_kernel");
  count_launch(2);

  LaOutParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.x_map = ap.x_map;
  bp.w_map = ap.w_map;
  rc = make_tmap_2d_bf16(&bp.mb_map, bd, 128, (uint64_t)B * C, 256, 64, C, "linear_attention_block(Mb)");
  if (rc) return rc;
  bp.inv = inv; bp.bias = out_b; bp.g = out_g;
  bp.x = reinterpret_cast<const bf16*>(x);
  bp.out = reinterpret_cast<bf16*>(out);
  bp.chunks = splits; bp.tiles_per_sample = tiles;
  if (C == 128) return launch_la_out<128>(bp, splits, B, st);
  return launch_la_out<256>(bp, splits, B, st);
}
