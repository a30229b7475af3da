// Ping-pong variants of the fused LinearAttention kernels for C = 128 (the 256^2 and 128^2 stages: 86 % of
// all LinearAttention pixels of the shipped U-Net).  Same math as linattn_fused.cu (model.py:307-324, 703);
// what changes is the schedule: the serial per-tile chain  MMA -> softmax warps -> MMA -> ...  of the first
// version left the tensor pipe idle while the epilogue ran and vice versa (ncu: 23 % / 12 % tensor-active,
// 6.9k / 13.4k clk per 128-pixel tile).  Here every CTA runs TWO independent tile pipelines ("groups") with
// their own TMEM columns, shared-memory operand tiles and 8 epilogue warps each; the single MMA-issuing
// thread is a polling scheduler that issues whichever group's next MMA batch has its inputs ready, so one
// group's GEMMs overlap the other group's softmax / normalisation work.  Weights stay resident in shared
// memory (C = 128: 64 KB / 96 KB), only x tiles stream through the TMA ring.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace srgd {

constexpr float kPpLog2e = 1.4426950408889634f;
constexpr float kPpLn2 = 0.6931471805599453f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void pp_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// spin guard for the polling scheduler: a protocol bug must trap, not hang the GPU
struct SpinGuard {
  long long t0;
  uint32_t spins;
  __device__ __forceinline__ void progress() { spins = 0; }
  __device__ __forceinline__ void idle(const char* what) {
    if ((++spins & 0xfffff) == 0) {
      if (spins == 0x100000u) t0 = clock64();
      else if (clock64() - t0 > 4000000000LL) {
        printf("srgd_b200: %s scheduler stalled (block %d,%d)\n", what, (int)blockIdx.x, (int)blockIdx.y);
        __trap();
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// kernel A: context partials, 64-pixel sub-tiles, two groups
// ---------------------------------------------------------------------------------------------
struct alignas(64) LaCtxPpParams {
  CUtensorMap x_map;                    // bf16 [B*N][128], box {64, 64}
  CUtensorMap w_map;                    // bf16 [384][128], box {64, 128}
  const float* inv;
  float* part;                          // [B][2*splits][34][128]
  int32_t splits, tiles_per_sample;
};
struct LaCtxPpSmem {
  static constexpr int kWOffset = 0;                          // Wk kb0 | Wk kb1 | Wv kb0 | Wv kb1 (16 KB each)
  static constexpr int kXOffset = 65536;                      // 4 slots x (kb0 8 KB | kb1 8 KB)
  static constexpr int kPOffset = kXOffset + 4 * 16384;       // P_0 | P_1   [128 ch][64 px]
  static constexpr int kVtOffset = kPOffset + 2 * 16384;      // Vt_0 | Vt_1 [128 ch][64 px]
  static constexpr int kInvOffset = kVtOffset + 2 * 16384;    // float [2 groups][2 parity][2 (inv, inv*log2e)][64]
  static constexpr int kBarOffset = kInvOffset + 2048;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

__global__ void __launch_bounds__(576, 1) la_ctx_pp_kernel(const __grid_constant__ LaCtxPpParams p) {
  using L = LaCtxPpSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* w_full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* x_full = w_full + 1;
  uint64_t* x_empty = x_full + 4;
  uint64_t* kv_full = x_empty + 4;
  uint64_t* pv_ready = kv_full + 2;
  uint64_t* ctx_full = pv_ready + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(ctx_full + 2);
  const uint32_t smem_base = ptx::smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, sp = blockIdx.x;
  const int per = (p.tiles_per_sample + p.splits - 1) / p.splits;
  const int t0 = min(sp * per, p.tiles_per_sample), t1 = min(t0 + per, p.tiles_per_sample);
  const int ntiles = t1 - t0;                               // 128-pixel tiles; each group takes one half of each
  const int64_t row0 = ((int64_t)b * p.tiles_per_sample + t0) * 128;

  if (warp == 1 && lane == 0) {
    ptx::mbar_init(w_full, 1);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&x_full[s], 1);
      ptx::mbar_init(&x_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(&kv_full[g], 1);
      ptx::mbar_init(&pv_ready[g], 8);
      ptx::mbar_init(&ctx_full[g], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&p.x_map);
      ptx::prefetch_tmap(&p.w_map);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr_smem, 512);   // group g: K^T [256g,+64) | V^T [256g+64,+64) | ctx tile [256g+128,+128)
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  const int nsub = 2 * ntiles;

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    if (ntiles > 0) {
      ptx::mbar_arrive_expect_tx(w_full, 65536);
      ptx::tma_load_2d(smem + L::kWOffset, &p.w_map, w_full, 0, 128);
      ptx::tma_load_2d(smem + L::kWOffset + 16384, &p.w_map, w_full, 64, 128);
      ptx::tma_load_2d(smem + L::kWOffset + 32768, &p.w_map, w_full, 0, 256);
      ptx::tma_load_2d(smem + L::kWOffset + 49152, &p.w_map, w_full, 64, 256);
    }
    for (int s = 0; s < nsub; ++s) {
      const int slot = s & 3;
      ptx::mbar_wait(&x_empty[slot], ((s >> 2) & 1) ^ 1);
      uint8_t* dst = smem + L::kXOffset + slot * 16384;
      ptx::mbar_arrive_expect_tx(&x_full[slot], 16384);
      ptx::tma_load_2d(dst, &p.x_map, &x_full[slot], 0, (int)(row0 + s * 64));
      ptx::tma_load_2d(dst + 8192, &p.x_map, &x_full[slot], 64, (int)(row0 + s * 64));
    }
  } else if (warp == 1 && lane == 0) {
    // ================================ MMA issuer: polling scheduler ================================
    constexpr uint32_t idesc_kv = ptx::make_idesc_bf16_f32(128, 64);
    constexpr uint32_t idesc_ctx = ptx::make_idesc_bf16_f32(128, 128);
    if (ntiles > 0) ptx::mbar_wait(w_full, 0);
    int it[2] = {0, 0};
    int st[2] = {0, 0};
    SpinGuard guard{0, 0};
    while (it[0] < ntiles || it[1] < ntiles) {
      bool issued = false;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (it[g] >= ntiles) continue;
        const uint32_t tb = tmem_base + g * 256;
        if (st[g] == 0) {
          const int s = 2 * it[g] + g, slot = s & 3;
          if (!ptx::mbar_test_wait(&x_full[slot], (s >> 2) & 1)) continue;
          ptx::tc_fence_after();
          const uint32_t xb = smem_base + L::kXOffset + slot * 16384;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t wk = ptx::make_sw128_kmajor_desc(smem_base + L::kWOffset + kb * 16384);
            const uint64_t wv = ptx::make_sw128_kmajor_desc(smem_base + L::kWOffset + 32768 + kb * 16384);
            const uint64_t xd = ptx::make_sw128_kmajor_desc(xb + kb * 8192);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
              ptx::umma_bf16_ss(tb, wk + 2 * k, xd + 2 * k, idesc_kv, accum);
              ptx::umma_bf16_ss(tb + 64, wv + 2 * k, xd + 2 * k, idesc_kv, accum);
            }
          }
          ptx::umma_commit(&x_empty[slot]);
          ptx::umma_commit(&kv_full[g]);
          st[g] = 1;
          issued = true;
        } else {
          if (!ptx::mbar_test_wait(&pv_ready[g], it[g] & 1)) continue;
          ptx::tc_fence_after();
          const uint64_t pd = ptx::make_sw128_kmajor_desc(smem_base + L::kPOffset + g * 16384);
          const uint64_t vd = ptx::make_sw128_kmajor_desc(smem_base + L::kVtOffset + g * 16384);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)                   // the context accumulates in TMEM across the tiles of this CTA
            ptx::umma_bf16_ss(tb + 128, pd + 2 * ks, vd + 2 * ks, idesc_ctx, (it[g] | ks) != 0 ? 1u : 0u);
          ptx::umma_commit(&ctx_full[g]);
          st[g] = 0;
          ++it[g];
          issued = true;
        }
      }
      if (issued) guard.progress();
      else guard.idle("la_ctx_pp");
    }
  } else if (warp >= 2) {
    // ================================== softmax / operand warps ==================================
    const int e = warp - 2;
    const int g = e >> 3;                                  // pipeline this warp belongs to
    const bool is_k = (e & 7) < 4;
    const int q = warp & 3;                                // TMEM lane quarter (= head)
    const int row = q * 32 + lane;                         // channel (h,d) resp. (h,e)
    const int gt = threadIdx.x - 64 - g * 256;             // 0..255 within the group
    const uint32_t tb = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    const uint32_t tile_addr = smem_base + (is_k ? L::kPOffset : L::kVtOffset) + g * 16384;
    const uint32_t inv_addr0 = smem_base + L::kInvOffset + g * 1024;
    // K warps: the un-normalised context of this thread's channel accumulates in TMEM across tiles (the MMA
    // adds every tile's P V), all relative to the exponent reference m_ref (log2 units); z_run likewise.
    float m_ref = -INFINITY, pending = -INFINITY, z_run = 0.f;

    // 1/||x|| of the group's next 64 pixels is fetched one iteration ahead (a dependent DRAM load at the top
    // of every iteration would stall all 256 threads of the group at the barrier below)
    float inv_next = (gt < 64 && ntiles > 0) ? p.inv[row0 + g * 64 + gt] : 0.f;
    for (int it = 0; it < ntiles; ++it) {
      const uint32_t par = it & 1;
      const uint32_t inv_addr = inv_addr0 + par * 512;     // [inv[64] | inv*log2e[64]]
      if (gt < 64) {
        const float v = inv_next;
        if (it + 1 < ntiles) inv_next = p.inv[row0 + (2 * (it + 1) + g) * 64 + gt];
        ptx::sts_f32(inv_addr + gt * 4, v);
        ptx::sts_f32(inv_addr + 256 + gt * 4, v * kPpLog2e);
      }
      pp_bar_sync(1 + g, 256);
      // kv_full(it) is a tcgen05.commit issued after the previous tile's P V MMA: once it fires, P / Vt and the
      // context accumulator of tile it-1 are no longer in use by the tensor core.
      ptx::mbar_wait(&kv_full[g], par);
      ptx::tc_fence_after();
      if (is_k) {
        float shift = pending;                             // == m_ref unless a re-reference is due
        if (it == 0) {
          // first tile: no reference yet -> one extra pass over the accumulator for the exact column max
          float ma = -INFINITY, mb2 = -INFINITY;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tb + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 il = ptx::lds_f4(inv_addr + 256 + (c * 32 + j) * 4);
              ma = fmaxf(ma, fmaxf(__uint_as_float(v[j]) * il.x, __uint_as_float(v[j + 1]) * il.y));
              mb2 = fmaxf(mb2, fmaxf(__uint_as_float(v[j + 2]) * il.z, __uint_as_float(v[j + 3]) * il.w));
            }
          }
          shift = fmaxf(ma, mb2);
        }
        // Single pass: P = 2^(k - shift) with the RUNNING reference as shift (P may exceed 1; bf16 has fp32's
        // range) while the tile's true max is tracked alongside.  If some column jumped far above the reference
        // the warp redoes the tile with the exact max -- never taken with normalised inputs, kept for safety.
        float m_t, z_t;
        bool redo = false;
        do {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          float z4[4] = {0.f, 0.f, 0.f, 0.f};
          uint32_t va[16], vb[16];
          ptx::tmem_ld_32x16(tb, va);
#pragma unroll
          for (int sc = 0; sc < 4; ++sc) {                   // 16 pixels at a time; the next load overlaps the math
            uint32_t* v = (sc & 1) ? vb : va;
            ptx::tmem_ld_wait();
            if (sc < 3) ptx::tmem_ld_32x16(tb + (sc + 1) * 16, (sc & 1) ? va : vb);
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 il = ptx::lds_f4(inv_addr + 256 + (sc * 16 + j) * 4);
              const float t0 = __uint_as_float(v[j]) * il.x, t1 = __uint_as_float(v[j + 1]) * il.y;
              const float t2 = __uint_as_float(v[j + 2]) * il.z, t3 = __uint_as_float(v[j + 3]) * il.w;
              m4[0] = fmaxf(m4[0], t0);
              m4[1] = fmaxf(m4[1], t1);
              m4[2] = fmaxf(m4[2], t2);
              m4[3] = fmaxf(m4[3], t3);
              const float p0 = fast_exp2(t0 - shift), p1 = fast_exp2(t1 - shift);
              const float p2 = fast_exp2(t2 - shift), p3 = fast_exp2(t3 - shift);
              z4[0] += p0; z4[1] += p1; z4[2] += p2; z4[3] += p3;
              w[j >> 1] = pack_bf16(p0, p1);
              w[(j >> 1) + 1] = pack_bf16(p2, p3);
            }
            ptx::sts_v4(tile_addr + ptx::sw128_offset(row, sc * 2), w[0], w[1], w[2], w[3]);
            ptx::sts_v4(tile_addr + ptx::sw128_offset(row, sc * 2 + 1), w[4], w[5], w[6], w[7]);
          }
          m_t = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
          z_t = (z4[0] + z4[1]) + (z4[2] + z4[3]);
          const bool far = (m_t - shift) > 40.f;
          redo = !redo && __any_sync(0xffffffffu, far);
          if (redo && far) shift = m_t;
        } while (redo);
        // the reference moved (pending re-reference or redo): rescale this channel's accumulated context in TMEM
        const bool moved = (it > 0) && (shift != m_ref);
        if (__any_sync(0xffffffffu, moved)) {
          const float f = moved ? fast_exp2(m_ref - shift) : 1.0f;
          uint32_t v[32];
          ptx::tmem_ld_32x32(tb + 128 + q * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
          ptx::tmem_st_32x32(tb + 128 + q * 32, v);
          ptx::tmem_st_wait();
          z_run *= f;
        }
        m_ref = shift;
        z_run += z_t;
        pending = (m_t > m_ref + 20.f) ? m_t : m_ref;      // keep later tiles well inside fp32 / bf16 range
      } else {
        uint32_t va[16], vb[16];
        ptx::tmem_ld_32x16(tb + 64, va);
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) {
          uint32_t* v = (sc & 1) ? vb : va;
          ptx::tmem_ld_wait();
          if (sc < 3) ptx::tmem_ld_32x16(tb + 64 + (sc + 1) * 16, (sc & 1) ? va : vb);
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 iv = ptx::lds_f4(inv_addr + (sc * 16 + j) * 4);
            w[j >> 1] = pack_bf16(__uint_as_float(v[j]) * iv.x, __uint_as_float(v[j + 1]) * iv.y);
            w[(j >> 1) + 1] = pack_bf16(__uint_as_float(v[j + 2]) * iv.z, __uint_as_float(v[j + 3]) * iv.w);
          }
          ptx::sts_v4(tile_addr + ptx::sw128_offset(row, sc * 2), w[0], w[1], w[2], w[3]);
          ptx::sts_v4(tile_addr + ptx::sw128_offset(row, sc * 2 + 1), w[4], w[5], w[6], w[7]);
        }
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&pv_ready[g]);
    }
    if (is_k) {
      float ctx[32];
      if (ntiles > 0) {
        ptx::mbar_wait(&ctx_full[g], (ntiles - 1) & 1);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld_32x32(tb + 128 + q * 32, v);           // diagonal 32x32 block of head q
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) ctx[i] = __uint_as_float(v[i]);
        ptx::tc_fence_before();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) ctx[i] = 0.f;
      }
      // record layout [34][128 channels]: coalesced here and in la_merge_mb_kernel
      float* dst = p.part + ((int64_t)b * (2 * p.splits) + 2 * sp + g) * (34 * 128) + row;
      dst[0] = m_ref * kPpLn2;                             // the merge kernel works in natural-log units
      dst[128] = z_run;
#pragma unroll
      for (int i = 0; i < 32; ++i) dst[(2 + i) * 128] = ctx[i];
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
// kernel B: q softmax, y = softmax(q) Mb^T, RMSNorm, residual; 128-pixel tiles, two groups
// ---------------------------------------------------------------------------------------------
struct alignas(64) LaOutPpParams {
  CUtensorMap x_map;                    // bf16 [B*N][128], box {64, 128}
  CUtensorMap w_map;                    // bf16 [384][128], box {64, 128} (rows 0..127 = W_q)
  CUtensorMap mb_map;                   // bf16 [B*128][128] (la_merge_mb_kernel), box {64, 128}
  const float* inv;
  const float* bias;
  const float* g;
  const bf16* x;
  bf16* out;
  int32_t chunks, tiles_per_sample;
};
#ifndef SRGD_PP_STAGE
#define SRGD_PP_STAGE 1
#endif
constexpr bool kPpStage = SRGD_PP_STAGE != 0;            // output rows through a shared-memory transpose (A/B knob)
struct LaOutPpSmem {
  static constexpr int kWqOffset = 0;                         // 2 k-blocks x 16 KB
  static constexpr int kMbOffset = 32768;                     // 2 k-blocks x 16 KB
  static constexpr int kXOffset = 65536;                      // ring of 4 k-block slots x 16 KB
  static constexpr int kQsOffset = kXOffset + 4 * 16384;      // per group: softmax(q) as bf16 A operand, 32 KB
  static constexpr int kBiasOffset = kQsOffset + 2 * 32768;   // bias[128] | g[128]
  static constexpr int kSsqOffset = kBiasOffset + 1024;       // float [2 groups][2 parity][2 halves][128]
  static constexpr int kBarOffset = kSsqOffset + 4096;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

__global__ void __launch_bounds__(576, 1) la_out_pp_kernel(const __grid_constant__ LaOutPpParams p) {
  using L = LaOutPpSmem;
  constexpr int C = 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* const_full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* x_full = const_full + 1;
  uint64_t* x_empty = x_full + 4;
  uint64_t* q_full = x_empty + 4;
  uint64_t* qs_ready = q_full + 2;
  uint64_t* y_full = qs_ready + 2;
  uint64_t* y_done = y_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(y_done + 2);
  const uint32_t smem_base = ptx::smem_u32(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, ck = blockIdx.x;
  const int per = (p.tiles_per_sample + p.chunks - 1) / p.chunks;
  const int t0 = min(ck * per, p.tiles_per_sample), t1 = min(t0 + per, p.tiles_per_sample);
  const int ntiles = t1 - t0;
  const int64_t row0 = ((int64_t)b * p.tiles_per_sample + t0) * 128;

  if (warp == 1 && lane == 0) {
    ptx::mbar_init(const_full, 1);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&x_full[s], 1);
      ptx::mbar_init(&x_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(&q_full[g], 1);
      ptx::mbar_init(&qs_ready[g], 8);
      ptx::mbar_init(&y_full[g], 1);
      ptx::mbar_init(&y_done[g], 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) {
      ptx::prefetch_tmap(&p.x_map);
      ptx::prefetch_tmap(&p.w_map);
      ptx::prefetch_tmap(&p.mb_map);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr_smem, 512);                  // group g: Q [256g,+128) | Y [256g+128,+128)
    ptx::tmem_relinquish();
  }
  if (warp >= 2) {
    const int et = threadIdx.x - 64;
    if (et < C) {
      ptx::sts_f32(smem_base + L::kBiasOffset + et * 4, p.bias[et]);
      ptx::sts_f32(smem_base + L::kBiasOffset + 512 + et * 4, p.g[et]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    if (ntiles > 0) {
      ptx::mbar_arrive_expect_tx(const_full, 2 * 32768);
      ptx::tma_load_2d(smem + L::kWqOffset, &p.w_map, const_full, 0, 0);
      ptx::tma_load_2d(smem + L::kWqOffset + 16384, &p.w_map, const_full, 64, 0);
      ptx::tma_load_2d(smem + L::kMbOffset, &p.mb_map, const_full, 0, b * 128);
      ptx::tma_load_2d(smem + L::kMbOffset + 16384, &p.mb_map, const_full, 64, b * 128);
    }
    for (int seq = 0; seq < 2 * ntiles; ++seq) {           // k-block `seq & 1` of tile `seq >> 1`
      const int slot = seq & 3;
      ptx::mbar_wait(&x_empty[slot], ((seq >> 2) & 1) ^ 1);
      ptx::mbar_arrive_expect_tx(&x_full[slot], 16384);
      ptx::tma_load_2d(smem + L::kXOffset + slot * 16384, &p.x_map, &x_full[slot], (seq & 1) * 64,
                       (int)(row0 + (seq >> 1) * 128));
    }
  } else if (warp == 1 && lane == 0) {
    // ================================ MMA issuer: polling scheduler ================================
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128);
    if (ntiles > 0) ptx::mbar_wait(const_full, 0);
    const int n_g[2] = {(ntiles + 1) >> 1, ntiles >> 1};
    int it[2] = {0, 0};
    int st[2] = {0, 0};
    SpinGuard guard{0, 0};
    while (it[0] < n_g[0] || it[1] < n_g[1]) {
      bool issued = false;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if (it[g] >= n_g[g]) continue;
        const uint32_t tb = tmem_base + g * 256;
        const uint32_t qs = smem_base + L::kQsOffset + g * 32768;
        if (st[g] == 0) {
          // Q[px][(h,d)] = x Wq^T.  Tile T = 2*it+g sits in ring slots {2T & 3, (2T+1) & 3}; group g always
          // gets the same slot pair, refilled only after its own previous MMA1 retired.
          const int tile = 2 * it[g] + g;
          const int s0 = (2 * tile) & 3, s1 = s0 + 1;
          const uint32_t xpar = ((2 * tile) >> 2) & 1;
          if (!ptx::mbar_test_wait(&x_full[s0], xpar) || !ptx::mbar_test_wait(&x_full[s1], xpar)) continue;
          ptx::tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const int slot = s0 + kb;
            const uint64_t xd = ptx::make_sw128_kmajor_desc(smem_base + L::kXOffset + slot * 16384);
            const uint64_t wq = ptx::make_sw128_kmajor_desc(smem_base + L::kWqOffset + kb * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss(tb, xd + 2 * k, wq + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            ptx::umma_commit(&x_empty[slot]);
          }
          ptx::umma_commit(&q_full[g]);
          st[g] = 1;
          issued = true;
        } else {
          // Y[px][c] = softmax(q)[px][(h,d)] * Mb[c][(h,d)]^T; the epilogue must have drained the previous Y
          if (!ptx::mbar_test_wait(&qs_ready[g], it[g] & 1)) continue;
          if (it[g] > 0 && !ptx::mbar_test_wait(&y_done[g], (it[g] - 1) & 1)) continue;
          ptx::tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t ad = ptx::make_sw128_kmajor_desc(qs + (ks >> 2) * 16384) + 2 * (ks & 3);
            const uint64_t bd = ptx::make_sw128_kmajor_desc(smem_base + L::kMbOffset + (ks >> 2) * 16384) + 2 * (ks & 3);
            ptx::umma_bf16_ss(tb + 128, ad, bd, idesc, ks != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&y_full[g]);
          st[g] = 0;
          ++it[g];
          issued = true;
        }
      }
      if (issued) guard.progress();
      else guard.idle("la_out_pp");
    }
  } else if (warp >= 2) {
    // ======================================= epilogue =======================================
    const int e = warp - 2;
    const int g = e >> 3;
    const int half = (e >> 2) & 1;                         // column half: heads {2*half, 2*half+1}, channels [64*half,+64)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t tb = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    const uint32_t qs_addr = smem_base + L::kQsOffset + g * 32768 + half * 16384;   // k-block `half` of the A tile
    const uint32_t bias_addr = smem_base + L::kBiasOffset + half * 256;
    const uint32_t gain_addr = bias_addr + 512;
    const int n_g = (ntiles + 1 - g) >> 1;
    const float sqrt_c = 11.313708498984761f;              // sqrt(128)

    for (int it = 0; it < n_g; ++it) {
      const uint32_t par = it & 1;
      const int64_t px = row0 + (int64_t)(2 * it + g) * 128 + row;
      // requested before the wait for this tile's Q: the load latency hides behind it (a value prefetched one
      // tile ahead lived in a register across the whole body and was spilled -- ncu: 9 % of the samples at its reload)
      const float my_inv = p.inv[px];
      // ---- q: softmax over the 32 channels of each head (model.py:315); 1/||x|| folded into the exponent ----
      ptx::mbar_wait(&q_full[g], par);
      ptx::tc_fence_after();
      const float my_invl = my_inv * kPpLog2e;
#pragma unroll 1
      for (int i = 0; i < 2; ++i) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tb + (half * 2 + i) * 32, v);
        ptx::tmem_ld_wait();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        const float ml = mx * my_invl;                     // my_invl > 0: the scaled max is the max of the scaled values
        float f[32];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          f[j] = fast_exp2(fmaf(__uint_as_float(v[j]), my_invl, -ml));
          f[j + 1] = fast_exp2(fmaf(__uint_as_float(v[j + 1]), my_invl, -ml));
          s0 += f[j];
          s1 += f[j + 1];
        }
        const float rs = 1.0f / (s0 + s1);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          ptx::sts_v4(qs_addr + ptx::sw128_offset(row, i * 4 + jj), pack_bf16(f[8 * jj] * rs, f[8 * jj + 1] * rs),
                      pack_bf16(f[8 * jj + 2] * rs, f[8 * jj + 3] * rs), pack_bf16(f[8 * jj + 4] * rs, f[8 * jj + 5] * rs),
                      pack_bf16(f[8 * jj + 6] * rs, f[8 * jj + 7] * rs));
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&qs_ready[g]);

      // residual rows (L2-hot: the TMA fetched the same lines), requested while the Y MMA runs.  Loaded COALESCED --
      // eight lanes cover one row's 128 bytes, a load instruction touches 4 full lines instead of 32 half sectors (the
      // L1 data pipe was 70 % busy) -- and transposed to one row per thread through the warp's staging rows below.
      // (Taking them from the x ring slot instead, slot handed back by the epilogue, was tried in round 2:
      // nondeterministic results at B=16, 128x128, cause not found; profiles/r02_experiments.txt.)
      uint4 xr[8];
      {
        const int64_t px0 = row0 + (int64_t)(2 * it + g) * 128 + q * 32;
        const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          xr[k] = *reinterpret_cast<const uint4*>(p.x + (px0 + k * 4 + sub) * C + half * 64 + chunk * 8);
      }

      // ---- y: + bias, RMSNorm over all 128 channels of the pixel (model.py:207), * g, + x ----
      ptx::mbar_wait(&y_full[g], par);
      ptx::tc_fence_after();
      {
        // the Y MMA has consumed the softmax(q) operand: this warp's 4 KB of it are free -> transpose the residual
        const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          ptx::sts_v4(qs_addr + ptx::sw128_offset(q * 32 + k * 4 + sub, chunk), xr[k].x, xr[k].y, xr[k].z, xr[k].w);
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; ++t) xr[t] = ptx::lds_v4(qs_addr + ptx::sw128_offset(row, t));
        __syncwarp();                                      // the rows are rewritten by the output staging below
      }
      float ssq = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tb + 128 + half * 64 + cc * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bv = ptx::lds_f4(bias_addr + (cc * 32 + j) * 4);
          const float y0 = __uint_as_float(v[j]) + bv.x, y1 = __uint_as_float(v[j + 1]) + bv.y;
          const float y2 = __uint_as_float(v[j + 2]) + bv.z, y3 = __uint_as_float(v[j + 3]) + bv.w;
          ssq += (y0 * y0 + y1 * y1) + (y2 * y2 + y3 * y3);
        }
      }
      const uint32_t ssq_addr = smem_base + L::kSsqOffset + (g * 2 + par) * 1024;
      ptx::sts_f32(ssq_addr + (half * 128 + row) * 4, ssq);
      pp_bar_sync(1 + g, 256);
      const float tot = ssq + ptx::lds_f32(ssq_addr + ((half ^ 1) * 128 + row) * 4);
      const float scale = sqrt_c / fmaxf(sqrtf(tot), 1e-12f);
      // Output rows go through this warp's own 4 KB of the softmax(q) operand tile (free once the Y MMA has
      // completed): written row-wise, read back so that eight lanes cover one row's 128 bytes -- a store instruction
      // then touches 4 full lines instead of 32 half sectors.
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {                     // unrolled (xr[] stays in registers); 16 columns at a time
        uint32_t v[16];                                    // keeps the live set under the 96-register cap of 576 threads
        ptx::tmem_ld_32x16(tb + 128 + half * 64 + cc * 16, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          float r[8], o[8];
          unpack8(xr[cc * 2 + jj], r);
          const float4 b0 = ptx::lds_f4(bias_addr + (cc * 16 + jj * 8) * 4);
          const float4 b1 = ptx::lds_f4(bias_addr + (cc * 16 + jj * 8 + 4) * 4);
          const float4 g0 = ptx::lds_f4(gain_addr + (cc * 16 + jj * 8) * 4);
          const float4 g1 = ptx::lds_f4(gain_addr + (cc * 16 + jj * 8 + 4) * 4);
          o[0] = fmaf(__uint_as_float(v[jj * 8 + 0]) + b0.x, scale * g0.x, r[0]);
          o[1] = fmaf(__uint_as_float(v[jj * 8 + 1]) + b0.y, scale * g0.y, r[1]);
          o[2] = fmaf(__uint_as_float(v[jj * 8 + 2]) + b0.z, scale * g0.z, r[2]);
          o[3] = fmaf(__uint_as_float(v[jj * 8 + 3]) + b0.w, scale * g0.w, r[3]);
          o[4] = fmaf(__uint_as_float(v[jj * 8 + 4]) + b1.x, scale * g1.x, r[4]);
          o[5] = fmaf(__uint_as_float(v[jj * 8 + 5]) + b1.y, scale * g1.y, r[5]);
          o[6] = fmaf(__uint_as_float(v[jj * 8 + 6]) + b1.z, scale * g1.z, r[6]);
          o[7] = fmaf(__uint_as_float(v[jj * 8 + 7]) + b1.w, scale * g1.w, r[7]);
          const uint4 pk = pack8(o);
          if (kPpStage) ptx::sts_v4(qs_addr + ptx::sw128_offset(row, cc * 2 + jj), pk.x, pk.y, pk.z, pk.w);
          else st_stream(p.out + px * C + half * 64 + cc * 16 + jj * 8, pk);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&y_done[g]);         // the accumulator is drained; the copy-out needs no TMEM
      if (kPpStage) {
        const int64_t px0 = row0 + (int64_t)(2 * it + g) * 128 + q * 32;
        const int sub = lane >> 3, chunk = lane & 7;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int rl = k * 4 + sub;                      // row of this warp's 32
          const uint4 pk = ptx::lds_v4(qs_addr + ptx::sw128_offset(q * 32 + rl, chunk));
          st_stream(p.out + (px0 + rl) * C + half * 64 + chunk * 8, pk);
        }
        __syncwarp();                                      // the tile is rewritten by the next softmax
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// merge kernel of linattn_fused.cu
int launch_la_merge(const float* part, const bf16* wout, bf16* mb, int B, int splits, int C, cudaStream_t st);

// Launches the C = 128 ping-pong pipeline: context partials -> merge -> output.  `inv` = 1/||x|| per pixel.
int launch_la_block_pp(const void* x, const void* qkv_w, const void* out_w, const float* out_b, const float* out_g,
                       void* out, int B, int N, const float* inv, float* part, bf16* bd, int splits, cudaStream_t st) {
  const int tiles = N / 128;
  const int64_t M = (int64_t)B * N;
  LaCtxPpParams ap;
  memset(&ap, 0, sizeof(ap));
  int rc = make_tmap_2d_bf16(&ap.x_map, x, 128, M, 256, 64, 64, "linear_attention_block(x/64)");
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&ap.w_map, qkv_w, 128, 384, 256, 64, 128, "linear_attention_block(qkv_w)");
  if (rc) return rc;
  ap.inv = inv; ap.part = part; ap.splits = splits; ap.tiles_per_sample = tiles;
  static uint64_t configured = 0;
  if (first_launch_on_device(configured)) {
    SRGD_CUDA_OK(cudaFuncSetAttribute(la_ctx_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LaCtxPpSmem::kTotal));
    SRGD_CUDA_OK(cudaFuncSetAttribute(la_out_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LaOutPpSmem::kTotal));
  }
  SRGD_CUDA_OK(launch_k(la_ctx_pp_kernel, dim3(splits, B), dim3(576), LaCtxPpSmem::kTotal, st, ap));
  SRGD_LAUNCH_OK("la_ctx_pp_kernel");
  rc = launch_la_merge(part, reinterpret_cast<const bf16*>(out_w), bd, B, 2 * splits, 128, st);
  if (rc) return rc;
  SRGD_LAUNCH_OK("la_merge_mb_kernel");

  LaOutPpParams bp;
  memset(&bp, 0, sizeof(bp));
  rc = make_tmap_2d_bf16(&bp.x_map, x, 128, M, 256, 64, 128, "linear_attention_block(x/128)");
  if (rc) return rc;
  bp.w_map = ap.w_map;
  rc = make_tmap_2d_bf16(&bp.mb_map, bd, 128, (uint64_t)B * 128, 256, 64, 128, "linear_attention_block(Mb)");
  if (rc) return rc;
  bp.inv = inv; bp.bias = out_b; bp.g = out_g;
  bp.x = reinterpret_cast<const bf16*>(x);
  bp.out = reinterpret_cast<bf16*>(out);
  bp.chunks = splits; bp.tiles_per_sample = tiles;
  SRGD_CUDA_OK(launch_k(la_out_pp_kernel, dim3(splits, B), dim3(576), LaOutPpSmem::kTotal, st, bp));
  SRGD_LAUNCH_OK("la_out_pp_kernel");
  count_launch(3);
  return SRGD_OK;
}

}  // namespace srgd
