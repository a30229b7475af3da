// Full attention core on tcgen05 tensor cores (reference: Attention.forward model.py:344-355 calling
// denoising_diffusion_pytorch's Attend, flash=False: softmax(q k^T * 32^-1/2) v).
//
// One CTA per (sample, head, 128-query tile); keys/values are consumed in blocks of 128:
//     S[128 q][128 keys]  = Q K_j^T          tcgen05.mma, A = Q tile, B = K block (both K-major in smem via TMA)
//     P = exp2(c (S - m)) ; online softmax   4 warps, thread = query row (TMEM lane), fp32
//     PV[128 q][32 d]     = P V_j            tcgen05.mma, A = P (bf16, written to smem by the softmax warps),
//                                            B = V_j^T (transposed into smem by the same warps)
//     O = O * corr + PV                      registers (32 fp32 per thread)
// dim_head = 32 is half a 128-byte swizzle row: the TMA box is 64 channels wide over a tensor whose inner
// extent is 32, so the upper half of every row is zero-filled and the known-good SWIZZLE_128B K-major layout
// is used unchanged; only the two K=16 steps that carry data are issued.
// 90 KB of shared memory and 256 TMEM columns per CTA -> two CTAs per SM overlap each other's softmax
// (MUFU-bound: N^2 exponentials per head) and MMA phases.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.h"

namespace srgd {

constexpr int kFtStages = 2;
struct alignas(64) FaTcParams {
  CUtensorMap qk_map;                   // bf16 view [B*N rows][12 segments][32 ch], box {64, 1, 128}
  const bf16* qkv;
  bf16* out;
  int32_t N, heads;
  float scale_log2e;                    // 32^-1/2 * log2(e)
};
struct FaTcSmem {
  static constexpr int kQOffset = 0;                          // [128 q][64] (upper 32 channels zero)
  static constexpr int kKOffset = 16384;                      // stages of [128 keys][64]
  static constexpr int kPOffset = kKOffset + kFtStages * 16384;   // 2 k-blocks of [128 q][64 keys]
  static constexpr int kVtOffset = kPOffset + 32768;          // 2 k-blocks of [32 d][64 keys]
  static constexpr int kBarOffset = kVtOffset + 8192;
  static constexpr int kTotal = kBarOffset + 128 + 1024;
};

__global__ void __launch_bounds__(192, 2) fa_tc_kernel(const __grid_constant__ FaTcParams p) {
  using L = FaTcSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + kFtStages;
  uint64_t* s_full = k_empty + kFtStages;
  uint64_t* p_ready = s_full + 1;
  uint64_t* pv_full = p_ready + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y / p.heads, h = blockIdx.y % p.heads;
  const int q0 = blockIdx.x * 128;
  const int nblocks = p.N / 128;
  const int hid = p.heads * 32;

  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < kFtStages; ++s) {
      ptx::mbar_init(&k_full[s], 1);
      ptx::mbar_init(&k_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_ready, 4);
    ptx::mbar_init(pv_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    if (lane == 0) ptx::prefetch_tmap(&p.qk_map);
    __syncwarp();
    ptx::tmem_alloc(tmem_ptr_smem, 256);                  // S [0,128) | PV [128,160)
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0 && lane == 0) {
    // ===================================== TMA producer =====================================
    ptx::mbar_arrive_expect_tx(q_full, 16384);
    ptx::tma_load_3d(smem + L::kQOffset, &p.qk_map, q_full, 0, h, b * p.N + q0);
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < nblocks; ++j) {
      ptx::mbar_wait(&k_empty[stage], phase ^ 1);
      ptx::mbar_arrive_expect_tx(&k_full[stage], 16384);
      ptx::tma_load_3d(smem + L::kKOffset + stage * 16384, &p.qk_map, &k_full[stage], 0, p.heads + h,
                       b * p.N + j * 128);
      if (++stage == kFtStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // ====================================== MMA issuer ======================================
    constexpr uint32_t idesc_s = ptx::make_idesc_bf16_f32(128, 128);
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_f32(128, 32);
    const uint64_t qd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kQOffset));
    const uint32_t pb = ptx::smem_u32(smem + L::kPOffset), vb = ptx::smem_u32(smem + L::kVtOffset);
    int stage = 0;
    uint32_t phase = 0;
    ptx::mbar_wait(q_full, 0);
    for (int j = 0; j <= nblocks; ++j) {
      if (j > 0) {
        // PV(j-1): the softmax warps have read S(j-1) and published P / V^T
        ptx::mbar_wait(p_ready, (j - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t pd = ptx::make_sw128_kmajor_desc(pb + (ks >> 2) * 16384) + 2 * (ks & 3);
          const uint64_t vd = ptx::make_sw128_kmajor_desc(vb + (ks >> 2) * 4096) + 2 * (ks & 3);
          ptx::umma_bf16_ss(tmem_base + 128, pd, vd, idesc_pv, ks != 0 ? 1u : 0u);
        }
        ptx::umma_commit(pv_full);
      }
      if (j < nblocks) {
        ptx::mbar_wait(&k_full[stage], phase);
        ptx::tc_fence_after();
        const uint64_t kd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + L::kKOffset + stage * 16384));
#pragma unroll
        for (int k = 0; k < 2; ++k)                        // channels 32..63 of every row are zero: skip them
          ptx::umma_bf16_ss(tmem_base, qd + 2 * k, kd + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(&k_empty[stage]);
        ptx::umma_commit(s_full);
        if (++stage == kFtStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2) {
    // ===================================== softmax warps =====================================
    const int q = warp & 3;
    const int row = q * 32 + lane;                         // query row == key row of the V block this thread moves
    const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
    const float c = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f;
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
    const bf16* vbase = p.qkv + (int64_t)b * p.N * 3 * hid + 2 * hid + h * 32;
    uint8_t* p_smem = smem + L::kPOffset;
    uint8_t* vt_smem = smem + L::kVtOffset;

    for (int j = 0; j < nblocks; ++j) {
      const uint32_t par = j & 1;
      // V block row (key j*128 + row): 32 channels, fetched before waiting on the scores
      uint4 vr[4];
      const bf16* vrow = vbase + (int64_t)(j * 128 + row) * 3 * hid;
#pragma unroll
      for (int t = 0; t < 4; ++t) vr[t] = ld_stream(vrow + t * 8);
      if (j > 0) {
        // O += PV(j-1) (its MMA must have consumed P / V^T before they are overwritten below)
        ptx::mbar_wait(pv_full, (j - 1) & 1);
        ptx::tc_fence_after();
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + lane_bits + 128, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] += __uint_as_float(v[d]);
      }
      // transpose V into the K-major B operand: Vt[d][key]
      {
        const int kb = row >> 6, kk = row & 63;
        uint8_t* dst = vt_smem + kb * 4096 + (kk & 7) * 2;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t wv[4] = {vr[t].x, vr[t].y, vr[t].z, vr[t].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int d = t * 8 + i * 2;
            *reinterpret_cast<uint16_t*>(dst + ptx::sw128_offset(d, kk >> 3)) = (uint16_t)(wv[i] & 0xffffu);
            *reinterpret_cast<uint16_t*>(dst + ptx::sw128_offset(d + 1, kk >> 3)) = (uint16_t)(wv[i] >> 16);
          }
        }
      }
      ptx::mbar_wait(s_full, par);
      ptx::tc_fence_after();
      float mx = m_run;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + lane_bits + cc * 32, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float corr = exp2f((m_run - mx) * c);          // 0 on the first block
      const float mc = mx * c;
      float sum = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + lane_bits + cc * 32, v);
        ptx::tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = exp2f(__uint_as_float(v[i]) * c - mc);
          const float p1 = exp2f(__uint_as_float(v[i + 1]) * c - mc);
          sum += p0 + p1;
          w[i >> 1] = pack_bf16(p0, p1);
        }
        uint8_t* kbase = p_smem + (cc >> 1) * 16384;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          *reinterpret_cast<uint4*>(kbase + ptx::sw128_offset(row, (cc & 1) * 4 + jj)) =
              make_uint4(w[4 * jj], w[4 * jj + 1], w[4 * jj + 2], w[4 * jj + 3]);
      }
      l_run = l_run * corr + sum;
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] *= corr;
      m_run = mx;
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_ready);
    }
    ptx::mbar_wait(pv_full, (nblocks - 1) & 1);
    ptx::tc_fence_after();
    {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_base + lane_bits + 128, v);
      ptx::tmem_ld_wait();
      const float inv = 1.0f / l_run;
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] = (o[d] + __uint_as_float(v[d])) * inv;
    }
    bf16* op = p.out + ((int64_t)b * p.N + q0 + row) * hid + h * 32;
#pragma unroll
    for (int t = 0; t < 4; ++t) st_stream(op + t * 8, pack8(o + t * 8));
    ptx::tc_fence_before();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

int make_tmap_qk_heads(CUtensorMap* m, const void* qkv, int64_t rows, int heads);   // conv_igemm.cu

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_attention_tc_supported(int32_t N, int32_t heads) {
  return (heads >= 1 && heads <= 4 && N > 0 && N % 128 == 0) ? 1 : 0;
}

extern "C" int srgd_attention_tc(const void* qkv, void* out, int32_t B, int32_t N, int32_t heads,
                                 srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(qkv && out && B > 0, "attention_tc: null argument");
  SRGD_REQUIRE(srgd_attention_tc_supported(N, heads), "attention_tc: unsupported shape N=%d heads=%d (N %% 128 == 0)", N,
               heads);
  SRGD_REQUIRE((int64_t)B * heads <= 65535, "attention_tc: B*heads too large");
  SRGD_REQUIRE(((uintptr_t)qkv | (uintptr_t)out) % 16 == 0, "attention_tc: pointers must be 16-byte aligned");
  FaTcParams kp;
  memset(&kp, 0, sizeof(kp));
  rc = make_tmap_qk_heads(&kp.qk_map, qkv, (int64_t)B * N, heads);
  if (rc) return rc;
  kp.qkv = reinterpret_cast<const bf16*>(qkv);
  kp.out = reinterpret_cast<bf16*>(out);
  kp.N = N;
  kp.heads = heads;
  kp.scale_log2e = 0.17677669529663687f * 1.4426950408889634f;
  static uint64_t configured = 0;
  if (first_launch_on_device(configured)) {
    SRGD_CUDA_OK(cudaFuncSetAttribute(fa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FaTcSmem::kTotal));
  }
  cudaStream_t st = as_stream(stream);
  ProfScope prof(SRGD_PK_FULL_ATTN, 4.0 * (double)B * heads * (double)N * N * 32, 2.0 * (double)B * N * heads * 32 * 4, st);
  SRGD_CUDA_OK(launch_k(fa_tc_kernel, dim3(N / 128, B * heads), dim3(192), FaTcSmem::kTotal, st, kp));
  SRGD_LAUNCH_OK("fa_tc_kernel");
  count_launch();
  return SRGD_OK;
}
