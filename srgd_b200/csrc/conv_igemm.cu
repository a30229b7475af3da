// Convolution as implicit GEMM on the Blackwell 5th-gen tensor cores (tcgen05.mma, TMEM
// accumulators, TMA operand staging).  Replaces every nn.Conv2d on the reference hot path:
// conv3x3 (model.py:246, 647, 668), conv1x1 (271, 300, 303, 341, 342), Downsample's
// space-to-depth + 1x1 (106-110), PixelShuffleUpsample's 1x1 + SiLU + PixelShuffle (70-98) and the
// 7x7 init conv (583) after srgd_pack_input.
//
// GEMM view:  D[M = B*Ho*Wo pixels, N = Cout] = sum_phases A_phase[M, C_src] * W[N, k_phase..]^T.
// A "phase" is one filter tap of one source tensor: the A tile of a phase is the NHWC activation
// patch of the CTA's 128 output pixels shifted by (dy,dx); TMA's tiled mode fetches it as a 4-D box
// {64 ch, TW, TH, TN} whose out-of-range pixels are zero-filled (= conv zero padding), directly in
// the 128-byte-swizzled K-major layout tcgen05.mma consumes.  A channel concat (model.py:713-722)
// is just more phases reading a second source: the cat is never materialised.
//
// Kernel organisation (persistent, one CTA per SM, 384 threads):
//   warps 0..7    : epilogue      (tcgen05.ld 32x32b, row_scale, bias from shared memory, GroupNorm partial
//                                  sums pre-reduced to one record per tile, SiLU, residual prefetched a chunk
//                                  ahead, bf16 store in NHWC or pixel-shuffled NHWC through a per-warp
//                                  transpose tile); two warps per TMEM lane quarter, each taking every other
//                                  32-column chunk
//   warp 8 lane 0 : TMA producer  (STAGES-deep smem ring, full/empty mbarriers)
//   warp 9 lane 0 : MMA issuer    (4 x tcgen05.mma K=16 per 64-wide k-block; commit -> empty barrier)
//   warp 10       : TMEM allocator (2 accumulator stages of BN fp32 columns)
// (the latency-critical single-thread roles get the highest warp ids: the SM's warp arbiter favours them)
// The double-buffered accumulator lets the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tiles.h"
#include "tmap.h"

namespace srgd {

constexpr int kBM = 128;          // output pixels per tile (UMMA M)
constexpr int kBK = 64;           // bf16 channels per k-block (one 128-byte swizzle row)
constexpr int kThreads = 384;         // 8 epilogue warps + 4 control warps (TMA, MMA, TMEM alloc, spare)
constexpr int kABytes = kBM * kBK * 2;   // 16 KiB
constexpr int kTmaWarp = 8, kMmaWarp = 9, kAllocWarp = 10;

struct alignas(64) ConvKernelParams {
  CUtensorMap a_maps[SRGD_CONV_MAX_SRC];
  CUtensorMap w_map;
  int32_t n_src;
  int32_t n_phase;
  int32_t total_kblocks;
  int8_t ph_src[SRGD_CONV_MAX_PHASE];
  int8_t ph_dy[SRGD_CONV_MAX_PHASE];
  int8_t ph_dx[SRGD_CONV_MAX_PHASE];
  int16_t ph_cblocks[SRGD_CONV_MAX_PHASE];
  int32_t ph_kblk[SRGD_CONV_MAX_PHASE];
  int32_t B, Ho, Wo, Cout;
  int32_t tw_log2, th_log2;       // tile = TN x TH x TW pixels, TN*TH*TW = 128
  int32_t tiles_x, tiles_y, tiles_b;
  int32_t m_tiles, n_tiles, total_tiles;
  int32_t group_size;             // Cout / 8 (GroupNorm group width in channels)
  int32_t act, out_mode;
  const float* bias;
  const float* row_scale;
  const bf16* residual;
  bf16* out;
  float* gn_partials;
  // split-K of the partial last wave (conv_igemm_kernel only; split == 1: off).  Work items [0, n_full) are whole
  // tiles; item n_full + slot * split + part is K part `part` of tile n_full + slot.
  int16_t ph_iter0[SRGD_CONV_MAX_PHASE];   // linear k-block index at which phase ph starts
  int32_t n_full, split, kb_per_part, total_items;
  float* sk_part;                 // helper parts' fp32 accumulators: [slot][part - 1][BN columns][128 rows]
  int32_t* sk_flags;              // [0, 256): helper-warp arrivals per slot; [256, 512): consumer-warp departures
};

// Split-K work decomposition.  With T tiles on G persistent CTAs the last wave holds T % G tiles; when that is at
// most G / 2 the K loop of each of those tiles is cut into `split` parts that run on different CTAs at the same
// time: part 0 ("main") owns the epilogue, the other parts ("helpers") dump their fp32 accumulators to a workspace
// and raise a flag.  All items of the split wave are the LAST item of their CTA and every CTA of the grid is
// resident, so the main part's wait can never deadlock.
struct WorkItem {
  int tile, kb0, kb1, part, slot;   // part: -1 = whole tile, 0 = main part, > 0 = helper part
};
__device__ __forceinline__ WorkItem decode_item(const ConvKernelParams& p, int item) {
  WorkItem w;
  if (item < p.n_full) {
    w.tile = item; w.kb0 = 0; w.kb1 = p.total_kblocks; w.part = -1; w.slot = 0;
  } else {
    const int r = item - p.n_full;
    w.slot = r / p.split;
    w.part = r - w.slot * p.split;
    w.tile = p.n_full + w.slot;
    w.kb0 = w.part * p.kb_per_part;
    w.kb1 = min(w.kb0 + p.kb_per_part, p.total_kblocks);
  }
  return w;
}
__device__ __forceinline__ int ld_acquire_gpu(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_cg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

constexpr int kMaxBiasSmem = 2048;   // widest conv of the U-Net (pixel-shuffle 1024 -> 2048)

template <int BN, int STAGES>
struct ConvSmem {
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem ptr, gn staging
  static constexpr int kGnOffset = kBarOffset + (2 * STAGES + 4) * 8 + 16;
  static constexpr int kGnBytes = 2 * 8 * (BN / 8) * 2 * 4;  // two parities x one staging row per epilogue warp
  static constexpr int kBiasOffset = kGnOffset + kGnBytes;   // the whole bias vector (Cout <= kMaxBiasSmem floats)
  static constexpr int kStoreOffset = kBiasOffset + kMaxBiasSmem * 4;   // per epilogue warp: 32 rows x 64 B transpose tile
  static constexpr int kTotal = kStoreOffset + 8 * 2048 + 1024;          // +1024: manual 1 KiB alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ ConvKernelParams p) {
  using L = ConvSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* gn_smem = reinterpret_cast<float*>(smem + L::kGnOffset);
  float* bias_smem = reinterpret_cast<float*>(smem + L::kBiasOffset);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * BN;                 // 128 / 256 / 512: powers of two >= 32

  // The bias vector is read by every epilogue chunk: a global load there is a ~500-clock long-scoreboard stall per
  // chunk (ncu source view of the short-K launches: 30 % of all samples); staged once, it is a broadcast LDS.
  // Weights are never written by a kernel, so this may precede pdl_wait().
  const bool bias_staged = p.bias != nullptr && p.Cout <= kMaxBiasSmem;
  if (bias_staged)
    for (int i = threadIdx.x; i < p.Cout; i += kThreads) bias_smem[i] = __ldg(p.bias + i);

  if (warp == kTmaWarp && lane == 0) {
    for (int i = 0; i < p.n_src; ++i) ptx::prefetch_tmap(&p.a_maps[i]);
    ptx::prefetch_tmap(&p.w_map);
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 8);                 // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == kAllocWarp) {
    ptx::tmem_alloc(tmem_ptr_smem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  const int tw = 1 << p.tw_log2, th = 1 << p.th_log2;
  const int tn_log2 = 7 - p.tw_log2 - p.th_log2;

  if (warp == kTmaWarp && lane == 0) {
    // ===================================== TMA producer =====================================
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const WorkItem wi = decode_item(p, item);
      const int tile = wi.tile;
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int tx = m_tile % p.tiles_x;
      const int ty = (m_tile / p.tiles_x) % p.tiles_y;
      const int tb = m_tile / (p.tiles_x * p.tiles_y);
      const int x0 = tx << p.tw_log2, y0 = ty << p.th_log2, b0 = tb << tn_log2;
      for (int ph = 0; ph < p.n_phase; ++ph) {
        const CUtensorMap* amap = &p.a_maps[p.ph_src[ph]];
        const int xs = x0 + p.ph_dx[ph], ys = y0 + p.ph_dy[ph];
        const int kblk0 = p.ph_kblk[ph];
        const int ncb = p.ph_cblocks[ph];
        const int it0 = p.ph_iter0[ph];                    // this item's share of the phase (whole phase unless split)
        const int cb_lo = max(0, wi.kb0 - it0), cb_hi = min(ncb, wi.kb1 - it0);
        for (int cb = cb_lo; cb < cb_hi; ++cb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + stage * L::kStageBytes;
          uint8_t* b_dst = a_dst + kABytes;
          ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
          ptx::tma_load_4d(a_dst, amap, &full_bar[stage], cb * kBK, xs, ys, b0);
          ptx::tma_load_2d(b_dst, &p.w_map, &full_bar[stage], (kblk0 + cb) * kBK, n_tile * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == kMmaWarp && lane == 0) {
    // ====================================== MMA issuer ======================================
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(kBM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const WorkItem wi = decode_item(p, item);
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);   // epilogue drained this accumulator
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = wi.kb0; kb < wi.kb1; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + stage * L::kStageBytes);
        const uint64_t a_desc = ptx::make_sw128_kmajor_desc(a_addr);
        const uint64_t b_desc = ptx::make_sw128_kmajor_desc(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
          ptx::umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, ((kb - wi.kb0) | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);             // smem slot reusable once these MMAs retire
        if (kb == wi.kb1 - 1) ptx::umma_commit(&tfull_bar[acc]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 8) {
    // ======================================= epilogue =======================================
    // 8 warps: two per TMEM lane quarter; `half` takes every other 32-column chunk (the short-K 1x1 convs are
    // epilogue-bound with 4 warps: ncu showed 17-31 % tensor-active on the pixel-shuffle / res_conv launches)
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int half = warp >> 2;
    const int r = q * 32 + lane;                         // tile row == output pixel slot
    const int w_i = r & (tw - 1);
    const int h_i = (r >> p.tw_log2) & (th - 1);
    const int n_i = r >> (p.tw_log2 + p.th_log2);
    constexpr int kGnRow = (BN / 8) * 2;                 // floats per staging row: [BN/8][2]
    int gn_par = 0;                                      // staging rows are double-buffered over tiles
    // Register copies of the parameters the chunk loop uses.  Left to itself the compiler re-reads them from the
    // constant bank inside the loop (LDC, long scoreboard: 7 % of the samples of the short-K launches); the empty
    // asm pins them.
    int Cout = p.Cout, Ho = p.Ho, Wo = p.Wo, act = p.act, out_mode = p.out_mode, group_size = p.group_size;
    bf16* outp = p.out;
    const bf16* resp = p.residual;
    float* gnp = p.gn_partials;
    asm volatile("" : "+r"(Cout), "+r"(Ho), "+r"(Wo), "+r"(act), "+r"(out_mode), "+r"(group_size));
    asm volatile("" : "+l"(outp), "+l"(resp), "+l"(gnp));
    const uint32_t bias_addr = ptx::smem_u32(bias_smem);
    // Output stores go through a per-warp 32 x 64 B transpose tile: a thread owns one pixel row, so storing its 64 B
    // directly makes every store instruction touch 32 different lines (the L1 store path bounds the short-K convs);
    // read back transposed, four lanes cover one row's 64 B and an instruction touches 8 lines.  16-byte pieces
    // are XOR-swizzled by (row >> 1) & 3, which keeps both directions free of bank conflicts.
    const uint32_t stile = ptx::smem_u32(smem + L::kStoreOffset) + warp * 2048;
    const int srow = lane >> 2, spiece = lane & 3;         // read-back role: row 8k + srow, piece spiece
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const WorkItem wi = decode_item(p, item);
      const int tile = wi.tile;
      if (wi.part > 0) {
        // -------- split-K helper part: raw fp32 accumulator -> workspace, column-major inside the tile so that a
        // warp's 32 rows of one column are one 128-byte line; then one release-arrival per warp --------
        ptx::mbar_wait(&tfull_bar[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t t_row_h = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
        float* dst = p.sk_part + ((int64_t)(wi.slot * (p.split - 1) + wi.part - 1) * BN) * 128 + r;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_row_h + c * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(dst + (c * 32 + j) * 128, __uint_as_float(v[j]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          ptx::mbar_arrive(&tempty_bar[acc]);
          __threadfence();
          atomicAdd(p.sk_flags + wi.slot, 1);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      float* gn_w = gn_smem + (gn_par * 8 + warp) * kGnRow;   // this warp's staging row for this tile
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int tx = m_tile % p.tiles_x;
      const int ty = (m_tile / p.tiles_x) % p.tiles_y;
      const int tb = m_tile / (p.tiles_x * p.tiles_y);
      const int x = (tx << p.tw_log2) + w_i;
      const int y = (ty << p.th_log2) + h_i;
      const int b = (tb << tn_log2) + n_i;
      const bool valid = (x < Wo) && (y < Ho) && (b < p.B);
      const int64_t pix = ((int64_t)b * Ho + y) * Wo + x;
      const int n0 = n_tile * BN;
      const float rs = (p.row_scale != nullptr && valid) ? p.row_scale[pix] : 1.0f;
      // element offsets of the four rows this lane writes back (channel 0 of the pixel; pixel-shuffle: of its 2x2 block)
      int64_t sbase[4];
      bool svalid[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rk = q * 32 + 8 * k + srow;
        const int xk = (tx << p.tw_log2) + (rk & (tw - 1));
        const int yk = (ty << p.th_log2) + ((rk >> p.tw_log2) & (th - 1));
        const int bk = (tb << tn_log2) + (rk >> (p.tw_log2 + p.th_log2));
        svalid[k] = (xk < Wo) && (yk < Ho) && (bk < p.B);
        sbase[k] = (out_mode == SRGD_OUT_PIXEL_SHUFFLE)
                       ? (((int64_t)bk * (2 * Ho) + 2 * yk) * (2 * Wo) + 2 * xk) * (Cout >> 2)
                       : (((int64_t)bk * Ho + yk) * Wo + xk) * Cout;
      }

      if (gnp != nullptr) {
        for (int i = lane; i < (BN / 8) * 2; i += 32) gn_w[i] = 0.f;
        __syncwarp();
      }

      // residual rows are requested one chunk ahead (and the first one before the accumulator is even complete):
      // a load issued where it is consumed is a full memory latency on the epilogue's critical path
      const bool has_res = resp != nullptr && valid && out_mode != SRGD_OUT_PIXEL_SHUFFLE;
      uint4 rres[4];
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rres[j] = ld_stream(resp + pix * Cout + n0 + half * 32 + j * 8);
      }

      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      const float* sk_src = nullptr;
      if (wi.part == 0) {
        // split-K main part: the helper parts of this tile have raised 8 arrivals each (bounded spin, like mbar_wait)
        if (lane == 0) {
          const int want = 8 * (p.split - 1);
          const long long t0 = clock64();
          while (ld_acquire_gpu(p.sk_flags + wi.slot) < want) {
            // ~60 s: the partner is another CTA that may be slowed down 100x by compute-sanitizer's instrumentation
            if (clock64() - t0 > 120000000000LL) {
              printf("srgd_b200: split-K wait timed out (block %d slot %d)\n", (int)blockIdx.x, wi.slot);
              __trap();
            }
          }
        }
        __syncwarp();
        sk_src = p.sk_part + ((int64_t)(wi.slot * (p.split - 1)) * BN) * 128 + r;
      }

#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_row + c * 32, v);
        ptx::tmem_ld_wait();
        const int nc = n0 + c * 32;                      // first output channel of this chunk
        float f[32];
        if (sk_src != nullptr) {                         // + the helpers' partial sums, in part order
          for (int h = 0; h < p.split - 1; ++h) {
            const float* src = sk_src + ((int64_t)h * BN + c * 32) * 128;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + ld_cg_f32(src + j * 128));
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * rs;
        if (bias_staged) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = ptx::lds_f4(bias_addr + (nc + j) * 4);       // explicit ld.shared (a generic LD is a long-scoreboard op)
            f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
          }
        } else if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + nc + j));
            f[j] += bv.x; f[j + 1] += bv.y; f[j + 2] += bv.z; f[j + 3] += bv.w;
          }
        }
        if (gnp != nullptr) {
          // per 8-channel sub-block sums over this warp's 32 pixels (masked rows contribute 0)
          float s8[4], q8[4];
#pragma unroll
          for (int sb = 0; sb < 4; ++sb) {
            float s = 0.f, qq = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float t = valid ? f[sb * 8 + j] : 0.f;
              s += t;
              qq += t * t;
            }
            s8[sb] = s;
            q8[sb] = qq;
          }
          if (group_size >= 32) {                      // whole chunk lies in one group
            float s = warp_sum(s8[0] + s8[1] + s8[2] + s8[3]);
            float qq = warp_sum(q8[0] + q8[1] + q8[2] + q8[3]);
            if (lane == 0) {
              const int g = (c * 32) / group_size;     // group index local to this tile
              gn_w[g * 2] += s;
              gn_w[g * 2 + 1] += qq;
            }
          } else {
#pragma unroll
            for (int sb = 0; sb < 4; ++sb) {
              float s = warp_sum(s8[sb]);
              float qq = warp_sum(q8[sb]);
              if (lane == 0) {
                const int g = (c * 32 + sb * 8) / group_size;
                gn_w[g * 2] += s;
                gn_w[g * 2 + 1] += qq;
              }
            }
          }
        }
        if (act == 1) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) silu2(f[j], f[j + 1]);
        }
        if (has_res) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float rr[8];
            unpack8(rres[j >> 3], rr);
#pragma unroll
            for (int t = 0; t < 8; ++t) f[j + t] += rr[t];
          }
          if (c + 2 < BN / 32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rres[j] = ld_stream(resp + pix * Cout + nc + 64 + j * 8);   // chunk c + 2
          }
        }
        // channel offset of this chunk inside a row (pixel-shuffle: plus the sub-pixel's displacement)
        int64_t coff;
        if (out_mode == SRGD_OUT_PIXEL_SHUFFLE) {
          const int cq = Cout >> 2;                      // C' output channels
          const int sub = nc / cq;                       // (i*2 + j) sub-pixel
          coff = ((int64_t)(sub >> 1) * (2 * Wo) + (sub & 1)) * cq + (nc - sub * cq);
        } else {
          coff = nc;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 pk = pack8(f + 8 * j);
          ptx::sts_v4(stile + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), pk.x, pk.y, pk.z, pk.w);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rl = 8 * k + srow;
          const uint4 pk = ptx::lds_v4(stile + rl * 64 + ((spiece ^ ((rl >> 1) & 3)) << 4));
          if (svalid[k]) st_stream(outp + sbase[k] + coff + spiece * 8, pk);
        }
        __syncwarp();                                    // the tile is rewritten by the next chunk
      }
      // accumulator fully read: hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if (wi.part == 0 && lane == 0) {
        // the last of the 8 consumer warps re-arms the slot's flags for the next launch that uses this workspace
        if (atomicAdd(p.sk_flags + 256 + wi.slot, 1) == 7) {
          p.sk_flags[wi.slot] = 0;
          p.sk_flags[256 + wi.slot] = 0;
        }
      }

      if (gnp != nullptr) {
        // One 64-byte record [8 groups][sum, sumsq] per (M tile, sample slot): the eight warps' staging rows are
        // folded in a fixed order by warp `slot` (a tile spans 1 << tn_log2 samples; quarter q belongs to slot
        // q >> (2 - tn_log2)).  The rows of this parity are not touched again before the next-but-one tile, and
        // every warp passes the barrier of the next tile in between, so one barrier per tile suffices.
        __syncwarp();
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (warp < (1 << tn_log2)) {
          const int groups_in_tile = BN / group_size;
          const int g0 = n0 / group_size;
          const int qshift = 2 - tn_log2;
          float* dst = gnp + (((int64_t)m_tile << tn_log2) + warp) * 16;
          for (int i = lane; i < groups_in_tile * 2; i += 32) {
            float sum = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w)
              if (((w & 3) >> qshift) == warp) sum += gn_smem[(gn_par * 8 + w) * kGnRow + i];
            const int g = g0 + (i >> 1);
            if (g < 8) dst[g * 2 + (i & 1)] = sum;
          }
        }
        gn_par ^= 1;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// "Channels-as-M" variant for Cout tiles of exactly 128 (Cout = 128, 384):
//   D^T[M = 128 output channels, N = 256 pixels] = W_tile[128, K] * Act[256 pixels, K]^T
// With pixels as M and N = 128 the tensor core reads A (4 KiB) + B (4 KiB) of shared memory per
// 64-clock K=16 instruction = 128 B/clk, the whole shared-memory bandwidth, while TMA is writing the
// same memory: measured 650 TFLOP/s on the 128->128 convs at 256^2 (profiles/r01_launches_v1.md).
// Making the two 128-pixel sub-tiles the N = 256 operand drops that to 96 B/clk (like the
// Cout >= 256 tiles, which run at ~1.4 PFLOP/s) and halves the weight traffic per pixel.
// The accumulator comes out channel-major (TMEM lane = channel), so the epilogue transposes
// through a 32 KiB shared-memory staging tile and writes fully coalesced 256-byte NHWC rows.
// ---------------------------------------------------------------------------------------------
constexpr int kTN = 256;                     // pixels per tile (two sub-tiles of 128)
// HALO variant (3x3 convs on W %% 256 == 0 images, one image-row segment of 256 pixels per tile): a pipeline stage is
// one (source, 64-channel block, dy) "row stage": the 258-pixel row segment x0-1 .. x0+256 is fetched ONCE and the
// three horizontal taps are three tcgen05.mma operand descriptors whose start address is shifted by 0 / 1 / 2
// rows of 128 bytes inside the swizzled tile (the swizzle is a function of the absolute shared-memory address, so a
// row-shifted start needs no descriptor change: tests/gpu_probe_umma_shift.py).  The plain variant re-reads every
// activation nine times from L2 (48 KB per 512 MMA clocks and SM, 2/3 of it unique data) and was L2-bound at ~60 %
// tensor-active on the 128->128 convs; this one moves 82 KB per 1536 clocks.
template <bool HALO>
struct ConvSmemT {
  static constexpr int kStages = HALO ? 2 : 4;
  static constexpr int kWBytes = (HALO ? 3 : 1) * 128 * kBK * 2;    // weights: A operand, 16 KiB per tap
  static constexpr int kXBytes = HALO ? 33 * 1024 : kTN * kBK * 2;  // activations: B operand (258 rows padded / 256)
  static constexpr int kXTxBytes = HALO ? 258 * 128 : kTN * kBK * 2;
  static constexpr int kStageBytes = kWBytes + kXBytes;
  static constexpr int kStagingOffset = kStages * kStageBytes;
  static constexpr int kStagingBytes = 128 * 128 * 2;      // 2 halves x [64 pixels][128 ch] bf16
  static constexpr int kBarOffset = kStagingOffset + kStagingBytes;
  static constexpr int kTotal = kBarOffset + (2 * kStages + 4) * 8 + 16 + 1024;
};

template <bool HALO>
__global__ void __launch_bounds__(kThreads, 1) conv_igemm_t_kernel(const __grid_constant__ ConvKernelParams p) {
  using L = ConvSmemT<HALO>;
  constexpr int kTStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  bf16* staging = reinterpret_cast<bf16*>(smem + L::kStagingOffset);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kTStages;
  uint64_t* tfull_bar = empty_bar + kTStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;                      // 2 accumulator stages x 256 pixel columns

  if (warp == kTmaWarp && lane == 0) {
    for (int i = 0; i < p.n_src; ++i) ptx::prefetch_tmap(&p.a_maps[i]);
    if (HALO)
      for (int i = 0; i < p.n_src; ++i) ptx::prefetch_tmap(&p.a_maps[2 + i]);
    ptx::prefetch_tmap(&p.w_map);
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kTStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kAllocWarp) {
    ptx::tmem_alloc(tmem_ptr_smem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  const int tw = 1 << p.tw_log2, th = 1 << p.th_log2;
  const int tn_log2 = 7 - p.tw_log2 - p.th_log2;
  const int tiles_xy = p.tiles_x * p.tiles_y;
  // tile -> (pair of pixel sub-tiles, 128-channel slab); n fastest so concurrent CTAs share activations

  if (warp == kTmaWarp && lane == 0) {
    // ===================================== TMA producer =====================================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int pair = tile / p.n_tiles;
      int x0[2], y0[2], b0[2];
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int st = pair * 2 + s;                       // may be == m_tiles (odd count): all OOB -> zeros
        x0[s] = (st % p.tiles_x) << p.tw_log2;
        y0[s] = ((st / p.tiles_x) % p.tiles_y) << p.th_log2;
        b0[s] = (st / tiles_xy) << tn_log2;
      }
      if (HALO) {
        // phases are ordered (ky, kx, source) (checked by the launcher); sub-tile 1 continues sub-tile 0's row
        for (int src = 0; src < p.n_src; ++src) {
          const int ncb = p.ph_cblocks[src];
          for (int cb = 0; cb < ncb; ++cb) {
            for (int dyi = 0; dyi < 3; ++dyi) {
              ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* w_dst = smem + stage * L::kStageBytes;
              uint8_t* x_dst = w_dst + L::kWBytes;
              ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kWBytes + L::kXTxBytes);
#pragma unroll
              for (int dxi = 0; dxi < 3; ++dxi)
                ptx::tma_load_2d(w_dst + dxi * 16384, &p.w_map, &full_bar[stage],
                                 (p.ph_kblk[(dyi * 3 + dxi) * p.n_src + src] + cb) * kBK, n_tile * 128);
              ptx::tma_load_4d(x_dst, &p.a_maps[src], &full_bar[stage], cb * kBK, x0[0] - 1, y0[0] + dyi - 1, b0[0]);
              ptx::tma_load_4d(x_dst + 256 * 128, &p.a_maps[2 + src], &full_bar[stage], cb * kBK, x0[0] + 255,
                               y0[0] + dyi - 1, b0[0]);
              if (++stage == kTStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      } else {
        for (int ph = 0; ph < p.n_phase; ++ph) {
          const CUtensorMap* amap = &p.a_maps[p.ph_src[ph]];
          const int dx = p.ph_dx[ph], dy = p.ph_dy[ph];
          const int kblk0 = p.ph_kblk[ph];
          const int ncb = p.ph_cblocks[ph];
          for (int cb = 0; cb < ncb; ++cb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* w_dst = smem + stage * L::kStageBytes;
            uint8_t* x_dst = w_dst + L::kWBytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            ptx::tma_load_2d(w_dst, &p.w_map, &full_bar[stage], (kblk0 + cb) * kBK, n_tile * 128);
            ptx::tma_load_4d(x_dst, amap, &full_bar[stage], cb * kBK, x0[0] + dx, y0[0] + dy, b0[0]);
            ptx::tma_load_4d(x_dst + kABytes, amap, &full_bar[stage], cb * kBK, x0[1] + dx, y0[1] + dy, b0[1]);
            if (++stage == kTStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp && lane == 0) {
    // ====================================== MMA issuer ======================================
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, kTN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * kTN;
      const int n_stages = HALO ? p.total_kblocks / 3 : p.total_kblocks;
      for (int kb = 0; kb < n_stages; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t w_addr = ptx::smem_u32(smem + stage * L::kStageBytes);
        if (HALO) {
#pragma unroll
          for (int dxi = 0; dxi < 3; ++dxi) {
            const uint64_t a_desc = ptx::make_sw128_kmajor_desc(w_addr + dxi * 16384);            // tap weights
            const uint64_t b_desc = ptx::make_sw128_kmajor_desc(w_addr + L::kWBytes + dxi * 128);  // row-shifted pixels
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              ptx::umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | dxi | k) != 0 ? 1u : 0u);
          }
        } else {
          const uint64_t a_desc = ptx::make_sw128_kmajor_desc(w_addr);                 // weights  [128 ch][64 k]
          const uint64_t b_desc = ptx::make_sw128_kmajor_desc(w_addr + L::kWBytes);    // pixels   [256 px][64 k]
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            ptx::umma_bf16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (kb == n_stages - 1) ptx::umma_commit(&tfull_bar[acc]);
        if (++stage == kTStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 8) {
    // ======================================= epilogue =======================================
    // 8 warps: q = TMEM lane quarter = 32 output channels; `half` = which 128-pixel sub-tile of the pair.
    const int q = warp & 3;
    const int half = warp >> 2;
    const int et = threadIdx.x & 127;                      // 0..127 within this half's four warps
    bf16* stg = staging + half * (64 * 128);               // this half's staging: [64 pixels][128 channels]
    const int bar_id = 1 + half;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int pair = tile / p.n_tiles;
      const int ch = n_tile * 128 + q * 32 + lane;         // this thread's output channel
      const float bias = (p.bias != nullptr) ? __ldg(p.bias + ch) : 0.f;
      const int st = pair * 2 + half;
      const int tx = st % p.tiles_x, ty = (st / p.tiles_x) % p.tiles_y, tb = st / tiles_xy;

      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kTN + half * 128;
      float gn_s = 0.f, gn_q = 0.f;

#pragma unroll 1
      for (int cp = 0; cp < 2; ++cp) {                     // 64 pixels per staging pass
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = cp * 2 + cc;                       // 32-pixel column chunk
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_row + c * 32, v);
          // lane j describes pixel column j of this chunk
          const int r = c * 32 + lane;
          const int px = (tx << p.tw_log2) + (r & (tw - 1));
          const int py = (ty << p.th_log2) + ((r >> p.tw_log2) & (th - 1));
          const int pb = (tb << tn_log2) + (r >> (p.tw_log2 + p.th_log2));
          const bool pvalid = (px < p.Wo) && (py < p.Ho) && (pb < p.B);
          const uint32_t vmask = __ballot_sync(0xffffffffu, pvalid);
          float rs_lane = 1.0f;
          if (p.row_scale != nullptr && pvalid) rs_lane = p.row_scale[((int64_t)pb * p.Ho + py) * p.Wo + px];
          ptx::tmem_ld_wait();
          float f[32];
          if (p.row_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * __shfl_sync(0xffffffffu, rs_lane, j) + bias;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + bias;
          }
          if (p.gn_partials != nullptr) {
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float t = ((vmask >> j) & 1u) ? f[j] : 0.f;
              sum += t;
              sq += t * t;
            }
            // reduce over the lanes of one GroupNorm group (16 or 32 channels)
            if (p.group_size >= 32) {
              sum += __shfl_xor_sync(0xffffffffu, sum, 16);
              sq += __shfl_xor_sync(0xffffffffu, sq, 16);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
              sum += __shfl_xor_sync(0xffffffffu, sum, o);
              sq += __shfl_xor_sync(0xffffffffu, sq, o);
            }
            // record layout shared with conv_igemm_kernel: [M tile][sample slot][8 groups][2]; the four 32-pixel
            // chunks of a sub-tile are accumulated in order and flushed when the sample slot changes
            gn_s += sum;
            gn_q += sq;
            const int qshift = 2 - tn_log2;
            if (c == 3 || ((c + 1) >> qshift) != (c >> qshift)) {
              const int lanes_per_group = p.group_size >= 32 ? 32 : 16;
              if ((lane & (lanes_per_group - 1)) == 0 && st < p.m_tiles) {
                const int g = ch / p.group_size;
                float* dst = p.gn_partials + (((((int64_t)st << tn_log2) + (c >> qshift)) * 8) + g) * 2;
                dst[0] = gn_s;
                dst[1] = gn_q;
              }
              gn_s = 0.f;
              gn_q = 0.f;
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) silu2(f[j], f[j + 1]);
          }
          // transpose: stg[pixel][channel]; one 64-byte row segment per store instruction
          bf16* col = stg + (cc * 32) * 128 + q * 32 + lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) col[j * 128] = __float2bfloat16(f[j]);
        }
        if (cp == 1) {
          // this warp has read its whole share of the accumulator: release it before the (slower) copy-out
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
        }
        // the four warps of this half have filled the staging tile -> coalesced copy-out of 64 rows x 256 B
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll 4
        for (int i = 0; i < 8; ++i) {
          const int vec = et + i * 128;                    // 1024 vectors of 16 B
          const int rr = vec >> 4, part = vec & 15;
          const int r = cp * 64 + rr;
          const int px = (tx << p.tw_log2) + (r & (tw - 1));
          const int py = (ty << p.th_log2) + ((r >> p.tw_log2) & (th - 1));
          const int pb = (tb << tn_log2) + (r >> (p.tw_log2 + p.th_log2));
          if ((px < p.Wo) && (py < p.Ho) && (pb < p.B)) {
            const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 128 + part * 8);
            bf16* dst = p.out + (((int64_t)pb * p.Ho + py) * p.Wo + px) * p.Cout + n_tile * 128 + part * 8;
            st_stream(dst, val);
          }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");    // staging is free again
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------
// CUDA-core direct evaluation of the same descriptor (debug / verification only)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv_direct_kernel(srgd_conv_desc d) {
  const int64_t total = (int64_t)d.B * d.Ho * d.Wo * d.Cout;
  const bf16* wt = reinterpret_cast<const bf16*>(d.weight);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx % d.Cout);
    const int64_t pix = idx / d.Cout;
    const int x = (int)(pix % d.Wo);
    const int y = (int)((pix / d.Wo) % d.Ho);
    const int b = (int)(pix / ((int64_t)d.Wo * d.Ho));
    float acc = 0.f;
    for (int ph = 0; ph < d.n_phase; ++ph) {
      const srgd_conv_src& s = d.srcs[d.phases[ph].src];
      const int yy = y + d.phases[ph].dy, xx = x + d.phases[ph].dx;
      if (yy < 0 || yy >= s.H || xx < 0 || xx >= s.W) continue;
      const bf16* a = reinterpret_cast<const bf16*>(s.ptr) + (int64_t)b * s.sb + (int64_t)yy * s.sy + (int64_t)xx * s.sx;
      const bf16* w = wt + (int64_t)n * d.Ktot + d.phases[ph].k_start;
      for (int c = 0; c < s.C; ++c) acc += __bfloat162float(a[c]) * __bfloat162float(w[c]);
    }
    if (d.row_scale) acc *= d.row_scale[pix];
    if (d.bias) acc += d.bias[n];
    if (d.act == 1) acc = silu_f(acc);
    int64_t off;
    if (d.out_mode == SRGD_OUT_PIXEL_SHUFFLE) {
      const int cq = d.Cout >> 2, sub = n / cq, cc = n - sub * cq;
      off = (((int64_t)b * (2 * d.Ho) + (2 * y + (sub >> 1))) * (2 * d.Wo) + (2 * x + (sub & 1))) * cq + cc;
    } else {
      off = pix * d.Cout + n;
    }
    if (d.residual) acc += __bfloat162float(reinterpret_cast<const bf16*>(d.residual)[off]);
    reinterpret_cast<bf16*>(d.out)[off] = __float2bfloat16(acc);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_rows, const char* what) {
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) {
    set_error("%s: cuTensorMapEncodeTiled not available from the driver", what);
    return SRGD_E_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)row_stride_bytes};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled(inner=%llu rows=%llu pitch=%llu box=%ux%u) failed with %d", what,
              (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_bytes, box_inner,
              box_rows, (int)r);
    return SRGD_E_CUDA;
  }
  return SRGD_OK;
}

// qkv tensor bf16 [rows][3*heads*32] viewed as [rows][3*heads segments][32 channels]; the box is 64 channels
// wide, so channels 32..63 of every shared-memory row are out of range and read as zero (attention_tc.cu).
int make_tmap_qk_heads(CUtensorMap* m, const void* qkv, int64_t rows, int heads) {
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) {
    set_error("attention_tc: cuTensorMapEncodeTiled not available from the driver");
    return SRGD_E_CUDA;
  }
  const cuuint64_t gdim[3] = {32, (cuuint64_t)(3 * heads), (cuuint64_t)rows};
  const cuuint64_t gstr[2] = {64, (cuuint64_t)(3 * heads * 32 * 2)};
  const cuuint32_t box[3] = {64, 1, 128};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention_tc: cuTensorMapEncodeTiled(rows=%lld heads=%d) failed with %d", (long long)rows, heads, (int)r);
    return SRGD_E_CUDA;
  }
  return SRGD_OK;
}

static int validate_desc(const srgd_conv_desc* d) {
  SRGD_REQUIRE(d != nullptr, "conv: null descriptor");
  SRGD_REQUIRE(d->B > 0 && d->Ho > 0 && d->Wo > 0, "conv: bad output extent %dx%dx%d", d->B, d->Ho, d->Wo);
  SRGD_REQUIRE(d->Cout > 0 && d->Cout % 64 == 0, "conv: Cout=%d must be a positive multiple of 64", d->Cout);
  SRGD_REQUIRE(d->n_src >= 1 && d->n_src <= SRGD_CONV_MAX_SRC, "conv: n_src=%d out of range", d->n_src);
  SRGD_REQUIRE(d->n_phase >= 1 && d->n_phase <= SRGD_CONV_MAX_PHASE, "conv: n_phase=%d out of range", d->n_phase);
  SRGD_REQUIRE(d->weight && d->out, "conv: null weight/out");
  SRGD_REQUIRE(((uintptr_t)d->weight | (uintptr_t)d->out | (uintptr_t)d->bias | (uintptr_t)d->residual) % 16 == 0,
               "conv: weight/out/bias/residual must be 16-byte aligned");
  for (int i = 0; i < d->n_src; ++i) {
    const srgd_conv_src& s = d->srcs[i];
    SRGD_REQUIRE(s.ptr && ((uintptr_t)s.ptr % 16) == 0, "conv: src %d pointer null or not 16-byte aligned", i);
    SRGD_REQUIRE(s.C > 0 && s.C % 64 == 0, "conv: src %d C=%d must be a multiple of 64", i, s.C);
    SRGD_REQUIRE(s.sx % 8 == 0 && s.sy % 8 == 0 && s.sb % 8 == 0 && s.sx > 0, "conv: src %d strides must be multiples of 8", i);
    SRGD_REQUIRE(s.H > 0 && s.W > 0, "conv: src %d bad extent", i);
  }
  for (int i = 0; i < d->n_phase; ++i) {
    const srgd_conv_phase& ph = d->phases[i];
    SRGD_REQUIRE(ph.src >= 0 && ph.src < d->n_src, "conv: phase %d src out of range", i);
    SRGD_REQUIRE(ph.k_start >= 0 && ph.k_start % 64 == 0 && ph.k_start + d->srcs[ph.src].C <= d->Ktot,
                 "conv: phase %d weight columns [%d,+%d) outside Ktot=%lld", i, ph.k_start, d->srcs[ph.src].C,
                 (long long)d->Ktot);
    SRGD_REQUIRE(ph.dy >= -64 && ph.dy <= 64 && ph.dx >= -64 && ph.dx <= 64, "conv: phase %d tap out of range", i);
  }
  SRGD_REQUIRE(d->Ktot % 64 == 0, "conv: Ktot must be a multiple of 64");
  SRGD_REQUIRE(d->out_mode == SRGD_OUT_BF16_NHWC || d->out_mode == SRGD_OUT_PIXEL_SHUFFLE, "conv: bad out_mode");
  if (d->out_mode == SRGD_OUT_PIXEL_SHUFFLE) {
    SRGD_REQUIRE(d->Cout % 128 == 0, "conv: pixel-shuffle output needs Cout %% 128 == 0");
    SRGD_REQUIRE(d->residual == nullptr, "conv: a residual cannot be combined with the pixel-shuffle output mode");
  }
  return SRGD_OK;
}

template <int BN, int STAGES>
static int launch_igemm(const ConvKernelParams& kp, cudaStream_t st) {
  using L = ConvSmem<BN, STAGES>;
  static uint64_t configured = 0;
  if (first_launch_on_device(configured)) {
    SRGD_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::kTotal));
  }
  int grid = kp.total_items < sm_count() ? kp.total_items : sm_count();
  SRGD_CUDA_OK(launch_k(conv_igemm_kernel<BN, STAGES>, dim3(grid), dim3(kThreads), L::kTotal, st, kp));
  count_launch();
  return SRGD_OK;
}

template <bool HALO>
static int launch_igemm_t(const ConvKernelParams& kp, cudaStream_t st) {
  using L = ConvSmemT<HALO>;
  static uint64_t configured = 0;
  if (first_launch_on_device(configured)) {
    SRGD_CUDA_OK(cudaFuncSetAttribute(conv_igemm_t_kernel<HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  }
  int grid = kp.total_tiles < sm_count() ? kp.total_tiles : sm_count();
  SRGD_CUDA_OK(launch_k(conv_igemm_t_kernel<HALO>, dim3(grid), dim3(kThreads), L::kTotal, st, kp));
  count_launch();
  return SRGD_OK;
}

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_conv_m_tiles(int32_t B, int32_t Ho, int32_t Wo) {
  if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
  const TileGeom g = tile_geom(B, Ho, Wo);
  return g.m_tiles << g.tn_log2;                         // gn_partials records: one per (M tile, sample slot)
}

extern "C" size_t srgd_conv_splitk_workspace_bytes(void) {
  // flags (4 KiB) + one 128 x 256 fp32 accumulator per helper part; helpers < CTAs of the grid <= SMs (160 covers B200)
  return 4096 + (size_t)160 * 128 * 256 * sizeof(float);
}

extern "C" int srgd_conv_igemm(const srgd_conv_desc* d, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  rc = validate_desc(d);
  if (rc) return rc;
  EncodeTiledFn encode = get_encode_tiled();
  if (encode == nullptr) {
    set_error("conv: cuTensorMapEncodeTiled not available from the driver");
    return SRGD_E_CUDA;
  }
  const TileGeom g = tile_geom(d->B, d->Ho, d->Wo);

  // channels-as-M variant for 128-wide Cout slabs (see conv_igemm_t_kernel)
  const char* force = getenv("SRGD_CONV_VARIANT");         // test knob: "n" = pixels-as-M only
  const bool swapped = d->Cout % 128 == 0 && d->Cout % 256 != 0 && d->residual == nullptr &&
                       d->out_mode == SRGD_OUT_BF16_NHWC && (d->gn_partials == nullptr || d->Cout / 8 <= 32) &&
                       !(force != nullptr && force[0] == 'n');
  int total_kb_all = 0;
  for (int i = 0; i < d->n_phase; ++i) total_kb_all += d->srcs[d->phases[i].src].C / kBK;
  // split-K needs the caller's workspace (flags zeroed once) and is a property of the pixels-as-M kernel
  const char* sk_env = getenv("SRGD_CONV_SPLITK");         // test knob: "0" = never split
  // (not in batch-invariant mode: whether a tile is split depends on the launch's tile count, i.e. on the batch, and
  // a split tile sums its products in another fp32 order)
  const bool sk_ok = d->splitk_ws != nullptr && !swapped && !(sk_env != nullptr && sk_env[0] == '0') &&
                     g_batch_invariant == 0 && d->splitk_ws_bytes >= (int64_t)srgd_conv_splitk_workspace_bytes();
  // pick the N tile: widest that divides Cout, but keep enough tiles to fill the SMs -- with split-K available a
  // short grid of 256-wide tiles (>= 24 of them, K parts of >= 8 k-blocks) fills the SMs through its K parts
  int BN = 64;
  const int64_t tiles256 = d->Cout % 256 == 0 ? (int64_t)g.m_tiles * (d->Cout / 256) : 0;
  if (tiles256 >= sm_count()) BN = 256;
  else if (sk_ok && tiles256 >= 24 && total_kb_all >= 32) BN = 256;
  else if (d->Cout % 128 == 0) BN = 128;
  if (d->gn_partials) {
    SRGD_REQUIRE(d->Cout % 8 == 0 && (d->Cout / 8) % 8 == 0, "conv: GroupNorm partials need Cout %% 64 == 0");
    SRGD_REQUIRE(g.tn_log2 <= 2, "conv: GroupNorm partials need H*W >= 32 (got %dx%d)", d->Ho, d->Wo);
    while (BN < d->Cout / 8) BN *= 2;                    // a tile must hold whole groups
      SRGD_REQUIRE(BN <= 256 && d->Cout % BN == 0, "conv: cannot tile Cout=%d for GroupNorm partials", d->Cout);
  }
  if (swapped) BN = 128;
  // halo-reuse variant: plain 3x3 window over 1 or 2 same-sized sources, phases ordered (ky, kx, source), tiles that
  // are image-row segments and pair up inside a row
  bool halo = swapped && d->n_src <= 2 && d->n_phase == 9 * d->n_src && g.tw_log2 == 7 && g.th_log2 == 0 &&
              d->Wo % 256 == 0 && !(force != nullptr && force[0] == 't');
  for (int i = 0; halo && i < d->n_phase; ++i) {
    const int tap = i / d->n_src, src = i % d->n_src;
    const srgd_conv_phase& ph = d->phases[i];
    halo = ph.src == src && ph.dy == tap / 3 - 1 && ph.dx == tap % 3 - 1 && d->srcs[src].H == d->Ho &&
           d->srcs[src].W == d->Wo;
  }

  ConvKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < d->n_src; ++i) {
    const srgd_conv_src& s = d->srcs[i];
    const cuuint64_t gdim[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)d->B};
    const cuuint64_t gstr[3] = {(cuuint64_t)s.sx * 2, (cuuint64_t)s.sy * 2, (cuuint64_t)s.sb * 2};
    const cuuint32_t box[4] = {(cuuint32_t)kBK, halo ? 256u : 1u << g.tw_log2, 1u << g.th_log2, 1u << g.tn_log2};
    CUresult r = encode(&kp.a_maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.ptr), gdim, gstr, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv: cuTensorMapEncodeTiled(src %d: C=%d W=%d H=%d B=%d sx=%lld sy=%lld sb=%lld) failed with %d", i,
                s.C, s.W, s.H, d->B, (long long)s.sx, (long long)s.sy, (long long)s.sb, (int)r);
      return SRGD_E_CUDA;
    }
  }
  if (halo) {                                               // 2-pixel tail boxes of the 258-pixel row segments
    for (int i = 0; i < d->n_src; ++i) {
      const srgd_conv_src& s = d->srcs[i];
      const cuuint64_t gdim[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)d->B};
      const cuuint64_t gstr[3] = {(cuuint64_t)s.sx * 2, (cuuint64_t)s.sy * 2, (cuuint64_t)s.sb * 2};
      const cuuint32_t box[4] = {(cuuint32_t)kBK, 2u, 1u, 1u};
      CUresult r = encode(&kp.a_maps[2 + i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s.ptr), gdim, gstr,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("conv: cuTensorMapEncodeTiled(halo tail, src %d) failed with %d", i, (int)r);
        return SRGD_E_CUDA;
      }
    }
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)d->Ktot, (cuuint64_t)d->Cout};
    const cuuint64_t gstr[1] = {(cuuint64_t)d->Ktot * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)BN};
    CUresult r = encode(&kp.w_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->weight), gdim, gstr, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv: cuTensorMapEncodeTiled(weight Ktot=%lld Cout=%d) failed with %d", (long long)d->Ktot, d->Cout,
                (int)r);
      return SRGD_E_CUDA;
    }
  }
  kp.n_src = d->n_src;
  kp.n_phase = d->n_phase;
  int total_kb = 0;
  for (int i = 0; i < d->n_phase; ++i) {
    kp.ph_src[i] = (int8_t)d->phases[i].src;
    kp.ph_dy[i] = (int8_t)d->phases[i].dy;
    kp.ph_dx[i] = (int8_t)d->phases[i].dx;
    kp.ph_cblocks[i] = (int16_t)(d->srcs[d->phases[i].src].C / kBK);
    kp.ph_kblk[i] = d->phases[i].k_start / kBK;
    total_kb += kp.ph_cblocks[i];
  }
  kp.total_kblocks = total_kb;
  kp.B = d->B; kp.Ho = d->Ho; kp.Wo = d->Wo; kp.Cout = d->Cout;
  kp.tw_log2 = g.tw_log2; kp.th_log2 = g.th_log2;
  kp.tiles_x = g.tiles_x; kp.tiles_y = g.tiles_y; kp.tiles_b = g.tiles_b;
  kp.m_tiles = g.m_tiles;
  kp.n_tiles = (d->Cout + BN - 1) / BN;
  kp.total_tiles = (swapped ? (kp.m_tiles + 1) / 2 : kp.m_tiles) * kp.n_tiles;
  {
    int it = 0;
    for (int i = 0; i < d->n_phase; ++i) {
      kp.ph_iter0[i] = (int16_t)it;
      it += kp.ph_cblocks[i];
    }
  }
  // split-K of the partial last wave (see WorkItem): rem tiles on sm_count CTAs, rem <= sm_count / 2
  kp.n_full = kp.total_tiles; kp.split = 1; kp.kb_per_part = total_kb; kp.total_items = kp.total_tiles;
  if (sk_ok) {
    const int G = sm_count();
    const int rem = kp.total_tiles % G;
    int split = rem > 0 ? G / rem : 1;
    if (split > 4) split = 4;
    while (split > 1 && total_kb / split < 8) --split;     // parts of at least 8 k-blocks
    if (split > 1) {
      int per = (total_kb + split - 1) / split;
      while (split > 1 && (split - 1) * per >= total_kb) { --split; per = (total_kb + split - 1) / split; }
      if (split > 1 && rem <= 256) {
        kp.n_full = kp.total_tiles - rem;
        kp.split = split;
        kp.kb_per_part = per;
        kp.total_items = kp.n_full + rem * split;
        kp.sk_flags = reinterpret_cast<int32_t*>(d->splitk_ws);
        kp.sk_part = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d->splitk_ws) + 4096);
      }
    }
  }
  kp.group_size = d->Cout / 8;
  kp.act = d->act; kp.out_mode = d->out_mode;
  kp.bias = d->bias; kp.row_scale = d->row_scale;
  kp.residual = reinterpret_cast<const bf16*>(d->residual);
  kp.out = reinterpret_cast<bf16*>(d->out);
  kp.gn_partials = d->gn_partials;

  cudaStream_t st = as_stream(stream);
  const double M = (double)d->B * d->Ho * d->Wo;
  ProfScope prof(SRGD_PK_CONV, 2.0 * M * d->Cout * total_kb * kBK,
                 2.0 * (M * total_kb * kBK + (double)d->Cout * d->Ktot + M * d->Cout), st);
  if (swapped) return halo ? launch_igemm_t<true>(kp, st) : launch_igemm_t<false>(kp, st);
  switch (BN) {
    case 64: return launch_igemm<64, 8>(kp, st);
    case 128: return launch_igemm<128, 6>(kp, st);
    default: return launch_igemm<256, 4>(kp, st);
  }
}

extern "C" int srgd_conv_direct(const srgd_conv_desc* d, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  rc = validate_desc(d);
  if (rc) return rc;
  SRGD_REQUIRE(d->gn_partials == nullptr, "conv_direct: GroupNorm partials are only produced by srgd_conv_igemm");
  const int64_t total = (int64_t)d->B * d->Ho * d->Wo * d->Cout;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 32) blocks = (int64_t)sm_count() * 32;
  conv_direct_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(*d);
  SRGD_LAUNCH_OK("conv_direct_kernel");
  count_launch();
  return SRGD_OK;
}
