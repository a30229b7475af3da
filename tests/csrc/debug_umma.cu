// Micro-experiment (test-only probe library built by tests/gpu_probe_umma_shift.py; NOT part of libsrgd_b200.so): does tcgen05.mma accept a K-major
// SWIZZLE_128B operand whose start address is shifted by whole 128-byte rows (not 1024-byte aligned), and which
// value of the descriptor's base_offset field (bits [49,52)) does it need?  A 3x3 convolution could then take its
// three horizontal taps from ONE shared-memory copy of an image-row segment (halo reuse) instead of three TMA loads.
//   D[128][128] = W[128][64] * X[shift : shift+128][64]^T   for shift in 0..7, base_offset mode in {0, shift, 8-shift}
#include <cuda.h>
#include <string.h>

#include "../../srgd_b200/csrc/common.cuh"
#include "../../srgd_b200/csrc/ptx.cuh"
#include "../../srgd_b200/csrc/tmap.h"

namespace srgd {

__global__ void __launch_bounds__(128, 1) debug_umma_shift_kernel(const __grid_constant__ CUtensorMap w_map,
                                                                  const __grid_constant__ CUtensorMap x_map,
                                                                  float* out, int shift, int base_off) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 16384 + 17408);
  uint64_t* done = full + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(full, 1);
    ptx::mbar_init(done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_ptr, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(full, 16384 + 136 * 128);
    ptx::tma_load_2d(smem, &w_map, full, 0, 0);                 // W: [128 rows][64]
    ptx::tma_load_2d(smem + 16384, &x_map, full, 0, 0);         // X: [136 rows][64]
    ptx::mbar_wait(full, 0);
    ptx::tc_fence_after();
    const uint64_t ad = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem));
    uint64_t bd = ptx::make_sw128_kmajor_desc(ptx::smem_u32(smem + 16384 + shift * 128));
    bd |= (uint64_t)(base_off & 7) << 49;
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 128);
#pragma unroll
    for (int k = 0; k < 4; ++k) ptx::umma_bf16_ss(tmem_base, ad + 2 * k, bd + 2 * k, idesc, k != 0 ? 1u : 0u);
    ptx::umma_commit(done);
  }
  ptx::mbar_wait(done, 0);
  ptx::tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[row * 128 + c * 32 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace srgd

using namespace srgd;

// w: bf16 [128][64], x: bf16 [136][64], out: fp32 [128][128] (device pointers)
extern "C" int srgd_debug_umma_shift(const void* w, const void* x, float* out, int32_t shift, int32_t base_off,
                                     srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(w && x && out && shift >= 0 && shift <= 8, "debug_umma_shift: bad arguments");
  CUtensorMap wm, xm;
  rc = make_tmap_2d_bf16(&wm, w, 64, 128, 128, 64, 128, "debug_umma_shift(w)");
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&xm, x, 64, 136, 128, 64, 136, "debug_umma_shift(x)");
  if (rc) return rc;
  const int smem = 16384 + 17408 + 64 + 1024;
  SRGD_CUDA_OK(cudaFuncSetAttribute(debug_umma_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  debug_umma_shift_kernel<<<1, 128, smem, as_stream(stream)>>>(wm, xm, out, shift, base_off);
  SRGD_LAUNCH_OK("debug_umma_shift_kernel");
  return SRGD_OK;
}
