// Shared helpers for the srgd_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/srgd_b200.h"

namespace srgd {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_device();                       // SRGD_OK if current device is sm_100, cached
int fail_cuda(cudaError_t e, const char* what);

#define SRGD_CUDA_OK(expr)                                         \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return ::srgd::fail_cuda(_e, #expr);    \
  } while (0)

#define SRGD_REQUIRE(cond, ...)                                    \
  do {                                                             \
    if (!(cond)) {                                                 \
      ::srgd::set_error(__VA_ARGS__);                              \
      return SRGD_E_ARG;                                           \
    }                                                              \
  } while (0)

#define SRGD_LAUNCH_OK(what)                                       \
  do {                                                             \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return ::srgd::fail_cuda(_e, what);     \
  } while (0)

static inline cudaStream_t as_stream(srgd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE function attribute and the library may be used on
// several GPUs from one process (srgd_device_check / ConditionalSRUnet.to(device)): a launcher keeps one bit per
// device index in a static mask and (re)configures its kernel on the first launch on each device.
static inline bool first_launch_on_device(uint64_t& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

// srgd_set_batch_invariant(): when non-zero, every reduction whose partition would otherwise depend on the batch size
// (the LinearAttention context partials: splits per sample = SMs / B) uses the partition of a fixed reference batch,
// so a row's result is bit-identical no matter which other rows share its launch.
extern int g_batch_invariant;
constexpr int kInvariantRefBatch = 16;

// launch counter (gpu_launches in bench.py / srgd_unet_last_launch_count)
extern thread_local long g_launches;
static inline void count_launch(int n = 1) { g_launches += n; }

// Optional per-launch timing (srgd_profile_*): a ProfScope brackets the launches made inside an API
// entry point with CUDA events when profiling is on; otherwise it costs one branch.
extern bool g_prof_on;
void prof_open(int kind, double flops, double bytes, cudaStream_t st);
void prof_close(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  bool on;
  ProfScope(int kind, double flops, double bytes, cudaStream_t s) : st(s), on(g_prof_on) {
    if (on) prof_open(kind, flops, bytes, st);
  }
  ~ProfScope() {
    if (on) prof_close(st);
  }
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Programmatic dependent launch: kernel N+1 is made resident while the last CTAs of kernel N drain, runs its
// prologue (barrier init, TMEM allocation, tensor-map prefetch) and blocks in pdl_wait() until kernel N has
// completed and flushed.  ~170 launches per U-Net forward make the launch gap 1-12 % of a step (B = 16 .. 1 tiles).
// Rule for every kernel launched through launch_k(): ALL threads call pdl_wait() before the first access to
// global memory that another kernel may have written, and before the first global write.  No kernel issues
// griddepcontrol.launch_dependents: an early trigger (dependents resident and parked for the whole primary) was
// measured SLOWER than stream order at B >= 8 and the implicit trigger at CTA exit faster at every batch
// (profiles/r01_pdl_ab.txt).  SRGD_PDL=0 in the environment restores plain stream serialisation.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

// SiLU of two values with ONE MUFU op: x*sigmoid(x) = 0.5x(1 + tanh(x/2)), tanh evaluated as a packed
// half2 (tanh.approx.f16x2, abs err ~2^-11).  B200 issues 16 MUFU/clk/SM, so exp+rcp per element
// (2 MUFU) makes streaming SiLU kernels and conv epilogues MUFU-bound; this form needs 0.5 MUFU per
// element.  Absolute error <= ~|x| * 4e-4, well below the bf16 rounding of the stored result.
__device__ __forceinline__ void silu2(float& a, float& b) {
  const __half2 h = __floats2half2_rn(0.5f * a, 0.5f * b);
  uint32_t t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&h)));
  const float2 tf = __half22float2(*reinterpret_cast<const __half2*>(&t));
  a = 0.5f * a * (1.0f + tf.x);
  b = 0.5f * b * (1.0f + tf.y);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t v) {
  __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(t);
}
// 8 bf16 <-> 8 floats through one 16-byte vector
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]);
  v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
  return v;
}
// streaming 16-byte accesses (activations are touched once per kernel: keep them out of L1)
__device__ __forceinline__ uint4 ld_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_f4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

#endif  // __CUDACC__
}  // namespace srgd
