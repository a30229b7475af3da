// Fused sampler-step kernels: CFG combine + x0 + clamp + posterior mean + noise add
// (reference: model.py:3150/3154, 3160-3168, 3187-3188), q_sample (3434-3447), final clamp (3237).
//
// HBM-bound elementwise work on the fp32 NCHW sampler state.  Algorithmic bytes per element:
// read x(4) + eps_cond(4) [+ eps_null(4)] [+ noise(4)], write img_next(4) [+ x_start(4)]
// = 20 B (no CFG) / 24 B (CFG) with both optional outputs.
// One float4 per thread per iteration, grid-stride, grid = multiple of the SM count.
#include "common.cuh"

namespace srgd {

// One element of the update.  Written with explicit round-to-nearest intrinsics so that nvcc
// does not contract a*b+c into FMA: the reference evaluates each torch op separately in fp32, and
// the result here is bit-identical to that sequence.
template <bool HAS_NULL, bool HAS_NOISE>
__device__ __forceinline__ void step_elem(float x, float ec, float en, float z, const srgd_step_scalars& s,
                                          float& out, float& x0) {
  // eps = null + (cond - null) * s                          model.py:3150 / 3154
  float eps = HAS_NULL ? __fadd_rn(en, __fmul_rn(__fsub_rn(ec, en), s.guidance_scale)) : ec;
  // x_start = (x - sigma * eps) / alpha                     model.py:3160
  float xs = __fdiv_rn(__fsub_rn(x, __fmul_rn(s.sigma, eps)), s.alpha);
  float mean;
  if (s.clip) {
    xs = fminf(fmaxf(xs, -1.0f), 1.0f);                                       // model.py:3163
    // alpha_next * (x * (1 - c) / alpha + c * x_start)     model.py:3164
    float a = __fdiv_rn(__fmul_rn(x, __fsub_rn(1.0f, s.c)), s.alpha);
    mean = __fmul_rn(s.alpha_next, __fadd_rn(a, __fmul_rn(s.c, xs)));
  } else {
    // alpha_next / alpha * (x - c * sigma * eps)            model.py:3166
    mean = __fmul_rn(__fdiv_rn(s.alpha_next, s.alpha),
                     __fsub_rn(x, __fmul_rn(__fmul_rn(s.c, s.sigma), eps)));
  }
  out = HAS_NOISE ? __fadd_rn(mean, __fmul_rn(s.noise_scale, z)) : mean;      // model.py:3188
  x0 = xs;
}

template <bool HAS_NULL, bool HAS_NOISE, bool HAS_X0>
__global__ void __launch_bounds__(256) sampler_step_kernel(
    const float* __restrict__ x, const float* __restrict__ ec, const float* __restrict__ en,
    const float* __restrict__ noise, float* img, float* __restrict__ x0out, int64_t n4, int64_t n,
    srgd_step_scalars s) {
  pdl_wait();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 xv = ld_stream_f4(x + 4 * i);
    float4 cv = ld_stream_f4(ec + 4 * i);
    float4 nv = make_float4(0, 0, 0, 0), zv = make_float4(0, 0, 0, 0);
    if (HAS_NULL) nv = ld_stream_f4(en + 4 * i);
    if (HAS_NOISE) zv = ld_stream_f4(noise + 4 * i);
    float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ca[4] = {cv.x, cv.y, cv.z, cv.w};
    float na[4] = {nv.x, nv.y, nv.z, nv.w}, za[4] = {zv.x, zv.y, zv.z, zv.w};
    float oa[4], sa[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) step_elem<HAS_NULL, HAS_NOISE>(xa[j], ca[j], na[j], za[j], s, oa[j], sa[j]);
    st_stream_f4(img + 4 * i, make_float4(oa[0], oa[1], oa[2], oa[3]));
    if (HAS_X0) st_stream_f4(x0out + 4 * i, make_float4(sa[0], sa[1], sa[2], sa[3]));
  }
  // scalar tail (n not a multiple of 4)
  if (blockIdx.x == 0) {
    for (int64_t i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) {
      float o, xs;
      step_elem<HAS_NULL, HAS_NOISE>(x[i], ec[i], HAS_NULL ? en[i] : 0.f, HAS_NOISE ? noise[i] : 0.f, s, o, xs);
      img[i] = o;
      if (HAS_X0) x0out[i] = xs;
    }
  }
}

__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0,
                                                       const float* __restrict__ noise, float* out,
                                                       int64_t n, float alpha, float sigma) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = noise[i] * sigma;
    if (x0) v = x0[i] * alpha + v;                                            // model.py:3442
    out[i] = v;
  }
}

__global__ void __launch_bounds__(256) finalize_kernel(const float* img, float* out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = fminf(fmaxf(img[i], -1.0f), 1.0f);                              // model.py:3237
    out[i] = (v + 1.0f) * 0.5f;                                               // model.py:44
  }
}

// ---- tiled sampling orchestration (model.py:3361-3396): batched tile gather / scatter, re-noise outside a hull ----
// canvas fp32 [C][H][W] (batch 1, like the reference's tiled_sample); tiles fp32 [n][C][T][T]; one float4 per thread.
template <bool SCATTER>
__global__ void __launch_bounds__(256) tile_copy_kernel(float* canvas, float* tiles, srgd_tile_coords tc, int C, int H,
                                                        int W, int T) {
  const int t4 = T / 4;
  const int64_t total = (int64_t)tc.n * C * T * t4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x4 = (int)(i % t4);
    const int y = (int)((i / t4) % T);
    const int c = (int)((i / ((int64_t)t4 * T)) % C);
    const int k = (int)(i / ((int64_t)t4 * T * C));
    float* cp = canvas + ((int64_t)c * H + tc.yx[k][0] + y) * W + tc.yx[k][1] + x4 * 4;
    float* tp = tiles + i * 4;
    if (SCATTER) st_stream_f4(cp, ld_stream_f4(tp));
    else st_stream_f4(tp, ld_stream_f4(cp));
  }
}

// img[c][y][x] = sigma * noise[c][y][x] outside the rectangle [y0,y1) x [x0,x1), unchanged inside (model.py:3392-3396:
// q_sample of zeros at the next noise level everywhere except the hull of the shifted tile grid)
__global__ void __launch_bounds__(256) renoise_outside_kernel(float* img, const float* __restrict__ noise, int64_t n,
                                                              int H, int W, int y0, int y1, int x0, int x1,
                                                              float sigma) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    if (y < y0 || y >= y1 || x < x0 || x >= x1) img[i] = noise[i] * sigma;
  }
}

static int grid_for(int64_t work_items, int threads, int ctas_per_sm) {
  int64_t want = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_sampler_step(const float* x, const float* eps_cond, const float* eps_null,
                                 const float* noise, float* img_next, float* x_start, int64_t n,
                                 const srgd_step_scalars* s, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x && eps_cond && img_next && s && n > 0, "srgd_sampler_step: null argument or n <= 0");
  SRGD_REQUIRE(((uintptr_t)x | (uintptr_t)eps_cond | (uintptr_t)eps_null | (uintptr_t)noise |
                (uintptr_t)img_next | (uintptr_t)x_start) % 16 == 0,
               "srgd_sampler_step: pointers must be 16-byte aligned");
  SRGD_REQUIRE(s->alpha > 0.f, "srgd_sampler_step: alpha must be > 0");
  const int64_t n4 = n / 4;
  const int grid = grid_for(n4, 256, 8);
  cudaStream_t st = as_stream(stream);
  const int key = (eps_null ? 4 : 0) | (noise ? 2 : 0) | (x_start ? 1 : 0);
  ProfScope prof(SRGD_PK_SAMPLER, 0.0,
                 4.0 * (double)n * (3 + (eps_null ? 1 : 0) + (noise ? 1 : 0) + (x_start ? 1 : 0)), st);
  cudaError_t launch_err = cudaSuccess;
#define SRGD_CASE(K, A, B_, C)                                                                    \
  case K:                                                                                         \
    launch_err = launch_k(sampler_step_kernel<A, B_, C>, dim3(grid), dim3(256), 0, st, x, eps_cond, \
                          eps_null, noise, img_next, x_start, n4, n, *s);                         \
    break;
  switch (key) {
    SRGD_CASE(0, false, false, false)
    SRGD_CASE(1, false, false, true)
    SRGD_CASE(2, false, true, false)
    SRGD_CASE(3, false, true, true)
    SRGD_CASE(4, true, false, false)
    SRGD_CASE(5, true, false, true)
    SRGD_CASE(6, true, true, false)
    SRGD_CASE(7, true, true, true)
  }
#undef SRGD_CASE
  SRGD_CUDA_OK(launch_err);
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_q_sample(const float* x_start, const float* noise, float* out, int64_t n,
                             float alpha, float sigma, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(noise && out && n > 0, "srgd_q_sample: null argument or n <= 0");
  ProfScope prof(SRGD_PK_OTHER, 0.0, 4.0 * (double)n * (x_start ? 3 : 2), as_stream(stream));
  q_sample_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(x_start, noise, out, n, alpha, sigma);
  SRGD_LAUNCH_OK("q_sample_kernel");
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_finalize_image(const float* img, float* out, int64_t n, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(img && out && n > 0, "srgd_finalize_image: null argument or n <= 0");
  ProfScope prof(SRGD_PK_OTHER, 0.0, 8.0 * (double)n, as_stream(stream));
  finalize_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(img, out, n);
  SRGD_LAUNCH_OK("finalize_kernel");
  count_launch();
  return SRGD_OK;
}

static int tile_copy(bool scatter, float* canvas, float* tiles, const srgd_tile_coords* tc, int32_t C, int32_t H,
                     int32_t W, int32_t T, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(canvas && tiles && tc, "tile gather/scatter: null argument");
  SRGD_REQUIRE(tc->n >= 1 && tc->n <= SRGD_MAX_TILES_PER_CALL, "tile gather/scatter: n=%d out of range", tc->n);
  SRGD_REQUIRE(C > 0 && T > 0 && T % 4 == 0 && W % 4 == 0 && H >= T && W >= T, "tile gather/scatter: bad geometry");
  SRGD_REQUIRE(((uintptr_t)canvas | (uintptr_t)tiles) % 16 == 0, "tile gather/scatter: pointers must be 16-byte aligned");
  for (int k = 0; k < tc->n; ++k)
    SRGD_REQUIRE(tc->yx[k][0] >= 0 && tc->yx[k][0] + T <= H && tc->yx[k][1] >= 0 && tc->yx[k][1] + T <= W &&
                     tc->yx[k][1] % 4 == 0,
                 "tile gather/scatter: tile %d at (%d,%d) outside the %dx%d canvas or x not a multiple of 4", k,
                 tc->yx[k][0], tc->yx[k][1], H, W);
  const int64_t total = (int64_t)tc->n * C * T * (T / 4);
  ProfScope prof(SRGD_PK_OTHER, 0.0, 32.0 * (double)total, as_stream(stream));
  if (scatter) tile_copy_kernel<true><<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(canvas, tiles, *tc, C, H, W, T);
  else tile_copy_kernel<false><<<grid_for(total, 256, 8), 256, 0, as_stream(stream)>>>(canvas, tiles, *tc, C, H, W, T);
  SRGD_LAUNCH_OK("tile_copy_kernel");
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_gather_tiles(const float* canvas, float* tiles, const srgd_tile_coords* tc, int32_t C, int32_t H,
                                 int32_t W, int32_t T, srgd_stream_t stream) {
  return tile_copy(false, const_cast<float*>(canvas), tiles, tc, C, H, W, T, stream);
}

extern "C" int srgd_scatter_tiles(float* canvas, const float* tiles, const srgd_tile_coords* tc, int32_t C, int32_t H,
                                  int32_t W, int32_t T, srgd_stream_t stream) {
  return tile_copy(true, canvas, const_cast<float*>(tiles), tc, C, H, W, T, stream);
}

extern "C" int srgd_renoise_outside(float* img, const float* noise, int32_t C, int32_t H, int32_t W, int32_t y0,
                                    int32_t y1, int32_t x0, int32_t x1, float sigma, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(img && noise && C > 0 && H > 0 && W > 0, "renoise_outside: bad arguments");
  const int64_t n = (int64_t)C * H * W;
  ProfScope prof(SRGD_PK_OTHER, 0.0, 8.0 * (double)n, as_stream(stream));
  renoise_outside_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(img, noise, n, H, W, y0, y1, x0, x1, sigma);
  SRGD_LAUNCH_OK("renoise_outside_kernel");
  count_launch();
  return SRGD_OK;
}
