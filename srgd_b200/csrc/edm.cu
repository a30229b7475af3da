// Fused elementwise kernels of the EDM sampler family on the same U-Net (reference:
// ConditionalElucidatedDiffusionSR, model.py:2059-2560): the stochastic perturbation of the state, the
// preconditioned-network combine + classifier-free guidance + clamp + Heun / Euler update, and the DPM-Solver++ (2M)
// update.  HBM-bound fp32 NCHW passes, one float4 per thread per iteration, grid-stride; every op is written with
// explicit round-to-nearest intrinsics in the reference's op order (no FMA contraction), like sampler.cu.
#include "common.cuh"

namespace srgd {

static int edm_grid(int64_t n4) {
  int64_t want = (n4 + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// images_hat = images + coef * (s_noise * noise)   (model.py:2270-2273);  x_in = c_in * images_hat (model.py:2141)
template <bool HAS_NOISE, bool HAS_XIN>
__global__ void __launch_bounds__(256) edm_perturb_kernel(const float* __restrict__ images,
                                                          const float* __restrict__ noise, float s_noise, float coef,
                                                          float c_in, float* hat, float* __restrict__ xin, int64_t n) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float h = images[i];
    if (HAS_NOISE) h = __fadd_rn(h, __fmul_rn(coef, __fmul_rn(s_noise, noise[i])));
    hat[i] = h;
    if (HAS_XIN) xin[i] = __fmul_rn(c_in, h);
  }
}

// denoised = clamp( null + (out - null) * s ),  out = c_skip x + c_out net      model.py:2148, 2160-2183
// d = (x - denoised) / sigma_eval                                              model.py:2279 / 2287
// images = x_base + step * (d_prev + d)                                        model.py:2281 / 2289
__global__ void __launch_bounds__(256) edm_update_kernel(const float* __restrict__ x_eval,
                                                         const float* __restrict__ net_c,
                                                         const float* __restrict__ net_n,
                                                         const float* __restrict__ x_base,
                                                         const float* __restrict__ d_prev, float* images_out,
                                                         float* d_out, float* den_out, float* xin_out, int64_t n,
                                                         srgd_edm_scalars s) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = x_eval[i];
    const float skip = __fmul_rn(s.c_skip, x);
    float den = __fadd_rn(skip, __fmul_rn(s.c_out, net_c[i]));
    if (net_n != nullptr) {
      const float nul = __fadd_rn(skip, __fmul_rn(s.c_out, net_n[i]));
      den = __fadd_rn(nul, __fmul_rn(__fsub_rn(den, nul), s.guidance_scale));
    }
    if (s.clip) den = fminf(fmaxf(den, -1.0f), 1.0f);
    if (den_out != nullptr) den_out[i] = den;
    if (images_out == nullptr && d_out == nullptr) continue;
    const float d = __fdiv_rn(__fsub_rn(x, den), s.sigma_eval);
    if (d_out != nullptr) d_out[i] = d;
    if (images_out != nullptr) {
      const float sum = d_prev != nullptr ? __fadd_rn(d_prev[i], d) : d;
      const float img = __fadd_rn(x_base[i], __fmul_rn(s.step, sum));
      images_out[i] = img;
      if (xin_out != nullptr) xin_out[i] = __fmul_rn(s.c_in_next, img);
    }
  }
}

// denoised_d = w_new * denoised + w_old * old;  images = a * images - b * denoised_d     model.py:2521-2530
__global__ void __launch_bounds__(256) edm_dpmpp_kernel(const float* __restrict__ images,
                                                        const float* __restrict__ den, const float* __restrict__ old,
                                                        float a, float b, float w_new, float w_old, float c_in_next,
                                                        float* images_out, float* xin_out, int64_t n) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float dd = den[i];
    if (old != nullptr) dd = __fadd_rn(__fmul_rn(w_new, dd), __fmul_rn(w_old, old[i]));
    const float img = __fsub_rn(__fmul_rn(a, images[i]), __fmul_rn(b, dd));
    images_out[i] = img;
    if (xin_out != nullptr) xin_out[i] = __fmul_rn(c_in_next, img);
  }
}

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_edm_perturb(const float* images, const float* noise, float s_noise, float coef, float c_in,
                                float* images_hat, float* x_in, int64_t n, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(images && images_hat && n > 0, "srgd_edm_perturb: null argument or n <= 0");
  cudaStream_t st = as_stream(stream);
  const int grid = edm_grid(n);
  ProfScope prof(SRGD_PK_SAMPLER, 0.0, 4.0 * (double)n * (2 + (noise ? 1 : 0) + (x_in ? 1 : 0)), st);
  cudaError_t e;
  if (noise && x_in) e = launch_k(edm_perturb_kernel<true, true>, dim3(grid), dim3(256), 0, st, images, noise, s_noise, coef, c_in, images_hat, x_in, n);
  else if (noise) e = launch_k(edm_perturb_kernel<true, false>, dim3(grid), dim3(256), 0, st, images, noise, s_noise, coef, c_in, images_hat, x_in, n);
  else if (x_in) e = launch_k(edm_perturb_kernel<false, true>, dim3(grid), dim3(256), 0, st, images, noise, s_noise, coef, c_in, images_hat, x_in, n);
  else e = launch_k(edm_perturb_kernel<false, false>, dim3(grid), dim3(256), 0, st, images, noise, s_noise, coef, c_in, images_hat, x_in, n);
  SRGD_CUDA_OK(e);
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_edm_update(const float* x_eval, const float* net_cond, const float* net_null, const float* x_base,
                               const float* d_prev, float* images_out, float* d_out, float* denoised_out,
                               float* x_in_out, int64_t n, const srgd_edm_scalars* s, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x_eval && net_cond && s && n > 0, "srgd_edm_update: null argument or n <= 0");
  SRGD_REQUIRE(images_out == nullptr || x_base != nullptr, "srgd_edm_update: images_out needs x_base");
  SRGD_REQUIRE((images_out == nullptr && d_out == nullptr) || s->sigma_eval > 0.f, "srgd_edm_update: sigma_eval must be > 0");
  cudaStream_t st = as_stream(stream);
  ProfScope prof(SRGD_PK_SAMPLER, 0.0,
                 4.0 * (double)n * (2 + (net_null ? 1 : 0) + (x_base ? 1 : 0) + (d_prev ? 1 : 0) + (images_out ? 1 : 0) +
                                    (d_out ? 1 : 0) + (denoised_out ? 1 : 0) + (x_in_out ? 1 : 0)), st);
  SRGD_CUDA_OK(launch_k(edm_update_kernel, dim3(edm_grid(n)), dim3(256), 0, st, x_eval, net_cond, net_null, x_base, d_prev,
                        images_out, d_out, denoised_out, x_in_out, n, *s));
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_edm_dpmpp(const float* images, const float* denoised, const float* old_denoised, float a, float b,
                              float w_new, float w_old, float c_in_next, float* images_out, float* x_in_out, int64_t n,
                              srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(images && denoised && images_out && n > 0, "srgd_edm_dpmpp: null argument or n <= 0");
  cudaStream_t st = as_stream(stream);
  ProfScope prof(SRGD_PK_SAMPLER, 0.0, 4.0 * (double)n * (3 + (old_denoised ? 1 : 0) + (x_in_out ? 1 : 0)), st);
  SRGD_CUDA_OK(launch_k(edm_dpmpp_kernel, dim3(edm_grid(n)), dim3(256), 0, st, images, denoised, old_denoised, a, b, w_new,
                        w_old, c_in_next, images_out, x_in_out, n));
  count_launch();
  return SRGD_OK;
}
