// Fused elementwise kernel of the discrete-time sampler family on the same U-Net (reference:
// ConditionalGaussianDiffusionSR, model.py:1311-1660): classifier-free-guidance combine, x_start from the
// objective (pred_noise / pred_x0 / pred_v), clamp, and either the DDPM posterior step (p_sample, model.py:1503-1514
// with q_posterior of the pip base class) or the DDIM step (ddim_sample, model.py:1599-1630).  One HBM-bound fp32
// NCHW pass; every op uses explicit round-to-nearest intrinsics in the reference's op order (no FMA contraction),
// like sampler.cu and edm.cu.  The per-step coefficients are table look-ups done by the host (srgd_b200/gaussian.py).
#include "common.cuh"

namespace srgd {

template <bool HAS_NULL, bool HAS_NOISE>
__global__ void __launch_bounds__(256) gauss_update_kernel(const float* __restrict__ xt, const float* __restrict__ out_c,
                                                           const float* __restrict__ out_n,
                                                           const float* __restrict__ noise, float* __restrict__ img,
                                                           float* __restrict__ x_start, float* __restrict__ pred_noise,
                                                           int64_t n, srgd_gauss_scalars s) {
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = xt[i];
    float o = out_c[i];
    if (HAS_NULL) {                                        // null + (cond - null) * scale     model.py:1462-1469
      const float nul = out_n[i];
      o = __fadd_rn(nul, __fmul_rn(__fsub_rn(o, nul), s.guidance_scale));
    }
    float x0;
    if (s.objective == SRGD_OBJ_PRED_NOISE)                // predict_start_from_noise          model.py:1474
      x0 = __fsub_rn(__fmul_rn(s.sqrt_recip_ac, x), __fmul_rn(s.sqrt_recipm1_ac, o));
    else if (s.objective == SRGD_OBJ_PRED_X0)              //                                   model.py:1481
      x0 = o;
    else                                                   // predict_start_from_v              model.py:1487
      x0 = __fsub_rn(__fmul_rn(s.sqrt_ac, x), __fmul_rn(s.sqrt_1m_ac, o));
    if (s.clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);        // maybe_clip / x_start.clamp_       model.py:1471, 1497
    if (x_start != nullptr) x_start[i] = x0;
    float pn = o;
    if (s.mode != SRGD_GAUSS_DDPM || pred_noise != nullptr) {
      // predict_noise_from_start; for pred_noise only when rederive_pred_noise asks for it after a clamp
      // (model.py:1477-1478), for the other objectives always (1483, 1489)
      if (s.objective != SRGD_OBJ_PRED_NOISE || (s.clip && s.rederive))
        pn = __fdiv_rn(__fsub_rn(__fmul_rn(s.sqrt_recip_ac, x), x0), s.sqrt_recipm1_ac);
      if (pred_noise != nullptr) pred_noise[i] = pn;
    }
    if (img == nullptr) continue;
    float r;
    if (s.mode == SRGD_GAUSS_DDPM) {
      // q_posterior mean + exp(0.5 log var) * noise                                   model.py:1499, 1512-1513
      r = __fadd_rn(__fmul_rn(s.coef1, x0), __fmul_rn(s.coef2, x));
      if (HAS_NOISE) r = __fadd_rn(r, __fmul_rn(s.noise_scale, noise[i]));
    } else if (s.mode == SRGD_GAUSS_DDIM_LAST) {
      r = x0;                                              // time_next < 0                     model.py:1604-1605
    } else {
      // x_start * sqrt(alpha_next) + c * pred_noise + sigma * noise                   model.py:1620-1622
      r = __fadd_rn(__fmul_rn(x0, s.sqrt_ac_next), __fmul_rn(s.c, pn));
      if (HAS_NOISE) r = __fadd_rn(r, __fmul_rn(s.noise_scale, noise[i]));
    }
    img[i] = r;
  }
}

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_gauss_update(const float* x_t, const float* out_cond, const float* out_null, const float* noise,
                                 float* img_out, float* x_start_out, float* pred_noise_out, int64_t n,
                                 const srgd_gauss_scalars* s, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x_t && out_cond && s && n > 0, "srgd_gauss_update: null argument or n <= 0");
  SRGD_REQUIRE(img_out || x_start_out || pred_noise_out, "srgd_gauss_update: no output requested");
  SRGD_REQUIRE(s->objective >= SRGD_OBJ_PRED_NOISE && s->objective <= SRGD_OBJ_PRED_V, "srgd_gauss_update: objective=%d",
               s->objective);
  SRGD_REQUIRE(s->mode >= SRGD_GAUSS_DDPM && s->mode <= SRGD_GAUSS_DDIM_LAST, "srgd_gauss_update: mode=%d", s->mode);
  cudaStream_t st = as_stream(stream);
  int64_t want = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(want < cap ? want : cap);
  ProfScope prof(SRGD_PK_SAMPLER, 0.0,
                 4.0 * (double)n * (2 + (out_null ? 1 : 0) + (noise ? 1 : 0) + (img_out ? 1 : 0) + (x_start_out ? 1 : 0) +
                                    (pred_noise_out ? 1 : 0)), st);
  cudaError_t e;
  if (out_null && noise) e = launch_k(gauss_update_kernel<true, true>, dim3(grid), dim3(256), 0, st, x_t, out_cond, out_null, noise, img_out, x_start_out, pred_noise_out, n, *s);
  else if (out_null) e = launch_k(gauss_update_kernel<true, false>, dim3(grid), dim3(256), 0, st, x_t, out_cond, out_null, noise, img_out, x_start_out, pred_noise_out, n, *s);
  else if (noise) e = launch_k(gauss_update_kernel<false, true>, dim3(grid), dim3(256), 0, st, x_t, out_cond, out_null, noise, img_out, x_start_out, pred_noise_out, n, *s);
  else e = launch_k(gauss_update_kernel<false, false>, dim3(grid), dim3(256), 0, st, x_t, out_cond, out_null, noise, img_out, x_start_out, pred_noise_out, n, *s);
  SRGD_CUDA_OK(e);
  count_launch();
  return SRGD_OK;
}
