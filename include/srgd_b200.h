/*
 * srgd_b200 -- C-ABI of the B200-native (sm_100a) Real-SRGD sampling hot path.
 *
 * The reference (yahoojapan/srgd) has no FFI: its hot path is Python/PyTorch
 * (model.py).  Each entry point below replaces the reference code it cites
 * (paths relative to the reference checkout); the Python host in `srgd_b200/`
 * binds them with ctypes (see INTEGRATION.md for the stub a reference
 * maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success or a negative SRGD_E_* code; the text
 *     of the last failure on the calling thread is srgd_last_error();
 *   - pointers named *_dev / documented "device" are CUDA device pointers owned
 *     by the caller (torch tensors via .data_ptr()); nothing is allocated,
 *     freed or synchronised inside (workspaces are passed in);
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - activations between kernels are bf16 NHWC ("pixel-major": [B][H][W][C]);
 *     sampler state, eps, noise and images are fp32 NCHW exactly like the
 *     reference tensors;
 *   - there is NO CPU fallback: without an sm_100 device every compute entry
 *     point fails with SRGD_E_DEVICE.
 */
#ifndef SRGD_B200_H
#define SRGD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRGD_B200_VERSION 100

enum {
  SRGD_OK = 0,
  SRGD_E_ARG = -1,      /* invalid argument / unsupported shape                   */
  SRGD_E_DEVICE = -2,   /* no CUDA device, or device is not sm_100                */
  SRGD_E_CUDA = -3,     /* a CUDA runtime/driver call failed (see last_error)     */
  SRGD_E_WORKSPACE = -4 /* workspace too small                                    */
};

typedef void* srgd_stream_t;

int srgd_version(void);
const char* srgd_last_error(void);
/* 0 if `device` is a usable sm_100 GPU, else SRGD_E_DEVICE. */
int srgd_device_check(int device);

/* ------------------------------------------------------------------------------------------
 * Sampler (model.py:3122-3188).  Scalars are the fp32 values of model.py:3127-3134, computed by
 * the host exactly like the reference (0-dim fp32 tensor math) and passed by value.
 * ------------------------------------------------------------------------------------------ */
typedef struct srgd_step_scalars {
  float alpha;          /* sqrt(sigmoid(log_snr))                         model.py:3131,3134 */
  float sigma;          /* sqrt(sigmoid(-log_snr))                        model.py:3132,3134 */
  float alpha_next;     /* sqrt(sigmoid(log_snr_next))                                        */
  float c;              /* -expm1(log_snr - log_snr_next)                 model.py:3129      */
  float noise_scale;    /* sqrt(sigmoid(-log_snr_next) * c); 0 on the last step (3168,3184)  */
  float guidance_scale; /* s in  eps = null + s (cond - null)             model.py:3150,3154 */
  int32_t clip;         /* clip_sample_denoised                           model.py:3162      */
} srgd_step_scalars;

/* Fused CFG combine + x0 + clamp + posterior mean + noise add (model.py:3150/3154, 3160-3168,
 * 3187-3188) over n fp32 elements.  eps_null, noise and x_start may be NULL.  img_next may alias x. */
int srgd_sampler_step(const float* x, const float* eps_cond, const float* eps_null,
                      const float* noise, float* img_next, float* x_start, int64_t n,
                      const srgd_step_scalars* s, srgd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * EDM sampler family on the same U-Net (ConditionalElucidatedDiffusionSR, model.py:2059-2560; SURVEY.md section 8
 * f-4).  Scalars are computed by the host like the reference does (python doubles / fp32 tensor math) and passed
 * as fp32.  All buffers fp32, n elements; optional pointers may be NULL.
 * ------------------------------------------------------------------------------------------ */
/* images_hat = images + coef * (s_noise * noise)  (model.py:2270-2273; noise NULL: images_hat = images) and
 * x_in = c_in * images_hat, the scaled input of the next U-Net evaluation (model.py:2141). */
int srgd_edm_perturb(const float* images, const float* noise, float s_noise, float coef, float c_in,
                     float* images_hat, float* x_in, int64_t n, srgd_stream_t stream);

typedef struct srgd_edm_scalars {
  float c_skip, c_out;      /* preconditioning at sigma_eval                              model.py:2147      */
  float guidance_scale;     /* s in  out = null + (out - null) * s                        model.py:2165, 2178 */
  float sigma_eval;         /* d = (x_eval - denoised) / sigma_eval                       model.py:2279, 2287 */
  float step;               /* images = x_base + step * (d_prev + d): sigma_next - sigma_hat for the Euler stage,
                               0.5 * (sigma_next - sigma_hat) for the Heun correction     model.py:2281, 2289 */
  float c_in_next;          /* x_in_out = c_in_next * images                                                  */
  int32_t clip;             /* clamp the denoised image to [-1, 1]                        model.py:2180      */
} srgd_edm_scalars;
/* preconditioned_network_forward's combine (c_skip x + c_out net, classifier-free guidance against net_null, clamp;
 * model.py:2148-2183) fused with the sampler update that consumes it.  Outputs (each optional): images_out (needs
 * x_base; d_prev optional), d_out = (x_eval - denoised) / sigma_eval, denoised_out, x_in_out = c_in_next * images_out. */
int srgd_edm_update(const float* x_eval, const float* net_cond, const float* net_null, const float* x_base,
                    const float* d_prev, float* images_out, float* d_out, float* denoised_out, float* x_in_out,
                    int64_t n, const srgd_edm_scalars* s, srgd_stream_t stream);
/* DPM-Solver++ (2M) update (model.py:2521-2530): denoised_d = w_new * denoised + w_old * old_denoised (old NULL:
 * denoised_d = denoised); images_out = a * images - b * denoised_d; x_in_out = c_in_next * images_out. */
int srgd_edm_dpmpp(const float* images, const float* denoised, const float* old_denoised, float a, float b,
                   float w_new, float w_old, float c_in_next, float* images_out, float* x_in_out, int64_t n,
                   srgd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Discrete-time sampler family on the same U-Net (ConditionalGaussianDiffusionSR, model.py:1311-1660: DDPM
 * ancestral sampling and DDIM).  The per-timestep coefficients are entries of the
 * reference's registered fp32 buffers (model.py:1395-1424), looked up by the host and passed as scalars.
 * ------------------------------------------------------------------------------------------ */
enum { SRGD_OBJ_PRED_NOISE = 0, SRGD_OBJ_PRED_X0 = 1, SRGD_OBJ_PRED_V = 2 };          /* model.py:1472-1489 */
enum { SRGD_GAUSS_DDPM = 0,        /* p_sample: posterior mean + exp(0.5 log var) * noise    model.py:1503-1514 */
       SRGD_GAUSS_DDIM = 1,        /* ddim_sample step with time_next >= 0                   model.py:1608-1622 */
       SRGD_GAUSS_DDIM_LAST = 2 }; /* ddim_sample step with time_next < 0: img = x_start     model.py:1604-1605 */
typedef struct srgd_gauss_scalars {
  int32_t objective, mode;
  int32_t clip;             /* clamp x_start to [-1, 1] (clip_x_start, or p_mean_variance's clamp_)  1471, 1497 */
  int32_t rederive;         /* rederive_pred_noise (only read for objective pred_noise)              1477-1478 */
  float guidance_scale;     /* s in  out = null + (cond - null) * s                                  1464, 1468 */
  float sqrt_recip_ac, sqrt_recipm1_ac;   /* sqrt_recip_alphas_cumprod[t], sqrt_recipm1_alphas_cumprod[t]        */
  float sqrt_ac, sqrt_1m_ac;              /* sqrt_alphas_cumprod[t], sqrt_one_minus_alphas_cumprod[t] (pred_v)   */
  float coef1, coef2;                     /* posterior_mean_coef1[t], posterior_mean_coef2[t]        (DDPM)      */
  float noise_scale;                      /* DDPM: exp(0.5 * posterior_log_variance_clipped[t]); DDIM: sigma    */
  float sqrt_ac_next, c;                  /* DDIM: sqrt(alphas_cumprod[t_next]), sqrt(1 - alpha_next - sigma^2) */
} srgd_gauss_scalars;
/* model_predictions (guidance combine, x_start, clamp, pred_noise; model.py:1449-1489) fused with the update that
 * consumes it.  out_null / noise may be NULL (no guidance / no noise term); each of img_out, x_start_out,
 * pred_noise_out is optional but at least one is required.  pred_noise_out is the noise AFTER the clamp when the
 * objective or `rederive` derives it from x_start. */
int srgd_gauss_update(const float* x_t, const float* out_cond, const float* out_null, const float* noise,
                      float* img_out, float* x_start_out, float* pred_noise_out, int64_t n,
                      const srgd_gauss_scalars* s, srgd_stream_t stream);

/* q_sample (model.py:3434-3447): out = x_start*alpha + noise*sigma.  x_start may be NULL (zeros). */
int srgd_q_sample(const float* x_start, const float* noise, float* out, int64_t n, float alpha,
                  float sigma, srgd_stream_t stream);

/* final clamp(-1,1) and (x+1)/2 (model.py:3237-3238, 3404-3405). out may alias img. */
int srgd_finalize_image(const float* img, float* out, int64_t n, srgd_stream_t stream);

/* Tiled sampling orchestration (tiled_sample, model.py:3361-3396) on a fp32 [C][H][W] canvas (batch 1):
 * gather up to SRGD_MAX_TILES_PER_CALL T x T tiles into a [n][C][T][T] minibatch (model.py:3364-3371),
 * scatter results back (3377-3380), and replace the state outside the hull [y0,y1) x [x0,x1) of the
 * shifted grid by sigma * noise (q_sample of zeros, 3392-3396).  Tile x offsets must be multiples of 4. */
enum { SRGD_MAX_TILES_PER_CALL = 64 };
typedef struct srgd_tile_coords {
  int32_t n;
  int32_t yx[SRGD_MAX_TILES_PER_CALL][2];   /* top-left (y, x) of each tile on the canvas */
} srgd_tile_coords;
int srgd_gather_tiles(const float* canvas, float* tiles, const srgd_tile_coords* tc, int32_t C, int32_t H,
                      int32_t W, int32_t T, srgd_stream_t stream);
int srgd_scatter_tiles(float* canvas, const float* tiles, const srgd_tile_coords* tc, int32_t C, int32_t H,
                       int32_t W, int32_t T, srgd_stream_t stream);
int srgd_renoise_outside(float* img, const float* noise, int32_t C, int32_t H, int32_t W, int32_t y0,
                         int32_t y1, int32_t x0, int32_t x1, float sigma, srgd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Convolution as implicit GEMM on tcgen05 tensor cores (nn.Conv2d call sites model.py:246, 271,
 * 300, 303, 341, 342, 109, 78, 647, 668, 583).  D[M=B*Ho*Wo, N=Cout] = sum over taps and input
 * channel blocks of A(pixel shifted by tap)[M, 64] * W[N, 64]^T, bf16 operands, fp32 TMEM
 * accumulators.
 * ------------------------------------------------------------------------------------------ */
enum { SRGD_CONV_MAX_SRC = 4, SRGD_CONV_MAX_PHASE = 20 };

enum {                     /* srgd_conv_desc.out_mode */
  SRGD_OUT_BF16_NHWC = 0,  /* out[b][y][x][n]                                                 */
  SRGD_OUT_PIXEL_SHUFFLE = 1 /* PixelShuffle(2) of the result (model.py:83): weights are packed
                                so that n = (i*2+j)*C' + c'; out[b][2y+i][2x+j][c'], C' = Cout/4 */
};

typedef struct srgd_conv_src {
  const void* ptr;         /* bf16, element (b,y,x,c) at ((b*sb + y*sy + x*sx) + c) elements   */
  int64_t sb, sy, sx;      /* element strides (multiples of 8)                                 */
  int32_t H, W, C;         /* extents seen by the kernel (C multiple of 64)                    */
} srgd_conv_src;

typedef struct srgd_conv_phase {
  int32_t src;             /* index into srcs[]                                                */
  int32_t dy, dx;          /* input pixel = output pixel + (dy,dx); out of range reads zero    */
  int32_t k_start;         /* first weight column (multiple of 64); covers srcs[src].C columns */
} srgd_conv_phase;

typedef struct srgd_conv_desc {
  int32_t B, Ho, Wo, Cout; /* Cout multiple of 64                                              */
  int32_t n_src, n_phase;
  srgd_conv_src srcs[SRGD_CONV_MAX_SRC];
  srgd_conv_phase phases[SRGD_CONV_MAX_PHASE];
  const void* weight;      /* bf16 [Cout][Ktot], K-major                                       */
  int64_t Ktot;
  const float* bias;       /* [Cout] or NULL                                                   */
  const float* row_scale;  /* per output pixel [B*Ho*Wo] fp32 or NULL: acc *= row_scale before bias
                              (RMSNorm folded into the following 1x1 conv, model.py:207,310,347) */
  const void* residual;    /* bf16, same layout as out, added last; or NULL                    */
  int32_t act;             /* 0 none, 1 SiLU (applied after bias, before residual)             */
  int32_t out_mode;
  void* out;               /* bf16                                                             */
  float* gn_partials;      /* NULL, or fp32 [m_tiles][S][8][2], S = samples spanned by one 128-pixel M tile
                              (1 when Ho*Wo >= 128): sum and sum of squares per GroupNorm group of the fp32
                              result (model.py:247) over the pixels of that (M tile, sample)               */
  void* splitk_ws;         /* NULL, or a device workspace of srgd_conv_splitk_workspace_bytes() bytes whose first
                              4096 bytes were zeroed once (the kernel re-arms them): when the tiles of a launch leave
                              the last wave of the persistent grid at most half full, the K loop of those tiles is
                              split across the idle SMs (fp32 partial accumulators through this workspace, summed in
                              a fixed order: deterministic).  One workspace serves any number of launches on a stream */
  int64_t splitk_ws_bytes;
} srgd_conv_desc;

/* Size of srgd_conv_desc.splitk_ws. */
size_t srgd_conv_splitk_workspace_bytes(void);

/* Number of 64-byte gn_partials records the kernel writes for (B,Ho,Wo): m_tiles * S. */
int srgd_conv_m_tiles(int32_t B, int32_t Ho, int32_t Wo);
int srgd_conv_igemm(const srgd_conv_desc* d, srgd_stream_t stream);
/* Same contract on CUDA cores, one thread per output element.  Debug/verification path used by
 * the GPU tests to cross-check the tensor-core kernel; never selected by the product path. */
int srgd_conv_direct(const srgd_conv_desc* d, srgd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm(8 groups, eps 1e-5) + (scale+1, shift) + SiLU (+ residual)   model.py:247-258, 285
 * ------------------------------------------------------------------------------------------ */
/* Reduce the conv epilogue partials to mean / rstd per (sample, group): stats[B][8][2].  The product path does
 * not launch this: srgd_groupnorm_apply* fold the partials themselves when given gn_partials instead of stats. */
int srgd_groupnorm_finalize(const float* gn_partials, float* stats, int32_t B, int32_t H, int32_t W,
                            int32_t C, srgd_stream_t stream);
/* Stand-alone statistics pass over a bf16 NHWC tensor (used when the producer was not a conv). */
int srgd_groupnorm_stats(const void* x, float* stats, int32_t B, int32_t H, int32_t W, int32_t C,
                         srgd_stream_t stream);
/* y = SiLU( (GN(x)*gamma+beta) * (scale+1) + shift ) [+ residual].  scale_shift: fp32, row b at
 * scale_shift + b*ss_stride holds [scale(C) | shift(C)] (model.py:279), or NULL.  Row b of the
 * output reads sample (b % Bx) of x and stats (CFG halves sharing one conv result).  y may alias
 * x when Bx == B.  Statistics: exactly one of `stats` ([Bx][8][2] mean, rstd) and `gn_partials` (the records
 * srgd_conv_igemm wrote for x with the same (Bx,H,W); every block folds its sample's records in fp64 in a
 * fixed order, which replaces the srgd_groupnorm_finalize launch) is non-NULL.  inv_out (optional, residual variant with C in {128,256} only): fp32 [B*H*W],
 * receives 1 / max(||y[pixel,:]||_2, 1e-12) of the stored bf16 row -- the RMSNorm statistic of the
 * attention block that consumes y (model.py:207), saving srgd_pixel_inv_norm's extra pass. */
int srgd_groupnorm_apply(const void* x, int32_t Bx, const float* stats, const float* gn_partials,
                         const float* gamma, const float* beta, const float* scale_shift, int64_t ss_stride,
                         const void* residual, void* y, float* inv_out, int32_t B, int32_t H, int32_t W,
                         int32_t C, srgd_stream_t stream);

/* The same with a broadcast residual: row b adds row (b % res_rows) of `residual` (B a multiple of res_rows).  Used
 * when the two halves of a class-guidance batch share init_conv's output as the identity residual of the first
 * ResnetBlock (model.py:3151-3154 run the same x twice; here it is computed once). */
int srgd_groupnorm_apply_ex(const void* x, int32_t Bx, const float* stats, const float* gn_partials,
                            const float* gamma, const float* beta, const float* scale_shift, int64_t ss_stride,
                            const void* residual, int32_t res_rows, void* y, float* inv_out, int32_t B, int32_t H,
                            int32_t W, int32_t C, srgd_stream_t stream);

/* Last ResnetBlock of the network fused with the final 1x1 conv (model.py:674-675, 724-725):
 * eps[b][o][y][x] = final_b[o] + sum_c final_w[o][c] * ( SiLU(GN(x)*gamma+beta) + residual )[b][y][x][c],
 * fp32 NCHW out, o < 3; the normalised activation itself is never stored.  C must be 128. */
int srgd_groupnorm_apply_final(const void* x, const float* stats, const float* gn_partials,
                               const float* gamma, const float* beta,
                               const void* residual, const float* final_w, const float* final_b,
                               float* eps, int32_t B, int32_t H, int32_t W, int32_t C,
                               srgd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * RMSNorm pieces (model.py:201-207)
 * ------------------------------------------------------------------------------------------ */
/* inv[m] = 1 / max(||x[m,:]||_2, 1e-12) for each of M pixels of a bf16 [M][C] tensor. */
int srgd_pixel_inv_norm(const void* x, float* inv, int64_t M, int32_t C, srgd_stream_t stream);
/* y = x / max(||x||,1e-12) * g * sqrt(C) [+ residual]   (to_out.1 + caller's "+ x", 304, 703) */
int srgd_rmsnorm_residual(const void* x, const float* g, const void* residual, void* y, int64_t M,
                          int32_t C, srgd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Attention cores on the packed qkv tensor bf16 [B][N][3*heads*32] (q | k | v, head-major)
 * ------------------------------------------------------------------------------------------ */
/* LinearAttention core, model.py:315-323: softmax(q) over d, softmax(k) over N, ctx = k v^T,
 * out = ctx^T q * 32^-1/2 -> bf16 [B][N][heads*32].  workspace: srgd_linear_attention_workspace(). */
size_t srgd_linear_attention_workspace(int32_t B, int32_t N, int32_t heads);
int srgd_linear_attention(const void* qkv, void* out, int32_t B, int32_t N, int32_t heads,
                          void* workspace, size_t workspace_bytes, srgd_stream_t stream);
/* Whole LinearAttention block fused on tcgen05 (model.py:307-324 + the caller's "+ x", 703/718):
 * out = RMSNorm_g(to_out(linear_attention(to_qkv(RMSNorm(x))))) + x, bf16 [B][N][C] in and out.
 * inv_norm: fp32 [B*N] = 1/max(||x[pixel,:]||,1e-12) if the producer already has it (srgd_groupnorm_apply
 * inv_out), or NULL to compute it here.
 * qkv_w: bf16 [3*heads*32][C] with the pre-norm gain g*sqrt(C) folded into its columns; out_w: bf16
 * [C][heads*32]; out_b, out_g: fp32 [C].  q/k/v/o never leave the SM (TMEM + shared memory).
 * Supported shapes only (srgd_linear_attention_block_supported: heads=4, C in {128,256}, N % 128 == 0);
 * other shapes use srgd_pixel_inv_norm + srgd_conv_igemm + srgd_linear_attention + srgd_rmsnorm_residual. */
int srgd_linear_attention_block_supported(int32_t N, int32_t C, int32_t heads);
size_t srgd_linear_attention_block_workspace(int32_t B, int32_t N, int32_t C, int32_t heads);
int srgd_linear_attention_block(const void* x, const float* inv_norm, const void* qkv_w, const void* out_w,
                                const float* out_b, const float* out_g, void* out, int32_t B, int32_t N,
                                int32_t C, int32_t heads, void* workspace, size_t workspace_bytes,
                                srgd_stream_t stream);
/* Full attention core (Attend, model.py:352): softmax(q k^T * 32^-1/2) v -> bf16 [B][N][heads*32]. */
int srgd_attention(const void* qkv, void* out, int32_t B, int32_t N, int32_t heads,
                   srgd_stream_t stream);

/* Same contract as srgd_attention on the tcgen05 tensor cores (flash-style: S = Q K^T and P V are
 * tcgen05.mma with TMEM accumulators, online softmax in fp32).  Needs N % 128 == 0, heads <= 4
 * (srgd_attention_tc_supported); the U-Net launcher falls back to srgd_attention otherwise. */
int srgd_attention_tc_supported(int32_t N, int32_t heads);
int srgd_attention_tc(const void* qkv, void* out, int32_t B, int32_t N, int32_t heads,
                      srgd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * U-Net entry and exit (model.py:681-687, 722-725) and embeddings (model.py:223-238, 603-619, 264-267)
 * ------------------------------------------------------------------------------------------ */
/* cat(x, cond) fp32 NCHW -> bf16 [B][H][W][64] rows for the 7x7 init conv: channel (dx*6+c) holds
 * input channel c at horizontal offset dx-3 (zero outside), channels 42..63 zero.  Row b reads
 * x[b % Bx]; cond likewise for rows b < n_cond_rows, zeros for the rest or if cond is NULL
 * (x_self_cond=None, model.py:682). */
int srgd_pack_input(const float* x, const float* cond, int32_t n_cond_rows, int32_t Bx, void* out,
                    int32_t B, int32_t H, int32_t W, srgd_stream_t stream);
/* eps[b][o][y][x] = bias[o] + sum_c w[o][c] * h[b][y][x][c]   (final_conv, 128 -> 3, fp32 NCHW out) */
int srgd_final_conv(const void* h, const float* w, const float* bias, float* eps, int32_t B,
                    int32_t H, int32_t W, int32_t C, int32_t Cout, srgd_stream_t stream);
/* Generic small dense layer on fp32 rows: y[m][n] = bias[n] + sum_k act_in(x[m][k]) * w[n][k],
 * act_in: 0 none, 1 SiLU, 2 GELU(erf).  accumulate!=0 adds into y. */
int srgd_dense_rows(const float* x, const float* w, const float* bias, float* y, int32_t M,
                    int32_t N, int32_t K, int32_t act_in, int32_t accumulate, srgd_stream_t stream);
/* [log_snr, sin(2 pi w log_snr), cos(...)]  (model.py:233-238): out fp32 [B][2*half+1]. */
int srgd_fourier_features(const float* log_snr, const float* weights, float* out, int32_t B,
                          int32_t half_dim, srgd_stream_t stream);
/* SinusoidalPosEmb (model.py:209-221): out fp32 [B][2*half], [sin(t * freq) | cos(t * freq)];  freq[k] =
 * exp(-k * ln(10000) / (half - 1)) is tabulated by the packer with the reference's own torch ops. */
int srgd_sinusoidal_pos_emb(const float* t, const float* freq, float* out, int32_t B, int32_t half_dim,
                            srgd_stream_t stream);
/* t[b][:] += table[labels[b]][:] for labels[b] >= 0   (t = t + class_mlp(label), model.py:692-694) */
int srgd_add_class_rows(float* t, const float* table, const int32_t* labels_dev, int32_t B,
                        int32_t dim, int32_t num_classes, srgd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Whole U-Net (ConditionalSRUnet.forward, model.py:678-725) over a packed parameter list.
 * ------------------------------------------------------------------------------------------ */
typedef struct srgd_unet_config {
  int32_t dim;             /* 128 */
  int32_t n_stages;        /* len(dim_mults), <= 6 */
  int32_t dim_mults[6];
  int32_t full_attn[6];
  int32_t heads, dim_head; /* 4, 32 (dim_head must be 32) */
  int32_t groups;          /* 8 */
  int32_t channels;        /* 3 */
  int32_t sinu_dim;        /* learned_sinusoidal_dim (32) */
  int32_t num_classes;     /* 3, or 0 for none */
  int32_t fixed_sinusoidal;/* 0: RandomOrLearnedSinusoidalPosEmb (sinu_dim + 1 features, model.py:596-598);
                              1: SinusoidalPosEmb(dim) (dim features, model.py:600; the discrete-time family) */
} srgd_unet_config;

typedef struct srgd_unet srgd_unet;

/* Number of device pointers srgd_unet_create expects for this config, and the name of the i-th one
 * (the packing contract; srgd_b200/weights.py produces exactly this list from the checkpoint). */
int srgd_unet_param_count(const srgd_unet_config* cfg);
const char* srgd_unet_param_name(const srgd_unet_config* cfg, int index);
int srgd_unet_create(const srgd_unet_config* cfg, const void* const* params_dev, int n_params,
                     srgd_unet** out);
void srgd_unet_destroy(srgd_unet* u);
size_t srgd_unet_workspace_bytes(const srgd_unet* u, int32_t B, int32_t H, int32_t W);
/* eps[B][3][H][W] = unet(x[b % Bx], log_snr[b], label[b], cond[b % Bx] for b < n_cond_rows).
 * The classifier-free-guidance 2x batch (model.py:3148-3154) is B = 2*Bx rows: rows [0,Bx) are the
 * conditional pass, rows [Bx,2Bx) the null pass (labels -1, or n_cond_rows = Bx).
 * labels_dev: int32 [B] device, entries < 0 select the null label (class_label=None, model.py:692);
 * NULL = all null.  cond_dev may be NULL (all zeros, model.py:682).
 * conv_impl: 0 = tcgen05 implicit GEMM (product path), 1 = CUDA-core direct (debug). */
int srgd_unet_forward(srgd_unet* u, const float* x_dev, const float* cond_dev,
                      const float* log_snr_dev, const int32_t* labels_dev, int32_t n_cond_rows,
                      int32_t Bx, float* eps_dev, int32_t B, int32_t H, int32_t W,
                      void* workspace_dev, size_t workspace_bytes, int32_t conv_impl,
                      srgd_stream_t stream);
/* Debug taps: after a forward, copy the named bf16 NHWC activation (e.g. "downs.0.0") into
 * out_dev; returns its element count through *n, or SRGD_E_ARG if unknown.  Enabled per forward
 * with srgd_unet_set_tap(name) before the call (one tap at a time; NULL disables). */
int srgd_unet_set_tap(srgd_unet* u, const char* name, void* out_dev, size_t out_bytes);
/* Kernels launched by the last srgd_unet_forward on this handle. */
int srgd_unet_last_launch_count(const srgd_unet* u);

/* ------------------------------------------------------------------------------------------
 * Measurement hooks (bench.py): when enabled, every kernel launch made through this library is
 * bracketed by CUDA events on its stream; srgd_profile_end() synchronises and folds them per kind.
 * ------------------------------------------------------------------------------------------ */
enum {
  SRGD_PK_CONV = 0,        /* conv_igemm_kernel (tcgen05)                    */
  SRGD_PK_GN_APPLY = 1,    /* gn_apply_kernel                                */
  SRGD_PK_SAMPLER = 2,     /* sampler_step_kernel                            */
  SRGD_PK_LINEAR_ATTN = 3, /* la_context_partial + merge + apply             */
  SRGD_PK_FULL_ATTN = 4,   /* full_attention_kernel                          */
  SRGD_PK_NORM_MISC = 5,   /* gn_finalize/stats, pixel_inv_norm, rmsnorm     */
  SRGD_PK_OTHER = 6,       /* pack_input, final_conv, embeddings, q_sample…  */
  SRGD_PK_COUNT = 7
};
/* Number of kernels this library has launched from the calling thread since it was loaded (bench.py `gpu_launches`). */
long long srgd_launch_count(void);

/* Batch-invariant reductions (returns the previous setting).  The only reduction of the path whose partition depends
 * on the launch's batch size is the LinearAttention context k.softmax(N) v^T (model.py:317-320): its per-sample partials
 * are split over SMs / B thread blocks.  With `on` != 0 the split of a fixed reference batch is used instead, so a row's
 * result no longer depends on which other rows share its launch: tile-sharded sampling (tiled_sample(shard_tiles=True))
 * then yields a bit-identical image for every world size and call partition.  Process-wide; default off. */
int srgd_set_batch_invariant(int on);

int srgd_profile_begin(void);
int srgd_profile_end(void);
/* Sums since srgd_profile_begin for one kind: device milliseconds, algorithmic FLOPs and bytes the
 * launches were credited with, number of API-level launches. */
int srgd_profile_get(int kind, double* ms, double* flops, double* bytes, int* launches);
/* The individual records behind those sums, in launch order (valid after srgd_profile_end). */
int srgd_profile_record_count(void);
int srgd_profile_record(int index, int* kind, double* ms, double* flops, double* bytes);

#ifdef __cplusplus
}
#endif
#endif /* SRGD_B200_H */
