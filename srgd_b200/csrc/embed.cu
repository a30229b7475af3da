// U-Net entry/exit kernels and the (tiny) embedding MLPs.
//   pack_input        : cat(x, x_self_cond) (model.py:681-684) -> bf16 row-im2col for the 7x7 init conv
//   final_conv        : final 1x1 conv C -> 3 writing fp32 NCHW eps (model.py:675, 725)
//   fourier_features  : RandomOrLearnedSinusoidalPosEmb (model.py:233-238)
//   sinusoidal_pos_emb: SinusoidalPosEmb (model.py:209-221), the time embedding of the discrete-time family
//   dense_rows        : nn.Linear on a handful of rows with SiLU / GELU on the input
//                       (time_mlp 603-608, class_mlp 612-619, ResnetBlock.mlp 264-267)
//   add_class_rows    : t = t + class_mlp(label) (model.py:692-694)
#include "common.cuh"

namespace srgd {

// out[b][y][x][dx*6 + c] = in6[b % Bx][c][y][x + dx - 3], c < 3 from x, c >= 3 from cond; 42..63 = 0
__global__ void __launch_bounds__(256) pack_input_kernel(const float* __restrict__ x, const float* __restrict__ cond,
                                                         int n_cond_rows, int Bx, bf16* __restrict__ out, int B, int H,
                                                         int W) {
  pdl_wait();
  const int64_t total = (int64_t)B * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int xx = (int)(i % W);
    const int yy = (int)((i / W) % H);
    const int b = (int)(i / ((int64_t)W * H));
    const int bs = b % Bx;
    const bool use_cond = (cond != nullptr) && (b < n_cond_rows);
    float v[64];
#pragma unroll
    for (int j = 42; j < 64; ++j) v[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const bool on = (c < 3) || use_cond;
      const float* plane = (c < 3 || !use_cond) ? x + ((int64_t)bs * 3 + (c % 3)) * H * W + (int64_t)yy * W
                                                : cond + ((int64_t)bs * 3 + (c - 3)) * H * W + (int64_t)yy * W;
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const int xs = xx + dx - 3;
        v[dx * 6 + c] = (on && xs >= 0 && xs < W) ? plane[xs] : 0.f;
      }
    }
    bf16* dst = out + i * 64;
#pragma unroll
    for (int j = 0; j < 64; j += 8) st_stream(dst + j, pack8(v + j));
  }
}

// eps[b][o][pix] = bias[o] + sum_c w[o][c] h[b][pix][c];  a warp owns 32 consecutive pixels,
// half-warps (16 lanes x 8 channels = 128) take one pixel each per iteration.
template <int COUT>
__global__ void __launch_bounds__(256) final_conv_kernel(const bf16* __restrict__ h, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ eps,
                                                         int B, int HW, int C) {
  extern __shared__ float sw[];                          // [COUT][C]
  pdl_wait();
  for (int i = threadIdx.x; i < COUT * C; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, half = lane >> 4, sub = lane & 15;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t n_groups = ((int64_t)B * HW + 31) / 32;   // HW % 32 == 0 is required by the launcher
  for (int64_t grp = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); grp < n_groups;
       grp += (int64_t)gridDim.x * warps_per_block) {
    const int64_t p0 = grp * 32;
    float res[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) res[o] = 0.f;
    for (int it = 0; it < 16; ++it) {
      const int64_t pix = p0 + it * 2 + half;
      float acc[COUT];
#pragma unroll
      for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
      for (int v = sub; v < C / 8; v += 16) {
        float f[8];
        unpack8(ld_stream(h + pix * C + v * 8), f);
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          const float* wr = sw + o * C + v * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[o] = fmaf(f[j], wr[j], acc[o]);
        }
      }
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
#pragma unroll
        for (int s = 8; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
      }
      // lane (it*2 + half) keeps the result for its pixel
      const int owner_lo = it * 2;                       // pixel of half 0
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const float v0 = __shfl_sync(0xffffffffu, acc[o], 0);    // half 0 result
        const float v1 = __shfl_sync(0xffffffffu, acc[o], 16);   // half 1 result
        if (lane == owner_lo) res[o] = v0;
        if (lane == owner_lo + 1) res[o] = v1;
      }
    }
    const int64_t pix = p0 + lane;
    const int b = (int)(pix / HW);
    const int64_t sp = pix % HW;
#pragma unroll
    for (int o = 0; o < COUT; ++o) eps[((int64_t)b * COUT + o) * HW + sp] = res[o] + bias[o];
  }
}

__global__ void fourier_features_kernel(const float* __restrict__ log_snr, const float* __restrict__ wts,
                                        float* __restrict__ out, int B, int half) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int width = 2 * half + 1;
  if (i >= B * width) return;
  const int b = i / width, j = i % width;
  const float t = log_snr[b];
  float v;
  if (j == 0) v = t;
  else {
    const int k = (j - 1) % half;
    // freqs = x * w * 2 * pi, evaluated left to right in fp32 like the reference (model.py:235)
    const float fr = __fmul_rn(__fmul_rn(__fmul_rn(t, wts[k]), 2.0f), 3.14159265358979323846f);
    v = (j - 1) < half ? sinf(fr) : cosf(fr);
  }
  out[i] = v;
}

// emb = t * freq[k]  (x[:, None] * emb[None, :], model.py:219);  out = [sin(emb) | cos(emb)]  (model.py:220)
__global__ void sinusoidal_pos_emb_kernel(const float* __restrict__ t, const float* __restrict__ freq,
                                          float* __restrict__ out, int B, int half) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int width = 2 * half;
  if (i >= B * width) return;
  const int b = i / width, j = i % width;
  const float e = __fmul_rn(t[b], freq[j % half]);
  out[i] = j < half ? sinf(e) : cosf(e);
}

__device__ __forceinline__ float act_in(float v, int act) {
  if (act == 1) return v / (1.0f + expf(-v));                               // SiLU
  if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); // exact GELU (nn.GELU default)
  return v;
}

// block: 8 warps = 8 output features, 8 input rows staged in smem (activation applied on load)
constexpr int kDrRows = 8;
__global__ void __launch_bounds__(256) dense_rows_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int M,
                                                         int N, int K, int act, int accumulate) {
  extern __shared__ float sx[];                          // [kDrRows][K]
  pdl_wait();
  const int m0 = blockIdx.y * kDrRows;
  const int rows = min(kDrRows, M - m0);
  for (int i = threadIdx.x; i < rows * K; i += blockDim.x) sx[i] = act_in(x[(int64_t)m0 * K + i], act);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  float acc[kDrRows];
#pragma unroll
  for (int r = 0; r < kDrRows; ++r) acc[r] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wv = w[(int64_t)n * K + k];
#pragma unroll
    for (int r = 0; r < kDrRows; ++r)
      if (r < rows) acc[r] = fmaf(sx[r * K + k], wv, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kDrRows; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float bv = bias ? bias[n] : 0.f;
    for (int r = 0; r < rows; ++r) {
      float* dst = y + (int64_t)(m0 + r) * N + n;
      *dst = accumulate ? (*dst + acc[r] + bv) : (acc[r] + bv);
    }
  }
}

__global__ void add_class_rows_kernel(float* __restrict__ t, const float* __restrict__ table,
                                      const int32_t* __restrict__ labels, int B, int dim, int num_classes) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * dim) return;
  const int b = i / dim, j = i % dim;
  const int lab = labels[b];
  if (lab >= 0 && lab < num_classes) t[i] += table[lab * dim + j];
}

}  // namespace srgd

using namespace srgd;

extern "C" int srgd_pack_input(const float* x, const float* cond, int32_t n_cond_rows, int32_t Bx, void* out,
                               int32_t B, int32_t H, int32_t W, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x && out && B > 0 && Bx > 0 && Bx <= B && H > 0 && W > 0, "pack_input: bad arguments");
  const int64_t total = (int64_t)B * H * W;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  ProfScope prof(SRGD_PK_OTHER, 0.0, (double)B * H * W * (6 * 4 + 128), as_stream(stream));
  SRGD_CUDA_OK(launch_k(pack_input_kernel, dim3((int)blocks), dim3(256), 0, as_stream(stream), x, cond, n_cond_rows, Bx,
                        reinterpret_cast<bf16*>(out), B, H, W));
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_final_conv(const void* h, const float* w, const float* bias, float* eps, int32_t B, int32_t H,
                               int32_t W, int32_t C, int32_t Cout, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(h && w && bias && eps && B > 0 && H > 0 && W > 0, "final_conv: bad arguments");
  SRGD_REQUIRE(C % 64 == 0 && C <= 1024, "final_conv: C=%d must be a multiple of 64", C);
  SRGD_REQUIRE(Cout == 3, "final_conv: only Cout=3 (channels=3, learned_variance=False) is built, got %d", Cout);
  SRGD_REQUIRE((H * W) % 32 == 0, "final_conv: H*W must be a multiple of 32");
  const int64_t n_groups = (int64_t)B * H * W / 32;
  int64_t blocks = (n_groups + 7) / 8;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  ProfScope prof(SRGD_PK_OTHER, 0.0, (double)B * H * W * (C * 2.0 + 12.0), as_stream(stream));
  SRGD_CUDA_OK(launch_k(final_conv_kernel<3>, dim3((int)blocks), dim3(256), (size_t)3 * C * sizeof(float),
                        as_stream(stream), reinterpret_cast<const bf16*>(h), w, bias, eps, B, H * W, C));
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_fourier_features(const float* log_snr, const float* weights, float* out, int32_t B,
                                     int32_t half_dim, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(log_snr && weights && out && B > 0 && half_dim > 0, "fourier_features: bad arguments");
  const int total = B * (2 * half_dim + 1);
  ProfScope prof(SRGD_PK_OTHER, 0.0, 0.0, as_stream(stream));
  SRGD_CUDA_OK(launch_k(fourier_features_kernel, dim3((total + 127) / 128), dim3(128), 0, as_stream(stream), log_snr,
                        weights, out, B, half_dim));
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_sinusoidal_pos_emb(const float* t, const float* freq, float* out, int32_t B, int32_t half_dim,
                                       srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(t && freq && out && B > 0 && half_dim > 0, "sinusoidal_pos_emb: bad arguments");
  const int total = B * 2 * half_dim;
  ProfScope prof(SRGD_PK_OTHER, 0.0, 0.0, as_stream(stream));
  SRGD_CUDA_OK(launch_k(sinusoidal_pos_emb_kernel, dim3((total + 127) / 128), dim3(128), 0, as_stream(stream), t, freq,
                        out, B, half_dim));
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_dense_rows(const float* x, const float* w, const float* bias, float* y, int32_t M, int32_t N,
                               int32_t K, int32_t act_in, int32_t accumulate, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(x && w && y && M > 0 && N > 0 && K > 0 && K <= 1536, "dense_rows: bad arguments (K <= 1536)");
  SRGD_REQUIRE((M + kDrRows - 1) / kDrRows <= 65535, "dense_rows: too many rows");
  dim3 grid((N + 7) / 8, (M + kDrRows - 1) / kDrRows);
  ProfScope prof(SRGD_PK_OTHER, 0.0, 4.0 * ((double)N * K + (double)M * (N + K)), as_stream(stream));
  SRGD_CUDA_OK(launch_k(dense_rows_kernel, grid, dim3(256), (size_t)kDrRows * K * sizeof(float), as_stream(stream), x, w,
                        bias, y, M, N, K, act_in, accumulate));
  count_launch();
  return SRGD_OK;
}

extern "C" int srgd_add_class_rows(float* t, const float* table, const int32_t* labels_dev, int32_t B, int32_t dim,
                                   int32_t num_classes, srgd_stream_t stream) {
  int rc = check_device();
  if (rc) return rc;
  SRGD_REQUIRE(t && table && labels_dev && B > 0 && dim > 0, "add_class_rows: bad arguments");
  const int total = B * dim;
  ProfScope prof(SRGD_PK_OTHER, 0.0, 0.0, as_stream(stream));
  SRGD_CUDA_OK(launch_k(add_class_rows_kernel, dim3((total + 255) / 256), dim3(256), 0, as_stream(stream), t, table,
                        labels_dev, B, dim, num_classes));
  count_launch();
  return SRGD_OK;
}
