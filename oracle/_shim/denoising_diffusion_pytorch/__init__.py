"""Test-only import shim for `denoising-diffusion-pytorch==1.8.15` (requirements.txt:1).

The pip package is absent (no network).  /root/reference/model.py:11-16 imports four names
from it; on the conditional-continuous hot path only two things survive:
  * `Unet.__init__` side effect `self.downsample_factor = 2 ** (len(dim_mults) - 1)`
    (read at model.py:679) -- every parametered attribute is overwritten by the subclass
    (model.py:583-675);
  * `Attend.forward` (see attend.py).
  * for the EDM sampler family (model.py:1731-2600) the `ElucidatedDiffusion` base class
    supplies the preconditioning coefficients and the sigma schedule (restated below, parity unpinned).
  * for the discrete-time family (model.py:781-1727) the `GaussianDiffusion` base class
    supplies `device`, `predict_start_from_noise`, `predict_noise_from_start`, `predict_start_from_v`, `q_posterior`
    and `q_sample` (restated below, parity unpinned); everything else (schedules, buffers) the reference subclass
    builds itself (model.py:1362-1424).
Nothing here is product code; it exists so tests/golden/make_golden.py can import the
UNMODIFIED reference.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Unet(nn.Module):
    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), *args, **kwargs):
        super().__init__()
        self.downsample_factor = 2 ** (len(dim_mults) - 1)


def _extract(a, t, x_shape):
    b, *_ = t.shape
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


class GaussianDiffusion(nn.Module):
    """PARITY UNPINNED restatement of what /root/reference/model.py's discrete-time classes inherit from the pip
    package's `GaussianDiffusion` (denoising_diffusion_pytorch.py of 1.8.15; Ho et al. 2020 eq. 4, 7, 15 and the
    v-parameterisation of Salimans & Ho 2022): the helpers below read the buffers the reference subclass registers
    itself (model.py:1395-1424).  Call sites: model.py:1474, 1478, 1483, 1487, 1489 (predict_*), 1499 (q_posterior),
    1526 / 1581 (q_sample), 1506 / 1521 (device).  The constructor of the real class builds the same buffers again;
    the subclass overwrites every one of them, so it is reduced to nn.Module bookkeeping here."""

    def __init__(self, model, *, image_size, **kwargs):
        super().__init__()
        self.model = model

    @property
    def device(self):
        return self.betas.device

    def predict_start_from_noise(self, x_t, t, noise):
        return (_extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t -
                _extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    def predict_noise_from_start(self, x_t, t, x0):
        return ((_extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0) /
                _extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape))

    def predict_v(self, x_start, t, noise):
        return (_extract(self.sqrt_alphas_cumprod, t, x_start.shape) * noise -
                _extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * x_start)

    def predict_start_from_v(self, x_t, t, v):
        return (_extract(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t -
                _extract(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v)

    def q_posterior(self, x_start, x_t, t):
        posterior_mean = (_extract(self.posterior_mean_coef1, t, x_t.shape) * x_start +
                          _extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        posterior_variance = _extract(self.posterior_variance, t, x_t.shape)
        posterior_log_variance_clipped = _extract(self.posterior_log_variance_clipped, t, x_t.shape)
        return posterior_mean, posterior_variance, posterior_log_variance_clipped

    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_start)
        return (_extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
                _extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)


class ElucidatedDiffusion(nn.Module):
    """PARITY UNPINNED restatement of what /root/reference/model.py's EDM classes inherit from the pip package's
    `ElucidatedDiffusion` (elucidated_diffusion.py of denoising-diffusion-pytorch 1.8.15; Karras et al. 2022, Table 1):
    the constructor bookkeeping, `device`, the preconditioning coefficients c_in / c_out / c_skip / c_noise and the
    rho-schedule of sigmas with a trailing zero.  Call sites: model.py:2086-2096 (super().__init__), 2141-2147
    (c_in, c_noise, c_skip, c_out), 2190 and 2232 (sample_schedule), 2247 (device)."""

    def __init__(self, net, *, image_size, channels=3, num_sample_steps=32, sigma_min=0.002, sigma_max=80,
                 sigma_data=0.5, rho=7, P_mean=-1.2, P_std=1.2, S_churn=80, S_tmin=0.05, S_tmax=50, S_noise=1.003):
        super().__init__()
        self.net = net
        self.sigma_min, self.sigma_max, self.sigma_data, self.rho = sigma_min, sigma_max, sigma_data, rho
        self.num_sample_steps = num_sample_steps

    @property
    def device(self):
        return next(self.net.parameters()).device

    def c_skip(self, sigma):
        return (self.sigma_data ** 2) / (sigma ** 2 + self.sigma_data ** 2)

    def c_out(self, sigma):
        return sigma * self.sigma_data * (self.sigma_data ** 2 + sigma ** 2) ** -0.5

    def c_in(self, sigma):
        return 1 * (sigma ** 2 + self.sigma_data ** 2) ** -0.5

    def c_noise(self, sigma):
        return torch.log(sigma.clamp(min=1e-20)) * 0.25

    def sample_schedule(self, num_sample_steps=None):
        N = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        inv_rho = 1 / self.rho
        steps = torch.arange(N, device=self.device, dtype=torch.float32)
        sigmas = (self.sigma_max ** inv_rho + steps / (N - 1) * (self.sigma_min ** inv_rho - self.sigma_max ** inv_rho)) ** self.rho
        return F.pad(sigmas, (0, 1), value=0.)          # last step is sigma value of 0.
