"""Test-only import shim for `denoising-diffusion-pytorch==1.8.15` (requirements.txt:1).

The pip package is absent (no network).  /root/reference/model.py:11-16 imports four names
from it; on the conditional-continuous hot path only two things survive:
  * `Unet.__init__` side effect `self.downsample_factor = 2 ** (len(dim_mults) - 1)`
    (read at model.py:679) -- every parametered attribute is overwritten by the subclass
    (model.py:583-675);
  * `Attend.forward` (see attend.py).
  * for the EDM sampler family (model.py:1731-2600) the `ElucidatedDiffusion` base class
    supplies the preconditioning coefficients and the sigma schedule (restated below, parity unpinned).
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


class GaussianDiffusion(nn.Module):
    pass


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
