"""Test-only import shim for `denoising-diffusion-pytorch==1.8.15` (requirements.txt:1).

The pip package is absent (no network).  /root/reference/model.py:11-16 imports four names
from it; on the conditional-continuous hot path only two things survive:
  * `Unet.__init__` side effect `self.downsample_factor = 2 ** (len(dim_mults) - 1)`
    (read at model.py:679) -- every parametered attribute is overwritten by the subclass
    (model.py:583-675);
  * `Attend.forward` (see attend.py).
Nothing here is product code; it exists so tests/golden/make_golden.py can import the
UNMODIFIED reference.
"""
import torch.nn as nn


class Unet(nn.Module):
    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), *args, **kwargs):
        super().__init__()
        self.downsample_factor = 2 ** (len(dim_mults) - 1)


class GaussianDiffusion(nn.Module):
    pass


class ElucidatedDiffusion(nn.Module):
    pass
