"""srgd_b200 -- B200-native (sm_100a) implementation of Real-SRGD's sampling hot path.

Python is the host (the reference's own language); all arithmetic runs in libsrgd_b200.so
(include/srgd_b200.h).  Importing the package never touches CUDA; the first compute call loads the
library and fails loudly if it is missing or the device is not sm_100.
"""
from .arch import UnetSpec, unet_keys
from .diffusion import (ConditionalContinuousTimeGaussianDiffusionSR, alpha_cosine_log_snr,
                        beta_linear_log_snr)
from .edm import ConditionalElucidatedDiffusionSR
from .gaussian import ConditionalGaussianDiffusionSR
from .sharding import gather_rows, sample_sharded, shard_counts, shard_range
from .tiling import TilePlan, get_area, get_coord_and_pad, get_coords
from .unet import ConditionalSRUnet

__all__ = ["UnetSpec", "unet_keys", "ConditionalSRUnet", "ConditionalContinuousTimeGaussianDiffusionSR",
           "ConditionalElucidatedDiffusionSR", "ConditionalGaussianDiffusionSR",
           "beta_linear_log_snr", "alpha_cosine_log_snr", "TilePlan", "get_coord_and_pad", "get_coords",
           "get_area", "shard_range", "shard_counts", "gather_rows", "sample_sharded"]
__version__ = "0.1.0"
