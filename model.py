"""Drop-in `model` module: the names `inference.py` and user code import from the reference's
model.py, backed by the srgd_b200 CUDA library.

Only the path the shipped configuration selects is built (conf.model == 'conditional_continuous',
reference model.py:3503-3515 + 3634-3651): `ConditionalSRUnet` and
`ConditionalContinuousTimeGaussianDiffusionSR`.  Other `conf.model` values raise
NotImplementedError (no configuration or weights ship for them).
"""
import copy

import torch
import torch.nn as nn

from srgd_b200 import (ConditionalContinuousTimeGaussianDiffusionSR, ConditionalSRUnet, alpha_cosine_log_snr,
                       beta_linear_log_snr, get_area, get_coord_and_pad, get_coords)

__all__ = ["ConditionalSRUnet", "ConditionalContinuousTimeGaussianDiffusionSR", "get_model", "ModelEma",
           "beta_linear_log_snr", "alpha_cosine_log_snr", "get_coord_and_pad", "get_coords", "get_area",
           "normalize_to_neg_one_to_one", "unnormalize_to_zero_to_one"]


def normalize_to_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_to_zero_to_one(t):
    return (t + 1) * 0.5


class ModelEma(nn.Module):
    """Inference-time stand-in for timm.utils.ModelEmaV2 (reference model.py:3657): holds an eval()
    deep copy as `.module`.  The EMA update itself is training-only and not provided."""

    def __init__(self, model, decay=0.9999, device=None):
        super().__init__()
        self.module = copy.deepcopy(model).eval()
        self.decay = decay
        if device is not None:
            self.module.to(device)


def get_model(conf, logger):
    """reference model.py:3500-3666, conditional_continuous branch."""
    if conf.model != 'conditional_continuous':
        raise NotImplementedError(
            f"conf.model={conf.model!r}: srgd_b200 builds the shipped 'conditional_continuous' sampler only")
    assert conf.learned_sinusoidal_cond
    dim_mults = tuple(int(v) for v in str(conf.ddpm_unet_dim_mults).split(','))
    full_attn = tuple(v.strip() == 'True' for v in str(conf.full_attn).split(','))
    unet = ConditionalSRUnet(dim=conf.unet_dim, dim_mults=dim_mults, full_attn=full_attn,
                             learned_variance=conf.learned_variance,
                             learned_sinusoidal_cond=conf.learned_sinusoidal_cond,
                             learned_sinusoidal_dim=conf.learned_sinusoidal_dim, flash_attn=conf.flash_attn,
                             pixel_shuffle_upsample=conf.pixel_shuffle_upsample, num_classes=conf.num_classes)
    logger.info(f"ConditionalSRUnet: channels=6 dim={conf.unet_dim} dim_mults={conf.ddpm_unet_dim_mults} "
                f"num_classes={conf.num_classes}")
    conf.use_dpmpp_solver = False
    diffusion = ConditionalContinuousTimeGaussianDiffusionSR(
        model=unet, image_size=conf.image_size, noise_schedule=conf.noise_schedule,
        num_sample_steps=conf.num_sample_steps, clip_sample_denoised=conf.clip_sample_denoised,
        learned_schedule_net_hidden_dim=conf.learned_schedule_net_hidden_dim,
        learned_noise_schedule_frac_gradient=conf.learned_noise_schedule_frac_gradient,
        min_snr_loss_weight=conf.min_snr_loss_weight, min_snr_gamma=conf.min_snr_gamma,
        cond_drop_prob=conf.cond_drop_prob, class_cond_drop_prob=conf.class_cond_drop_prob,
        loss_type=conf.loss_type)
    logger.info(f"ConditionalContinuousTimeGaussianDiffusionSR: image_size={conf.image_size} "
                f"num_sample_steps={conf.num_sample_steps}")
    ema_model = ModelEma(diffusion, decay=conf.ema_decay)
    if conf.ckpt_path:
        ckpt = torch.load(conf.ckpt_path, map_location='cpu', weights_only=True)
        check = ema_model.module.load_state_dict(ckpt['ema_model'], strict=conf.load_strict)
        logger.info(f"load ema_model weight from : {conf.ckpt_path}")
        logger.info(f"check: {check}")
    return ema_model
