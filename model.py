"""Drop-in `model` module: the names `inference.py` and user code import from the reference's
model.py, backed by the srgd_b200 CUDA library.

Built: the path the shipped configuration selects (conf.model == 'conditional_continuous', reference
model.py:3503-3515 + 3634-3651: `ConditionalSRUnet` + `ConditionalContinuousTimeGaussianDiffusionSR`) and the two
other class-conditional sampler families on the same U-Net: EDM (conf.model == 'conditional_elucidated',
model.py:3593-3614: `ConditionalElucidatedDiffusionSR`) and discrete-time DDPM / DDIM (conf.model ==
'conditional_gaussian', model.py:3552-3570: `ConditionalGaussianDiffusionSR`).  Other `conf.model` values raise
NotImplementedError (the non-conditional twins; no configuration or weights ship for them).
"""
import copy
import os

import torch
import torch.nn as nn

from srgd_b200 import (ConditionalContinuousTimeGaussianDiffusionSR, ConditionalElucidatedDiffusionSR,
                       ConditionalGaussianDiffusionSR, ConditionalSRUnet, alpha_cosine_log_snr, beta_linear_log_snr,
                       get_area, get_coord_and_pad, get_coords)

__all__ = ["ConditionalSRUnet", "ConditionalContinuousTimeGaussianDiffusionSR", "ConditionalElucidatedDiffusionSR",
           "ConditionalGaussianDiffusionSR", "get_model", "ModelEma",
           "beta_linear_log_snr", "alpha_cosine_log_snr", "get_coord_and_pad", "get_coords", "get_area",
           "normalize_to_neg_one_to_one", "unnormalize_to_zero_to_one"]


def normalize_to_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_to_zero_to_one(t):
    return (t + 1) * 0.5


class ModelEma(nn.Module):
    """Inference-time stand-in for timm.utils.ModelEmaV2 (reference model.py:3657): holds an eval()
    deep copy as `.module`.  The EMA update itself is training-only and not provided."""

    def __init__(self, model, decay=0.9999, device=None, copy_model=True):
        super().__init__()
        # copy_model=False: take ownership instead of deep-copying 550 MB (get_model keeps no other reference)
        self.module = (copy.deepcopy(model) if copy_model else model).eval()
        self.decay = decay
        if device is not None:
            self.module.to(device)


def get_model(conf, logger):
    """reference model.py:3500-3666: the shipped conditional_continuous branch (3503-3515, 3634-3651), the
    conditional_elucidated branch (EDM sampler family on the same U-Net, 3593-3614) and the conditional_gaussian
    branch (discrete-time DDPM / DDIM, 3552-3570)."""
    if conf.model not in ('conditional_continuous', 'conditional_elucidated', 'conditional_gaussian'):
        raise NotImplementedError(
            f"conf.model={conf.model!r}: srgd_b200 builds the class-conditional samplers 'conditional_continuous' "
            "(shipped), 'conditional_elucidated' and 'conditional_gaussian' only")
    if conf.model == 'conditional_gaussian':
        assert not conf.learned_sinusoidal_cond                                # model.py:3553
    else:
        assert conf.learned_sinusoidal_cond
    dim_mults = tuple(int(v) for v in str(conf.ddpm_unet_dim_mults).split(','))
    full_attn = tuple(v.strip() == 'True' for v in str(conf.full_attn).split(','))
    from srgd_b200 import weights as _weights
    from srgd_b200.arch import UnetSpec
    ckpt_path = conf.ckpt_path or None
    # ingest cache: the packed device weights of an earlier start of this checkpoint
    spec = UnetSpec(dim=conf.unet_dim, dim_mults=dim_mults, full_attn=full_attn,
                    learned_sinusoidal_dim=conf.learned_sinusoidal_dim, num_classes=conf.num_classes,
                    learned_sinusoidal_cond=bool(conf.learned_sinusoidal_cond))
    use_cache = bool(ckpt_path) and os.environ.get("SRGD_B200_PACK_CACHE", "1") != "0"
    cached = _weights.load_pack_cache(ckpt_path, spec) if use_cache else None
    unet = ConditionalSRUnet(dim=conf.unet_dim, dim_mults=dim_mults, full_attn=full_attn,
                             learned_variance=conf.learned_variance,
                             learned_sinusoidal_cond=conf.learned_sinusoidal_cond,
                             learned_sinusoidal_dim=conf.learned_sinusoidal_dim, flash_attn=conf.flash_attn,
                             pixel_shuffle_upsample=conf.pixel_shuffle_upsample, num_classes=conf.num_classes,
                             _init_weights=not ckpt_path)
    assert unet.spec == spec
    logger.info(f"ConditionalSRUnet: channels=6 dim={conf.unet_dim} dim_mults={conf.ddpm_unet_dim_mults} "
                f"num_classes={conf.num_classes}")
    if conf.model == 'conditional_elucidated':
        diffusion = ConditionalElucidatedDiffusionSR(
            unet, image_size=conf.image_size, num_sample_steps=conf.num_sample_steps, sigma_min=conf.sigma_min,
            sigma_max=conf.sigma_max, sigma_data=conf.sigma_data, rho=conf.rho, P_mean=conf.P_mean, P_std=conf.P_std,
            S_churn=conf.S_churn, S_tmin=conf.S_tmin, S_tmax=conf.S_tmax, S_noise=conf.S_noise,
            cond_drop_prob=conf.cond_drop_prob, class_cond_drop_prob=conf.class_cond_drop_prob,
            use_dpmpp_solver=conf.use_dpmpp_solver, loss_type=conf.loss_type)
        logger.info(f"ConditionalElucidatedDiffusionSR: image_size={conf.image_size} "
                    f"num_sample_steps={conf.num_sample_steps}")
        return _finish(conf, logger, diffusion, unet, spec, ckpt_path, use_cache, cached, _weights)
    conf.use_dpmpp_solver = False
    if conf.model == 'conditional_gaussian':
        diffusion = ConditionalGaussianDiffusionSR(
            model=unet, image_size=conf.image_size, timesteps=conf.timesteps,
            sampling_timesteps=conf.sampling_timesteps, objective=conf.objective, beta_schedule=conf.beta_schedule,
            offset_noise_strength=conf.offset_noise_strength, min_snr_loss_weight=conf.min_snr_loss_weight,
            min_snr_gamma=conf.min_snr_gamma, cond_drop_prob=conf.cond_drop_prob,
            class_cond_drop_prob=conf.class_cond_drop_prob, loss_type=conf.loss_type)
        logger.info(f"ConditionalGaussianDiffusionSR: image_size={conf.image_size} timesteps={conf.timesteps} "
                    f"sampling_timesteps={conf.sampling_timesteps}")
        return _finish(conf, logger, diffusion, unet, spec, ckpt_path, use_cache, cached, _weights)
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
    return _finish(conf, logger, diffusion, unet, spec, ckpt_path, use_cache, cached, _weights)


def _finish(conf, logger, diffusion, unet, spec, ckpt_path, use_cache, cached, _weights):
    """EMA holder + checkpoint (model.py:3657-3664), through the ingest cache when there is one."""
    ema_model = ModelEma(diffusion, decay=conf.ema_decay, copy_model=False)
    if ckpt_path and cached is not None:
        unet.attach_pack_cache(cached, ckpt_path,
                               prefix="net." if isinstance(diffusion, ConditionalElucidatedDiffusionSR) else "model.")
        logger.info(f"load ema_model weight from : {ckpt_path} (ingest cache {_weights.pack_cache_path(ckpt_path)}; "
                    "fp32 parameters are read from the checkpoint on first state_dict())")
    elif ckpt_path:
        ckpt = torch.load(ckpt_path, map_location='cpu', weights_only=True)
        check = ema_model.module.load_state_dict(ckpt['ema_model'], strict=conf.load_strict)
        logger.info(f"load ema_model weight from : {ckpt_path}")
        logger.info(f"check: {check}")
        if use_cache and not check.missing_keys and not check.unexpected_keys:
            unet.save_pack_cache_after_first_pack(ckpt_path)
    return ema_model
