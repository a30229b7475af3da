"""Drop-in `config` module for the reference CLI surface (reference: config.py:5-194).

`load_config(path) -> Config` accepts the shipped yaml verbatim.  The reference's Config is a flat
dataclass of ~130 mostly training-only fields; only the handful read by `get_model`/`inference.py`
matter to srgd_b200, but every reference field name is accepted (with the reference default) and an
unknown key raises TypeError exactly like `Config(**opts)` does there.
"""
import yaml

# field -> default, grouped by who reads them
_MODEL_FIELDS = dict(
    model='continuous', noise_schedule='linear', num_sample_steps=32, clip_sample_denoised=True,
    image_size=128, unet_dim=64, ddpm_unet_dim_mults='1,2,4,8', full_attn='False,False,False,True',
    learned_variance=False, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, flash_attn=False,
    pixel_shuffle_upsample=True, num_classes=3, ema_decay=0.995, ema_device='cuda', ckpt_path='',
    load_strict=True, learned_schedule_net_hidden_dim=1024, learned_noise_schedule_frac_gradient=1.,
    min_snr_loss_weight=False, min_snr_gamma=5, cond_drop_prob=0.1, class_cond_drop_prob=0.1, loss_type='l2',
    use_dpmpp_solver=True,
)
_SAMPLING_FIELDS = dict(
    cond_scale=1., class_cond_scale=1., test_label=0, guidance_start_steps=0, class_guidance_start_steps=0,
    generation_start_steps=0, seed=71, amp=False, amp_dtype='float16', device='cuda',
)
_OTHER_SAMPLER_FIELDS = dict(          # discrete-time / EDM families (not built by srgd_b200)
    objective='pred_noise', beta_schedule='linear', timesteps=1000, sampling_timesteps=250,
    offset_noise_strength=0., sigma_min=0.002, sigma_max=80, sigma_data=0.5, rho=7, P_mean=-1.2, P_std=1.2,
    S_churn=80, S_tmin=0.05, S_tmax=50, S_noise=1.003,
)
_TRAINING_FIELDS = dict(               # accepted and ignored: the reference ships no trainer
    save_dir='srgd', prefix='conditional_continuous_linear', base_dir='./input/',
    dataset_name='cropped_df2kost_400x400_overlap200', conditional_task_type='realsr_denoise_sr',
    val_num_sample_steps=32, n_fold=10, train_fold='0', skip_sample=False, skip_val=False, validation_ratio=0.5,
    val_realsrv3=False, val_drealsr=False, val_realsrv3_scale=4, val_drealsr_scale=4, crop_size=256,
    hr_image_size=256, lr_image_size=128, crop_rate=2, scale_size=256, crop_size_limit=False, batch_size=32,
    sample_size=16, hflip=False, rotate=False, interpolation='BICUBIC', shuffle=True, torch_compile=False,
    optimizer='adamw', lr=1e-4, min_lr=1e-4, weight_decay=0., momentum=0.9, nesterov=False, amsgrad=False,
    madgrad_decoupled_decay=True, epochs=300, warmup_epochs=0, warmup_lr_init=1e-6, plateau_mode='min',
    factor=0.1, patience=4, plateau_eps=1e-8, scheduler='cosine', cosine_interval_type='step',
    train_preprocess='randomcrop', valid_preprocess='centercrop', train_trans_mode='realesrgan',
    valid_trans_mode='simple', usm_sharpener=False, blur_prob=0.5, advance_blur_prob=0.5, gaussian_blur_prob=0.5,
    sinc_blur_prob=0.5, sinc_blur_factor_min=0.9, sinc_blur_factor_max=1.1, image_compression_prob=0.5,
    quality_lower=50, quality_upper=100, noise_prob=0.5, gauss_noise_prob=0.5, iso_noise_prob=0.5,
    multiplicative_noise_prob=0.5, train=True, test=False, debug=False, save_validation_sample=False,
    save_validation_hr_sample=False, save_every_epoch=False, test_target='best_loss', num_workers=4,
    pin_memory=True, model_dir='models', log_dir='logs', print_freq=0,
)
_DEFAULTS = {**_MODEL_FIELDS, **_SAMPLING_FIELDS, **_OTHER_SAMPLER_FIELDS, **_TRAINING_FIELDS}


class Config:
    """Flat attribute bag with the reference's field names and defaults."""

    def __init__(self, **overrides):
        unknown = [k for k in overrides if k not in _DEFAULTS]
        if unknown:
            raise TypeError(f"Config.__init__() got an unexpected keyword argument '{unknown[0]}'")
        self.__dict__.update(_DEFAULTS)
        self.__dict__.update(overrides)

    def __repr__(self):
        changed = {k: v for k, v in self.__dict__.items() if _DEFAULTS.get(k, object()) != v}
        return f"Config({', '.join(f'{k}={v!r}' for k, v in changed.items())})"

    def __eq__(self, other):
        return isinstance(other, Config) and self.__dict__ == other.__dict__


def load_config(config_file):
    with open(config_file, 'r') as fp:
        return Config(**(yaml.safe_load(fp) or {}))
