"""Host-side mirror of the reference's `ConditionalGaussianDiffusionSR` (model.py:1311-1660, sampling half): the
discrete-time sampler family -- DDPM ancestral sampling over all `timesteps` and DDIM over `sampling_timesteps` --
on the same class- and LR-conditioned U-Net, here with the fixed `SinusoidalPosEmb` time embedding
(model.py:209-221, 600) -- SURVEY.md section 8 f-4.

Kept from the reference: constructor keywords, the registered buffers (names, fp32 values computed in float64 with the
reference's own schedule functions, model.py:744-777 / 1362-1424, so a reference checkpoint loads strictly),
`model_predictions`, `p_mean_variance`, `p_sample`, `p_sample_loop`, `ddim_sample`, `sample`, the RNG call order on
torch's generator and the `NotImplementedError` for two guidance scales.  Five formulas live in the pip package's
`GaussianDiffusion` base class in the reference (`predict_start_from_noise`, `predict_noise_from_start`,
`predict_start_from_v`, `q_posterior`, `q_sample`); they are restated here from the published DDPM algebra
(parity unpinned at that boundary, like `Attend`).

Underneath: the U-Net forward is `srgd_unet_forward` with the guidance pair as ONE 2x batch (the reference runs two
forwards, model.py:1460-1469); guidance combine + x_start + clamp + noise re-derivation + posterior / DDIM update are
one fused kernel (`srgd_gauss_update`, csrc/gaussian.cu); the per-step coefficients are host look-ups in fp32 copies
of the buffers.  No shipped configuration or weights select this family (conf.model == 'conditional_gaussian'), and
the reference CLI cannot drive it either (inference.py:84 calls `tiled_sample`, which this class does not have).
"""
from __future__ import annotations

import ctypes as C
import math
from collections import namedtuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

try:
    from tqdm import tqdm as _tqdm
except Exception:                      # pragma: no cover
    _tqdm = None

ModelPrediction = namedtuple('ModelPrediction', ['pred_noise', 'pred_x_start'])           # model.py:36


def linear_beta_schedule(timesteps):
    """model.py:744-751"""
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    """model.py:753-763"""
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    ac = torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def sigmoid_beta_schedule(timesteps, start=-3, end=3, tau=1, clamp_min=1e-5):
    """model.py:765-777"""
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def extract(a, t, x_shape):
    """model.py:730-733"""
    b, *_ = t.shape
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


_OBJECTIVES = {'pred_noise': _lib.OBJ_PRED_NOISE, 'pred_x0': _lib.OBJ_PRED_X0, 'pred_v': _lib.OBJ_PRED_V}
_BUFFERS = ('betas', 'alphas_cumprod', 'alphas_cumprod_prev', 'sqrt_alphas_cumprod', 'sqrt_one_minus_alphas_cumprod',
            'log_one_minus_alphas_cumprod', 'sqrt_recip_alphas_cumprod', 'sqrt_recipm1_alphas_cumprod',
            'posterior_variance', 'posterior_log_variance_clipped', 'posterior_mean_coef1', 'posterior_mean_coef2',
            'loss_weight')


class ConditionalGaussianDiffusionSR(nn.Module):
    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None, objective='pred_v',
                 beta_schedule='sigmoid', schedule_fn_kwargs=dict(), ddim_sampling_eta=0., auto_normalize=True,
                 offset_noise_strength=0., min_snr_loss_weight=False, min_snr_gamma=5, cond_drop_prob=0.,
                 class_cond_drop_prob=0., loss_type='l2'):
        super().__init__()
        assert not model.random_or_learned_sinusoidal_cond                                   # model.py:1349
        self.model = model
        self.channels = self.model.channels
        self.self_condition = self.model.self_condition
        self.image_size = image_size
        self.objective = objective
        assert objective in _OBJECTIVES, \
            'objective must be either pred_noise (predict noise) or pred_x0 (predict image start) or pred_v (predict v)'
        if beta_schedule == 'linear':
            beta_schedule_fn = linear_beta_schedule
        elif beta_schedule == 'cosine':
            beta_schedule_fn = cosine_beta_schedule
        elif beta_schedule == 'sigmoid':
            beta_schedule_fn = sigmoid_beta_schedule
        else:
            raise ValueError(f'unknown beta schedule {beta_schedule}')
        betas = beta_schedule_fn(timesteps, **schedule_fn_kwargs)
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, dim=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.)
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)
        self.sampling_timesteps = timesteps if sampling_timesteps is None else sampling_timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = ddim_sampling_eta

        reg = lambda name, val: self.register_buffer(name, val.to(torch.float32))           # model.py:1393
        reg('betas', betas)
        reg('alphas_cumprod', alphas_cumprod)
        reg('alphas_cumprod_prev', alphas_cumprod_prev)
        reg('sqrt_alphas_cumprod', torch.sqrt(alphas_cumprod))
        reg('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - alphas_cumprod))
        reg('log_one_minus_alphas_cumprod', torch.log(1. - alphas_cumprod))
        reg('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alphas_cumprod))
        reg('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        reg('posterior_variance', posterior_variance)
        reg('posterior_log_variance_clipped', torch.log(posterior_variance.clamp(min=1e-20)))
        reg('posterior_mean_coef1', betas * torch.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        reg('posterior_mean_coef2', (1. - alphas_cumprod_prev) * torch.sqrt(alphas) / (1. - alphas_cumprod))
        self.offset_noise_strength = offset_noise_strength
        snr = alphas_cumprod / (1 - alphas_cumprod)
        maybe_clipped_snr = snr.clone()
        if min_snr_loss_weight:
            maybe_clipped_snr.clamp_(max=min_snr_gamma)
        if objective == 'pred_noise':
            reg('loss_weight', maybe_clipped_snr / snr)
        elif objective == 'pred_x0':
            reg('loss_weight', maybe_clipped_snr)
        else:
            reg('loss_weight', maybe_clipped_snr / (snr + 1))
        self.auto_normalize = auto_normalize
        self.cond_drop_prob, self.class_cond_drop_prob = cond_drop_prob, class_cond_drop_prob
        self.loss_type = loss_type
        self.progress = True
        self.rng_device = None          # "cpu": draw from torch's global CPU generator (parity runs vs a CPU reference)
        self.last_step_launches = 0
        self.__dict__["_host"] = None   # fp32 CPU copies of the buffers for the per-step look-ups
        self.register_load_state_dict_post_hook(ConditionalGaussianDiffusionSR._after_load)

    # -- plumbing ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _after_load(module, incompatible_keys) -> None:
        module.__dict__["_host"] = None

    def _apply(self, fn, *a, **kw):
        r = super()._apply(fn, *a, **kw)
        self.__dict__["_host"] = None
        return r

    def _tables(self):
        host = self.__dict__.get("_host")
        if host is None:
            host = self.__dict__["_host"] = {k: getattr(self, k).detach().to("cpu", torch.float32) for k in _BUFFERS}
        return host

    def normalize(self, img):
        return img * 2 - 1 if self.auto_normalize else img

    def unnormalize(self, t):
        return (t + 1) * 0.5 if self.auto_normalize else t

    @property
    def device(self):
        return self.betas.device                                 # pip base class

    def set_seed(self, seed):
        torch.cuda.manual_seed(seed)

    def _randn(self, shape, device):
        if self.rng_device is not None and torch.device(self.rng_device).type == "cpu":
            return torch.randn(tuple(shape)).to(device)
        return torch.randn(tuple(shape), device=device)

    def _iter(self, seq, total):
        if self.progress and _tqdm is not None:
            return _tqdm(seq, desc='sampling loop time step', total=total)
        return seq

    @staticmethod
    def _uniform_time(t) -> int:
        """The timestep of a batch that shares one (every sampling call); per-sample timesteps are a training-time use."""
        if torch.is_tensor(t):
            tt = t.reshape(-1)
            if tt.numel() > 1 and not bool((tt == tt[0]).all()):
                raise NotImplementedError("per-sample timesteps are only used by the training loss (not shipped)")
            return int(tt[0])
        return int(t)

    # -- pip base class (restated; parity unpinned) ------------------------------------------------------------------
    def predict_start_from_noise(self, x_t, t, noise):
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t -
                extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    def predict_noise_from_start(self, x_t, t, x0):
        return ((extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0) /
                extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape))

    def predict_v(self, x_start, t, noise):
        return (extract(self.sqrt_alphas_cumprod, t, x_start.shape) * noise -
                extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * x_start)

    def predict_start_from_v(self, x_t, t, v):
        return (extract(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t -
                extract(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v)

    def q_posterior(self, x_start, x_t, t):
        posterior_mean = (extract(self.posterior_mean_coef1, t, x_t.shape) * x_start +
                          extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        return (posterior_mean, extract(self.posterior_variance, t, x_t.shape),
                extract(self.posterior_log_variance_clipped, t, x_t.shape))

    def q_sample(self, x_start, t, noise=None):
        """sqrt_alphas_cumprod[t] * x_start + sqrt_one_minus_alphas_cumprod[t] * noise."""
        if noise is None:
            noise = self._randn(x_start.shape, x_start.device)                              # randn_like(x_start)
        else:
            _lib.require_like(noise, x_start, "q_sample noise")
        tt = t.reshape(-1) if torch.is_tensor(t) else torch.tensor([int(t)])
        if x_start.is_cuda and (tt.numel() == 1 or bool((tt == tt[0]).all())):
            tab, k = self._tables(), int(tt[0])
            out = torch.empty_like(x_start, dtype=torch.float32)
            with torch.cuda.device(x_start.device):
                rc = _lib.load().srgd_q_sample(_lib.ptr(x_start.contiguous().float()),
                                               _lib.ptr(noise.contiguous().float()), _lib.ptr(out), out.numel(),
                                               float(tab['sqrt_alphas_cumprod'][k]),
                                               float(tab['sqrt_one_minus_alphas_cumprod'][k]), _lib.current_stream())
            _lib.check(rc, "srgd_q_sample")
            return out
        tt = tt.to(x_start.device)
        return (extract(self.sqrt_alphas_cumprod, tt, x_start.shape) * x_start +
                extract(self.sqrt_one_minus_alphas_cumprod, tt, x_start.shape) * noise)

    # -- denoiser call with classifier-free guidance (model.py:1453-1469) ---------------------------------------------
    def _predict(self, x, t: int, condition_x, class_label, cond_scale, class_cond_scale):
        """Returns (out_cond, out_null or None, guidance scale)."""
        if (cond_scale != 1.0) and (class_cond_scale != 1.0):
            raise NotImplementedError(
                "Currently, you cannot specify both cond_scale and class_cond_scale at the same time.")
        unet, B, dev = self.model, x.shape[0], x.device
        if cond_scale != 1.0:                    # LR-condition guidance: null rows drop condition_x
            rows, n_cond, scale = 2 * B, B, cond_scale
            labels = unet.labels_for(class_label, B, dev)
            labels = None if labels is None else torch.cat((labels, labels))
        elif class_cond_scale != 1.0:            # class guidance: null rows drop the label
            rows, n_cond, scale = 2 * B, 2 * B, class_cond_scale
            labels = unet.labels_for(class_label, B, dev)
            labels = None if labels is None else torch.cat((labels, torch.full_like(labels, -1)))
        else:
            rows, n_cond, scale = B, B, 1.0
            labels = unet.labels_for(class_label, B, dev)
        if condition_x is None:
            n_cond = 0
        times = torch.full((rows,), float(t), device=dev, dtype=torch.float32)     # long -> float, model.py:219
        out = unet.run(x, times, labels, condition_x, rows, n_cond)
        self.last_step_launches = unet.last_launches
        return (out, None, 1.0) if rows == B else (out[:B], out[B:], scale)

    def _scalars(self, t: int, mode: int, clip: bool, rederive: bool, scale: float, time_next: int = -1):
        tab = self._tables()
        s = _lib.GaussScalars()
        s.objective, s.mode, s.clip, s.rederive = _OBJECTIVES[self.objective], mode, int(bool(clip)), int(bool(rederive))
        s.guidance_scale = float(scale)
        s.sqrt_recip_ac = float(tab['sqrt_recip_alphas_cumprod'][t])
        s.sqrt_recipm1_ac = float(tab['sqrt_recipm1_alphas_cumprod'][t])
        s.sqrt_ac = float(tab['sqrt_alphas_cumprod'][t])
        s.sqrt_1m_ac = float(tab['sqrt_one_minus_alphas_cumprod'][t])
        s.coef1, s.coef2 = float(tab['posterior_mean_coef1'][t]), float(tab['posterior_mean_coef2'][t])
        if mode == _lib.GAUSS_DDPM:
            s.noise_scale = float((0.5 * tab['posterior_log_variance_clipped'][t]).exp())   # model.py:1513
        elif mode == _lib.GAUSS_DDIM:
            alpha, alpha_next = tab['alphas_cumprod'][t], tab['alphas_cumprod'][time_next]  # model.py:1608-1612
            sigma = self.ddim_sampling_eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            s.noise_scale = float(sigma)
            s.c = float((1 - alpha_next - sigma ** 2).sqrt())
            s.sqrt_ac_next = float(alpha_next.sqrt())
        return s

    def _update(self, x, out_c, out_n, noise, s: _lib.GaussScalars, want_img=True, want_x0=True, want_noise=False):
        mk = lambda want: torch.empty_like(x) if want else None
        img, x0, pn = mk(want_img), mk(want_x0), mk(want_noise)
        with torch.cuda.device(x.device):
            rc = _lib.load().srgd_gauss_update(_lib.ptr(x), _lib.ptr(out_c), _lib.ptr(out_n), _lib.ptr(noise),
                                               _lib.ptr(img), _lib.ptr(x0), _lib.ptr(pn), x.numel(), C.byref(s),
                                               _lib.current_stream())
        _lib.check(rc, "srgd_gauss_update")
        self.last_step_launches += 1
        return img, x0, pn

    # -- reference surface -------------------------------------------------------------------------------------------
    def model_predictions(self, x, t, condition_x=None, class_label=None, cond_scale=1.0, class_cond_scale=1.0,
                          clip_x_start=False, rederive_pred_noise=False):
        """model.py:1449-1489"""
        _lib.require_cuda(x, "model_predictions")
        x, k = x.contiguous().float(), self._uniform_time(t)
        out_c, out_n, scale = self._predict(x, k, condition_x, class_label, cond_scale, class_cond_scale)
        s = self._scalars(k, _lib.GAUSS_DDIM_LAST, clip_x_start, rederive_pred_noise, scale)
        _, x0, pn = self._update(x, out_c, out_n, None, s, want_img=False, want_noise=True)
        return ModelPrediction(pn, x0)

    def p_mean_variance(self, x, t, condition_x=None, class_label=None, cond_scale=1.0, class_cond_scale=1.0,
                        clip_denoised=True):
        """model.py:1491-1500"""
        _lib.require_cuda(x, "p_mean_variance")
        x, k = x.contiguous().float(), self._uniform_time(t)
        out_c, out_n, scale = self._predict(x, k, condition_x, class_label, cond_scale, class_cond_scale)
        s = self._scalars(k, _lib.GAUSS_DDPM, clip_denoised, False, scale)
        mean, x0, _ = self._update(x, out_c, out_n, None, s)
        tab, shape = self._tables(), (x.shape[0],) + (1,) * (x.ndim - 1)
        var = tab['posterior_variance'][k].to(x.device).expand(shape)
        logvar = tab['posterior_log_variance_clipped'][k].to(x.device).expand(shape)
        return mean, var, logvar, x0

    @torch.inference_mode()
    def p_sample(self, x, t: int, condition_x=None, class_label=None, cond_scale=1.0, class_cond_scale=1.0, noise=None):
        """model.py:1503-1514"""
        _lib.require_cuda(x, "p_sample")
        x, k = x.contiguous().float(), int(t)
        out_c, out_n, scale = self._predict(x, k, condition_x, class_label, cond_scale, class_cond_scale)
        s = self._scalars(k, _lib.GAUSS_DDPM, True, False, scale)
        if k > 0 and noise is None:
            noise = self._randn(x.shape, x.device)                                          # model.py:1512
        elif noise is not None:
            _lib.require_like(noise, x, "p_sample noise")
        img, x0, _ = self._update(x, out_c, out_n, noise.contiguous().float() if k > 0 else None, s)
        return img, x0

    def _start_image(self, shape, condition_x, target_time, generation_start_steps):
        if generation_start_steps > 0:
            return self.q_sample(x_start=condition_x, t=int(target_time))                   # model.py:1526, 1581
        return self._randn(shape, self.device)

    @torch.inference_mode()
    def p_sample_loop(self, shape, condition_x, class_label, cond_scale, guidance_start_steps, class_cond_scale,
                      class_guidance_start_steps, generation_start_steps, sampling_timesteps, with_images,
                      with_x0_images):
        """model.py:1517-1563: ancestral sampling over ALL num_timesteps (sampling_timesteps is not read here)."""
        img = self._start_image(shape, condition_x, self.num_timesteps - generation_start_steps, generation_start_steps)
        image_list = [img.clone().cpu()] if with_images else None
        x0_list = [img.clone().cpu()] if with_x0_images else None       # the reference's typo (`img.clne()`, :1538) aside
        steps = reversed(range(0, self.num_timesteps))
        for i, t in enumerate(self._iter(steps, self.num_timesteps)):
            if i < generation_start_steps:
                continue
            cs = 1.0 if i < guidance_start_steps else cond_scale
            ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
            img, x_start = self.p_sample(img, t, condition_x, class_label, cs, ccs)
            if with_images:
                image_list.append(img.clone().cpu())
            if with_x0_images:
                x0_list.append(x_start.clone().cpu())
        out = self.unnormalize(img)
        if with_images:
            return (out, image_list, x0_list) if with_x0_images else (out, image_list)
        return out

    @torch.inference_mode()
    def ddim_sample(self, shape, condition_x, class_label, cond_scale, guidance_start_steps, class_cond_scale,
                    class_guidance_start_steps, generation_start_steps, sampling_timesteps, with_images, with_x0_images):
        """model.py:1566-1641"""
        times = torch.linspace(-1, self.num_timesteps - 1, steps=sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        time_pairs = list(zip(times[:-1], times[1:]))
        img = self._start_image(shape, condition_x,
                                time_pairs[generation_start_steps][0] if generation_start_steps > 0 else 0,
                                generation_start_steps)
        image_list = [img.clone().cpu()] if with_images else None
        x0_list = [img.clone().cpu()] if with_x0_images else None
        for i, (time, time_next) in enumerate(self._iter(time_pairs, len(time_pairs))):
            if i < generation_start_steps:
                continue
            cs = 1.0 if i < guidance_start_steps else cond_scale
            ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
            x = img.contiguous().float()
            out_c, out_n, scale = self._predict(x, time, condition_x, class_label, cs, ccs)
            if time_next < 0:
                s = self._scalars(time, _lib.GAUSS_DDIM_LAST, True, True, scale)
                img, x_start, _ = self._update(x, out_c, out_n, None, s, want_x0=False)
                x_start = img
            else:
                s = self._scalars(time, _lib.GAUSS_DDIM, True, True, scale, time_next)
                noise = self._randn(x.shape, x.device)                                      # model.py:1618
                img, x_start, _ = self._update(x, out_c, out_n, noise, s, want_x0=with_x0_images)
            if with_images:
                image_list.append(img.clone().cpu())
            if with_x0_images:
                x0_list.append(x_start.clone().cpu())
        out = self.unnormalize(img)
        if with_images:
            return (out, image_list, x0_list) if with_x0_images else (out, image_list)
        return out

    @torch.inference_mode()
    def sample(self, batch_size=16, condition_x=None, class_label=None, cond_scale=1.0, guidance_start_steps=0,
               class_cond_scale=1.0, class_guidance_start_steps=0, generation_start_steps=0, num_sample_steps=None,
               with_images=False, with_x0_images=False):
        """model.py:1645-1660"""
        sampling_timesteps = self.sampling_timesteps if num_sample_steps is None else num_sample_steps
        _lib.require_cuda(condition_x, "sample")
        _n, _c, h, w = condition_x.shape
        condition_x = (condition_x * 2 - 1).contiguous().float()                            # normalize_to_neg_one_to_one
        sample_fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return sample_fn((batch_size, self.channels, h, w), condition_x, class_label, cond_scale, guidance_start_steps,
                         class_cond_scale, class_guidance_start_steps, generation_start_steps, sampling_timesteps,
                         with_images, with_x0_images)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("srgd_b200 implements the sampling path only (the reference ships no trainer)")
