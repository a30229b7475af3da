"""Host-side mirror of the reference's `ConditionalContinuousTimeGaussianDiffusionSR`
(model.py:3054-3495, sampling half) on top of the srgd_b200 CUDA library.

Kept from the reference: constructor keywords, attributes, and the call surface
`p_mean_variance / p_sample / p_sample_loop / sample / tiled_sample / q_sample` with the same
argument meaning, value ranges, RNG call order (torch's generator, same shapes in the same order)
and error behaviour.  What changed underneath:

  * classifier-free guidance runs as ONE 2x-batch U-Net launch sequence (cond rows | null rows)
    instead of two sequential forwards (model.py:3148-3154);
  * guidance combine + x0 + clamp + posterior mean + noise add are one fused kernel
    (`srgd_sampler_step`) instead of ~15 elementwise kernels (model.py:3150-3168, 3187-3188);
  * the per-step schedule scalars are evaluated on the host (no device sync for `time_next == 0`,
    model.py:3184) with the same fp32 tensor ops as the reference (model.py:3127-3134).
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .tiling import TilePlan

try:                                   # progress bars are cosmetic (reference uses tqdm)
    from tqdm import tqdm as _tqdm
except Exception:                      # pragma: no cover
    _tqdm = None


def _log(t: torch.Tensor, eps: float = 1e-20) -> torch.Tensor:
    return torch.log(t.clamp(min=eps))


def beta_linear_log_snr(t: torch.Tensor) -> torch.Tensor:
    """-log(expm1(1e-4 + 10 t^2))   (model.py:2632-2633)"""
    return -_log(torch.special.expm1(1e-4 + 10 * (t ** 2)))


def alpha_cosine_log_snr(t: torch.Tensor, s: float = 0.008) -> torch.Tensor:
    """model.py:2635-2636"""
    return -_log((torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2) - 1, eps=1e-5)


def _host_scalar(t) -> torch.Tensor:
    """0-dim fp32 CPU tensor from a python float or a (possibly CUDA) tensor."""
    if torch.is_tensor(t):
        return t.detach().to(device="cpu", dtype=torch.float32).reshape(())
    return torch.tensor(float(t), dtype=torch.float32)


class ConditionalContinuousTimeGaussianDiffusionSR(nn.Module):
    def __init__(self, model, *, image_size, channels=3, noise_schedule='linear', num_sample_steps=500,
                 clip_sample_denoised=True, learned_schedule_net_hidden_dim=1024,
                 learned_noise_schedule_frac_gradient=1., min_snr_loss_weight=False, min_snr_gamma=5,
                 cond_drop_prob=0., class_cond_drop_prob=0., loss_type='l2'):
        super().__init__()
        assert model.random_or_learned_sinusoidal_cond
        self.model = model
        self.channels = channels
        self.image_size = image_size
        if noise_schedule == 'linear':
            self.log_snr = beta_linear_log_snr
        elif noise_schedule == 'cosine':
            self.log_snr = alpha_cosine_log_snr
        elif noise_schedule == 'learned':
            raise NotImplementedError("the learned noise schedule is a training-time feature; no shipped "
                                      "configuration uses it")
        else:
            raise ValueError(f'unknown noise schedule {noise_schedule}')
        self.num_sample_steps = num_sample_steps
        self.clip_sample_denoised = clip_sample_denoised
        self.min_snr_loss_weight = min_snr_loss_weight
        self.min_snr_gamma = min_snr_gamma
        self.cond_drop_prob = cond_drop_prob
        self.class_cond_drop_prob = class_cond_drop_prob
        self.loss_type = loss_type
        self.progress = True
        self.last_step_launches = 0
        # None: noise is drawn by the generator of the state's device (what the reference does on a GPU).
        # "cpu": every draw comes from torch's global CPU generator and is copied over -- same shapes, same order,
        # so a run reproduces the noise stream of the reference run on the CPU with the same seed (parity tests).
        self.rng_device = None

    @property
    def device(self):
        return next(self.model.parameters()).device

    def set_seed(self, seed):
        torch.cuda.manual_seed(seed)

    def _randn(self, shape, device):
        if self.rng_device is not None and torch.device(self.rng_device).type == "cpu":
            return torch.randn(tuple(shape)).to(device, non_blocking=False)
        return torch.randn(tuple(shape), device=device)

    # ---------------------------------------------------------------------------------------
    # per-step scalars (model.py:3127-3134, 3168) -- same fp32 tensor ops, evaluated on the host
    # ---------------------------------------------------------------------------------------
    def step_scalars(self, time, time_next, guidance_scale: float = 1.0) -> _lib.StepScalars:
        t, tn = _host_scalar(time), _host_scalar(time_next)
        log_snr, log_snr_next = self.log_snr(t), self.log_snr(tn)
        c = -torch.special.expm1(log_snr - log_snr_next)
        alpha, sigma = log_snr.sigmoid().sqrt(), (-log_snr).sigmoid().sqrt()
        alpha_next = log_snr_next.sigmoid().sqrt()
        var = (-log_snr_next).sigmoid() * c
        s = _lib.StepScalars()
        s.alpha, s.sigma, s.alpha_next, s.c = float(alpha), float(sigma), float(alpha_next), float(c)
        s.noise_scale = float(var.sqrt()) if float(tn) != 0.0 else 0.0
        s.guidance_scale = float(guidance_scale)
        s.clip = int(bool(self.clip_sample_denoised))
        s._log_snr = float(log_snr)
        s._var = float(var)
        return s

    # ---------------------------------------------------------------------------------------
    # denoise with classifier-free guidance (model.py:3136-3158)
    # ---------------------------------------------------------------------------------------
    def _predict(self, x, log_snr: float, condition_x, class_label, cond_scale, class_cond_scale):
        """Returns (eps_cond, eps_null or None, guidance scale)."""
        if (cond_scale != 1.0) and (class_cond_scale != 1.0):
            raise NotImplementedError(
                "Currently, you cannot specify both cond_scale and class_cond_scale at the same time.")
        unet = self.model
        B = x.shape[0]
        dev = x.device
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
        lsnr = torch.full((rows,), log_snr, device=dev, dtype=torch.float32)     # model.py:3136
        eps = unet.run(x, lsnr, labels, condition_x, rows, n_cond)
        self.last_step_launches = unet.last_launches
        if rows == B:
            return eps, None, 1.0
        return eps[:B], eps[B:], scale

    def _update(self, x, eps_cond, eps_null, noise, s: _lib.StepScalars, want_x0=True):
        lib = _lib.load()
        img = torch.empty_like(x)
        x0 = torch.empty_like(x) if want_x0 else None
        with torch.cuda.device(x.device):
            rc = lib.srgd_sampler_step(_lib.ptr(x), _lib.ptr(eps_cond), _lib.ptr(eps_null), _lib.ptr(noise),
                                       _lib.ptr(img), _lib.ptr(x0), x.numel(), C.byref(s), _lib.current_stream())
        _lib.check(rc, "srgd_sampler_step")
        self.last_step_launches += 1
        return img, x0

    def p_mean_variance(self, x, time, condition_x, class_label, cond_scale, class_cond_scale, time_next):
        _lib.require_cuda(x, "p_mean_variance")
        x = x.contiguous().float()
        s = self.step_scalars(time, time_next)
        eps_c, eps_n, scale = self._predict(x, s._log_snr, condition_x, class_label, cond_scale, class_cond_scale)
        s.guidance_scale = scale
        mean, x_start = self._update(x, eps_c, eps_n, None, s)
        return mean, torch.tensor(s._var, device=x.device, dtype=torch.float32), x_start

    @torch.inference_mode()
    def p_sample(self, x, time, condition_x, class_label, cond_scale, class_cond_scale, time_next, noise=None):
        _lib.require_cuda(x, "p_sample")
        x = x.contiguous().float()
        s = self.step_scalars(time, time_next)
        eps_c, eps_n, scale = self._predict(x, s._log_snr, condition_x, class_label, cond_scale, class_cond_scale)
        s.guidance_scale = scale
        if float(_host_scalar(time_next)) == 0.0:                              # model.py:3184
            return self._update(x, eps_c, eps_n, None, s)
        if noise is None:
            noise = self._randn(x.shape, x.device)                              # model.py:3187
        else:
            _lib.require_like(noise, x, "p_sample noise")
        return self._update(x, eps_c, eps_n, noise.contiguous().float(), s)

    # ---------------------------------------------------------------------------------------
    # q_sample (model.py:3434-3447)
    # ---------------------------------------------------------------------------------------
    def q_sample(self, x_start, times, noise=None, return_alpha_sigma_sum=False):
        if noise is None:
            noise = self._randn(x_start.shape, x_start.device)
        else:
            _lib.require_like(noise, x_start, "q_sample noise")
        times_t = times if torch.is_tensor(times) else torch.tensor(times)
        log_snr = self.log_snr(times_t.float())
        if x_start.is_cuda and log_snr.numel() == 1:
            lsn = _host_scalar(log_snr)
            alpha, sigma = float(lsn.sigmoid().sqrt()), float((-lsn).sigmoid().sqrt())
            out = torch.empty_like(x_start, dtype=torch.float32)
            with torch.cuda.device(x_start.device):
                rc = _lib.load().srgd_q_sample(_lib.ptr(x_start.contiguous().float()),
                                               _lib.ptr(noise.contiguous().float()), _lib.ptr(out), out.numel(),
                                               alpha, sigma, _lib.current_stream())
            _lib.check(rc, "srgd_q_sample")
            log_snr = log_snr.to(x_start.device)
            if return_alpha_sigma_sum:
                return out, torch.tensor(alpha + sigma, device=x_start.device)
            return out, log_snr
        # per-sample times (training-style call): tiny broadcast, left to torch
        log_snr = log_snr.to(x_start.device)
        pad = log_snr.reshape(*log_snr.shape, *((1,) * max(0, x_start.ndim - log_snr.ndim)))
        alpha, sigma = pad.sigmoid().sqrt(), (-pad).sigmoid().sqrt()
        x_noised = x_start * alpha + noise * sigma
        return (x_noised, alpha + sigma) if return_alpha_sigma_sum else (x_noised, log_snr)

    def _pure_noise_at(self, noise, time):
        """q_sample(zeros, time) without reading the zeros: sigma(time) * noise (model.py:3395, 3442)."""
        lsn = self.log_snr(_host_scalar(time))
        out = torch.empty_like(noise)
        with torch.cuda.device(noise.device):
            rc = _lib.load().srgd_q_sample(None, _lib.ptr(noise), _lib.ptr(out), out.numel(),
                                           float(lsn.sigmoid().sqrt()), float((-lsn).sigmoid().sqrt()),
                                           _lib.current_stream())
        _lib.check(rc, "srgd_q_sample")
        return out

    def _finalize(self, img):
        out = torch.empty_like(img)
        with torch.cuda.device(img.device):
            rc = _lib.load().srgd_finalize_image(_lib.ptr(img.contiguous()), _lib.ptr(out), out.numel(),
                                                 _lib.current_stream())
        _lib.check(rc, "srgd_finalize_image")
        return out

    def _iter(self, n):
        rng = range(n)
        if self.progress and _tqdm is not None:
            return _tqdm(rng, desc='sampling loop time step', total=n)
        return rng

    # ---------------------------------------------------------------------------------------
    # batched sampling (model.py:3191-3246, 3417-3430)
    # ---------------------------------------------------------------------------------------
    def p_sample_loop(self, shape, condition_x, class_label, cond_scale, guidance_start_steps, class_cond_scale,
                      class_guidance_start_steps, generation_start_steps, num_sample_steps, with_images,
                      with_x0_images):
        batch = shape[0]
        dev = self.device
        if generation_start_steps > 0:
            start = 1. - torch.tensor(generation_start_steps / num_sample_steps)       # fp32, model.py:3199
            img, _ = self.q_sample(condition_x, start)
        else:
            img = self._randn(shape, dev)                                       # RNG draw #0 (model.py:3203)
        images = [img.clone().cpu()] if with_images else None
        x0_images = [img.clone().cpu()] if with_x0_images else None
        steps = torch.linspace(1., 0., num_sample_steps + 1)                    # host copy of model.py:3213
        for i in self._iter(num_sample_steps):
            if i < generation_start_steps:
                continue
            cs = 1.0 if i < guidance_start_steps else cond_scale
            ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
            img, x_start = self.p_sample(img, steps[i], condition_x, class_label, cs, ccs, steps[i + 1])
            if with_images:
                images.append(img.clone().cpu())
            if with_x0_images:
                x0_images.append(x_start.clone().cpu())
        img = self._finalize(img)                                               # clamp + [0,1] (3237-3238)
        if with_images:
            return (img, images, x0_images) if with_x0_images else (img, images)
        return img

    def sample(self, batch_size=16, condition_x=None, class_label=None, cond_scale=1.0, guidance_start_steps=0,
               class_cond_scale=1.0, class_guidance_start_steps=0, generation_start_steps=0,
               num_sample_steps=None, with_images=False, with_x0_images=False, x0=None):
        num_sample_steps = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        condition_x = condition_x * 2 - 1                                       # [0,1] -> [-1,1] (model.py:40)
        return self.p_sample_loop((batch_size, self.channels, self.image_size, self.image_size), condition_x,
                                  class_label, cond_scale, guidance_start_steps, class_cond_scale,
                                  class_guidance_start_steps, generation_start_steps, num_sample_steps,
                                  with_images, with_x0_images)

    # ---------------------------------------------------------------------------------------
    # large images: alternating tile grids (model.py:3288-3413)
    # ---------------------------------------------------------------------------------------
    def tiled_sample(self, batch_size=4, tile_size=256, tile_stride=256, condition_x=None, class_label=None,
                     cond_scale=1.0, guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
                     generation_start_steps=0, num_sample_steps=None, with_images=False, with_x0_images=False,
                     start_white_noise=True, amp=False, shard_tiles=False, shard_group=None):
        """`shard_tiles=True` (extension, srgd_b200/tiled.py "exact mode"): the tiles of every step are split into
        contiguous ranges over the ranks of the initialised torch.distributed group (`shard_group`, default: the world)
        and exchanged with one all-gather per step; denoiser calls are regrouped (up to 64 rows instead of
        `batch_size`) with the library in batch-invariant mode, so every rank returns the same image, bit-identical
        for every world size (1 included).  `batch_size` then only shapes the noise draws, as in the reference.

        `condition_x` with a batch of N > 1 same-sized images (extension; the reference's loop only works for N = 1):
        the images advance together, their tiles stacked into one denoiser batch, and share ONE noise stream -- the
        result equals N consecutive single-image calls each preceded by the same reseed, which is what the reference
        CLI does (inference.py:81)."""
        num_sample_steps = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        _lib.require_cuda(condition_x, "tiled_sample")
        condition_x = condition_x * 2 - 1
        batch, ch, h, w = condition_x.shape
        plan = TilePlan(h, w, tile_size, tile_stride)
        condition_x = F.pad(condition_x, plan.canvas_pad, mode='reflect')
        one = (1,) + tuple(condition_x.shape[1:])

        def shared(noise):                                                      # one draw, used by every image
            return noise if batch == 1 else noise.expand(batch, -1, -1, -1).contiguous()

        if generation_start_steps > 0:
            start = 1. - torch.tensor(generation_start_steps / num_sample_steps)       # fp32, model.py:3306
            img, _ = self.q_sample(condition_x, start, noise=shared(self._randn(one, self.device)))
        elif start_white_noise:
            img = shared(self._randn(one, self.device))                         # RNG draw #0 (model.py:3311)
        else:
            img, _ = self.q_sample(condition_x, torch.tensor(1., dtype=torch.float32),
                                   noise=shared(self._randn(one, self.device)))
        top, bottom, left, right = plan.crop
        images = [img[:, :, top:bottom, left:right].clone().cpu()] if with_images else None
        x0_images = [img[:, :, top:bottom, left:right].clone().cpu()] if with_x0_images else None
        steps = torch.linspace(1., 0., num_sample_steps + 1)
        # the LR condition is zeroed outside the hull of the shifted grid (model.py:3337-3342)
        it, ib, il, ir = plan.inner
        cond_canvas = torch.zeros_like(condition_x)
        cond_canvas[:, :, it:ib, il:ir] = condition_x[:, :, it:ib, il:ir]
        x_start = img.clone() if with_x0_images else None
        bar = self._iter(num_sample_steps)
        bar_it = iter(bar)

        def on_step(i, cur, cur_x0):
            next(bar_it, None)                                                  # progress bar tick
            if with_images:                                                     # the reference keeps the
                images.append(cur.clone().cpu())                                # UNCROPPED canvas for every step
            if with_x0_images:                                                  # after the first frame
                x0_images.append(cur_x0.clone().cpu())                          # (model.py:3398-3401)

        from .tiled import CudaTiledOps, run_tiled
        img = img.contiguous()
        cond_canvas = cond_canvas.contiguous()
        img, x_start = run_tiled(CudaTiledOps(self), img, cond_canvas, plan, steps, num_sample_steps, batch_size,
                                 class_label, cond_scale, guidance_start_steps, class_cond_scale,
                                 class_guidance_start_steps, generation_start_steps, x_start=x_start, on_step=on_step,
                                 shard=shard_tiles, group=shard_group)
        for _ in bar_it:
            pass
        img = self._finalize(img[:, :, top:bottom, left:right].contiguous())
        if with_images:
            return (img, images, x0_images) if with_x0_images else (img, images)
        return img

    # training entry points are out of scope (no trainer is shipped with the reference)
    def forward(self, *args, **kwargs):
        raise NotImplementedError("srgd_b200 implements the sampling path only (the reference ships no trainer)")
