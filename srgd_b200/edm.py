"""Host-side mirror of the reference's `ConditionalElucidatedDiffusionSR` (model.py:2059-2560): the EDM sampler
family (Karras et al. 2022) on the same class- and LR-conditioned U-Net -- SURVEY.md section 8 f-4.

Kept from the reference: constructor keywords, attributes, `preconditioned_network_forward`, `get_noised_images`,
`sample` (-> `sample_org`, stochastic Heun, or `sample_using_dpmpp`, DPM-Solver++ 2M, by `use_dpmpp_solver`),
`tiled_sample`, the RNG call order on torch's generator, the `NotImplementedError` for two guidance scales, and the
state-dict layout (`net.<U-Net keys>`).  The preconditioning coefficients and the sigma schedule come from the pip
package's `ElucidatedDiffusion` base class in the reference; they are restated here from the published algorithm
(c_in, c_out, c_skip, c_noise, rho-schedule with a trailing zero).

Underneath: the U-Net forward is `srgd_unet_forward` with the class-guidance pair as ONE 2x batch (the reference
runs two forwards, model.py:2140-2178); everything elementwise is three fused kernels (`srgd_edm_perturb`,
`srgd_edm_update`, `srgd_edm_dpmpp`, csrc/edm.cu) that also emit the scaled input of the next U-Net evaluation.
No shipped configuration or weights select this family (conf.model == 'conditional_elucidated').
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .tiling import TilePlan

try:
    from tqdm import tqdm as _tqdm
except Exception:                      # pragma: no cover
    _tqdm = None


class ConditionalElucidatedDiffusionSR(nn.Module):
    def __init__(self, net, *, image_size, channels=3, num_sample_steps=32, sigma_min=0.002, sigma_max=80,
                 sigma_data=0.5, rho=7, P_mean=-1.2, P_std=1.2, S_churn=80, S_tmin=0.05, S_tmax=50, S_noise=1.003,
                 cond_drop_prob=0., class_cond_drop_prob=0., use_dpmpp_solver=False, loss_type='l2'):
        super().__init__()
        assert net.random_or_learned_sinusoidal_cond
        self.self_condition = net.self_condition
        self.net = net
        self.channels, self.image_size = channels, image_size
        self.sigma_min, self.sigma_max, self.sigma_data, self.rho = sigma_min, sigma_max, sigma_data, rho
        self.P_mean, self.P_std = P_mean, P_std
        self.num_sample_steps = num_sample_steps
        self.S_churn, self.S_tmin, self.S_tmax, self.S_noise = S_churn, S_tmin, S_tmax, S_noise
        self.cond_drop_prob, self.class_cond_drop_prob = cond_drop_prob, class_cond_drop_prob
        self.use_dpmpp_solver = use_dpmpp_solver
        self.loss_type = loss_type
        self.progress = True
        self.rng_device = None          # "cpu": draw from torch's global CPU generator (parity runs vs a CPU reference)
        self.last_step_launches = 0

    # -- pip base class (restated; parity unpinned) ----------------------------------------------------------------
    @property
    def device(self):
        return next(self.net.parameters()).device

    def set_seed(self, seed):
        torch.cuda.manual_seed(seed)

    def c_skip(self, sigma):
        return (self.sigma_data ** 2) / (sigma ** 2 + self.sigma_data ** 2)

    def c_out(self, sigma):
        return sigma * self.sigma_data * (self.sigma_data ** 2 + sigma ** 2) ** -0.5

    def c_in(self, sigma):
        return 1 * (sigma ** 2 + self.sigma_data ** 2) ** -0.5

    def c_noise(self, sigma):
        return torch.log(sigma.clamp(min=1e-20)) * 0.25

    def sample_schedule(self, num_sample_steps=None):
        """Host fp32 copy of the schedule (the reference builds it on the device and reads it back with .item())."""
        n = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        inv_rho = 1 / self.rho
        steps = torch.arange(n, dtype=torch.float32)
        sigmas = (self.sigma_max ** inv_rho + steps / (n - 1) * (self.sigma_min ** inv_rho - self.sigma_max ** inv_rho)) ** self.rho
        return F.pad(sigmas, (0, 1), value=0.)

    # -- helpers -----------------------------------------------------------------------------------------------------
    def _coeffs(self, sigma: float):
        """(c_in, c_out, c_skip, c_noise) at a scalar sigma with the reference's fp32 tensor arithmetic
        (`torch.full((b,), sigma)` then the base-class formulas, model.py:2134-2147)."""
        s = torch.full((1,), sigma, dtype=torch.float32)
        return float(self.c_in(s)), float(self.c_out(s)), float(self.c_skip(s)), float(self.c_noise(s))

    def _randn(self, shape, device):
        if self.rng_device is not None and torch.device(self.rng_device).type == "cpu":
            return torch.randn(tuple(shape)).to(device)
        return torch.randn(tuple(shape), device=device)

    def _iter(self, n):
        rng = range(n)
        if self.progress and _tqdm is not None:
            return _tqdm(rng, desc='sampling time step', total=n)
        return rng

    def _net(self, x_in, sigma: float, condition_x, class_label, cond_scale, class_cond_scale):
        """The U-Net evaluation(s) of preconditioned_network_forward on the ALREADY SCALED input x_in = c_in * x:
        returns (net_cond, net_null or None, guidance scale)."""
        if (cond_scale != 1.0) and (class_cond_scale != 1.0):
            raise NotImplementedError(
                "Currently, you cannot specify both cond_scale and class_cond_scale at the same time.")
        unet, B, dev = self.net, x_in.shape[0], x_in.device
        if cond_scale != 1.0:
            rows, n_cond, scale = 2 * B, B, cond_scale
            labels = unet.labels_for(class_label, B, dev)
            labels = None if labels is None else torch.cat((labels, labels))
        elif class_cond_scale != 1.0:
            rows, n_cond, scale = 2 * B, 2 * B, class_cond_scale
            labels = unet.labels_for(class_label, B, dev)
            labels = None if labels is None else torch.cat((labels, torch.full_like(labels, -1)))
        else:
            rows, n_cond, scale = B, B, 1.0
            labels = unet.labels_for(class_label, B, dev)
        if condition_x is None:
            n_cond = 0
        c_noise = self._coeffs(sigma)[3]
        t = torch.full((rows,), c_noise, device=dev, dtype=torch.float32)
        out = unet.run(x_in, t, labels, condition_x, rows, n_cond)
        self.last_step_launches += unet.last_launches
        return (out, None, 1.0) if rows == B else (out[:B], out[B:], scale)

    def _update(self, x_eval, net_c, net_n, scal: _lib.EdmScalars, x_base=None, d_prev=None, want_images=False,
                want_d=False, want_denoised=False, want_xin=False):
        lib = _lib.load()
        mk = lambda want: torch.empty_like(x_eval) if want else None
        images, d, den, xin = mk(want_images), mk(want_d), mk(want_denoised), mk(want_xin and want_images)
        with torch.cuda.device(x_eval.device):
            rc = lib.srgd_edm_update(_lib.ptr(x_eval), _lib.ptr(net_c), _lib.ptr(net_n), _lib.ptr(x_base),
                                     _lib.ptr(d_prev), _lib.ptr(images), _lib.ptr(d), _lib.ptr(den), _lib.ptr(xin),
                                     x_eval.numel(), C.byref(scal), _lib.current_stream())
        _lib.check(rc, "srgd_edm_update")
        self.last_step_launches += 1
        return images, d, den, xin

    def _perturb(self, images, noise, coef: float, c_in: float, want_xin=True):
        lib = _lib.load()
        hat = torch.empty_like(images)
        xin = torch.empty_like(images) if want_xin else None
        with torch.cuda.device(images.device):
            rc = lib.srgd_edm_perturb(_lib.ptr(images), _lib.ptr(noise), float(self.S_noise), float(coef), float(c_in),
                                      _lib.ptr(hat), _lib.ptr(xin), images.numel(), _lib.current_stream())
        _lib.check(rc, "srgd_edm_perturb")
        self.last_step_launches += 1
        return hat, xin

    def _finalize(self, img):
        out = torch.empty_like(img)
        with torch.cuda.device(img.device):
            rc = _lib.load().srgd_finalize_image(_lib.ptr(img.contiguous()), _lib.ptr(out), out.numel(),
                                                 _lib.current_stream())
        _lib.check(rc, "srgd_finalize_image")
        return out

    # -- reference surface -------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def preconditioned_network_forward(self, noised_images, sigma, condition_x, class_label, cond_scale=1.0,
                                       class_cond_scale=1.0, clamp=False):
        """model.py:2128-2183 for a scalar sigma (the sampling paths; per-sample sigmas are a training-time call)."""
        _lib.require_cuda(noised_images, "preconditioned_network_forward")
        if torch.is_tensor(sigma):
            if sigma.numel() != 1:
                raise NotImplementedError("per-sample sigmas are only used by the training loss (not shipped)")
            sigma = float(sigma)
        x = noised_images.contiguous().float()
        c_in, c_out, c_skip, _ = self._coeffs(sigma)
        _, x_in = self._perturb(x, None, 0.0, c_in)
        net_c, net_n, scale = self._net(x_in, sigma, condition_x, class_label, cond_scale, class_cond_scale)
        s = _lib.EdmScalars(c_skip, c_out, scale, 1.0, 0.0, 0.0, int(bool(clamp)))
        return self._update(x, net_c, net_n, s, want_denoised=True)[2]

    @torch.inference_mode()
    def get_noised_images(self, condition_x, target_step, num_sample_steps=None):
        """model.py:2186-2195: condition_x (already in [-1,1]) + sigma[target_step] * noise."""
        sigmas = self.sample_schedule(num_sample_steps)
        noise = self._randn(condition_x.shape, condition_x.device)
        hat, _ = self._perturb_raw(condition_x.contiguous().float(), noise, float(sigmas[target_step]))
        return hat

    def _perturb_raw(self, base, noise, sigma: float):
        """base + sigma * noise (no S_noise factor)."""
        lib = _lib.load()
        out = torch.empty_like(base)
        with torch.cuda.device(base.device):
            rc = lib.srgd_edm_perturb(_lib.ptr(base), _lib.ptr(noise), 1.0, float(sigma), 0.0, _lib.ptr(out), None,
                                      base.numel(), _lib.current_stream())
        _lib.check(rc, "srgd_edm_perturb")
        return out, None

    def _heun(self, x_hat, x_in, sigma_hat, sigma_next, condition_x, class_label, cs, ccs, clamp):
        """One Heun step from images_hat (model.py:2276-2289): returns (images_next, x0-side record)."""
        c_in, c_out, c_skip, _ = self._coeffs(sigma_hat)
        net_c, net_n, scale = self._net(x_in, sigma_hat, condition_x, class_label, cs, ccs)
        last = sigma_next == 0
        c_in_n = 0.0 if last else self._coeffs(sigma_next)[0]
        s = _lib.EdmScalars(c_skip, c_out, scale, sigma_hat, sigma_next - sigma_hat, c_in_n, int(bool(clamp)))
        nxt, d, _, xin2 = self._update(x_hat, net_c, net_n, s, x_base=x_hat, want_images=True, want_d=True,
                                       want_xin=not last)
        if last:
            return nxt, d
        _, c_out2, c_skip2, _ = self._coeffs(sigma_next)
        net_c, net_n, scale = self._net(xin2, sigma_next, condition_x, class_label, cs, ccs)
        s2 = _lib.EdmScalars(c_skip2, c_out2, scale, sigma_next, 0.5 * (sigma_next - sigma_hat), 0.0, int(bool(clamp)))
        out, d2, _, _ = self._update(nxt, net_c, net_n, s2, x_base=x_hat, d_prev=d, want_images=True, want_d=True)
        return out, d2

    def _schedule(self, num_sample_steps):
        sigmas = self.sample_schedule(num_sample_steps)
        gammas = torch.where((sigmas >= self.S_tmin) & (sigmas <= self.S_tmax),
                             min(self.S_churn / num_sample_steps, math.sqrt(2) - 1), 0.)       # model.py:2234-2238
        return sigmas, gammas

    def _init_images(self, shape, condition_x, sigmas, generation_start_steps, zero_init, dev):
        if generation_start_steps > 0:
            # like the reference, on the schedule of self.num_sample_steps (model.py:2241 passes no step count)
            return self.get_noised_images(condition_x, generation_start_steps)
        if zero_init:
            return torch.zeros(shape, device=dev)
        z = self._randn(shape, dev)
        return self._perturb_raw(torch.zeros_like(z), z, float(sigmas[0]))[0]       # init_sigma * randn (model.py:2247)

    @torch.inference_mode()
    def sample(self, batch_size=16, condition_x=None, class_label=None, cond_scale=1.0, guidance_start_steps=0,
               class_cond_scale=1.0, class_guidance_start_steps=0, generation_start_steps=0, num_sample_steps=None,
               clamp=True, with_images=False, with_x0_images=False, zero_init=False):
        fn = self.sample_using_dpmpp if self.use_dpmpp_solver else self.sample_org              # model.py:2201-2211
        return fn(batch_size, condition_x, class_label, cond_scale, guidance_start_steps, class_cond_scale,
                  class_guidance_start_steps, generation_start_steps, num_sample_steps, clamp, with_images,
                  with_x0_images, zero_init)

    @torch.inference_mode()
    def sample_org(self, batch_size=16, condition_x=None, class_label=None, cond_scale=1.0, guidance_start_steps=0,
                   class_cond_scale=1.0, class_guidance_start_steps=0, generation_start_steps=0, num_sample_steps=None,
                   clamp=True, with_images=False, with_x0_images=False, zero_init=False):
        """Stochastic Heun sampler, model.py:2213-2307."""
        num_sample_steps = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        _lib.require_cuda(condition_x, "sample")
        _n, _c, h, w = condition_x.shape
        shape, dev = (batch_size, self.channels, h, w), condition_x.device
        condition_x = (condition_x * 2 - 1).contiguous().float()
        sigmas, gammas = self._schedule(num_sample_steps)
        images = self._init_images(shape, condition_x, sigmas, generation_start_steps, zero_init, dev)
        image_list = [images.clone().cpu()] if with_images else None
        x0_list = [images.clone().cpu()] if with_x0_images else None
        for i in self._iter(num_sample_steps):
            if i < generation_start_steps:
                continue
            cs = 1.0 if i < guidance_start_steps else cond_scale
            ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
            sigma, sigma_next, gamma = sigmas[i].item(), sigmas[i + 1].item(), gammas[i].item()
            self.last_step_launches = 0
            noise = self._randn(shape, dev)                                           # S_noise * randn, model.py:2270
            sigma_hat = sigma + gamma * sigma
            x_hat, x_in = self._perturb(images, noise, math.sqrt(sigma_hat ** 2 - sigma ** 2), self._coeffs(sigma_hat)[0])
            images, x0 = self._heun(x_hat, x_in, sigma_hat, sigma_next, condition_x, class_label, cs, ccs, clamp)
            if with_images:
                image_list.append(images.clone().cpu())
            if with_x0_images:
                x0_list.append(x0.clone().cpu())
        out = self._finalize(images)
        if with_images:
            return (out, image_list, x0_list) if with_x0_images else (out, image_list)
        return out

    @torch.inference_mode()
    def sample_using_dpmpp(self, batch_size=16, condition_x=None, class_label=None, cond_scale=1.0,
                           guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
                           generation_start_steps=0, num_sample_steps=None, clamp=True, with_images=False,
                           with_x0_images=False, zero_init=False):
        """DPM-Solver++ (2M), model.py:2466-2544."""
        num_sample_steps = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        _lib.require_cuda(condition_x, "sample")
        _n, _c, h, w = condition_x.shape
        shape, dev = (batch_size, self.channels, h, w), condition_x.device
        condition_x = (condition_x * 2 - 1).contiguous().float()
        sigmas = self.sample_schedule(num_sample_steps)
        images = self._init_images(shape, condition_x, sigmas, generation_start_steps, zero_init, dev)
        image_list = [images.clone().cpu()] if with_images else None
        x0_list = [images.clone().cpu()] if with_x0_images else None
        t_fn = lambda s: s.log().neg()
        sigma_fn = lambda t: t.neg().exp()
        lib = _lib.load()
        old = None
        x_in = None
        for i in self._iter(len(sigmas) - 1):
            if i < generation_start_steps:
                continue
            cs = 1.0 if i < guidance_start_steps else cond_scale
            ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
            sigma = sigmas[i].item()
            self.last_step_launches = 0
            c_in, c_out, c_skip, _ = self._coeffs(sigma)
            if x_in is None:
                _, x_in = self._perturb(images, None, 0.0, c_in)
            net_c, net_n, scale = self._net(x_in, sigma, condition_x, class_label, cs, ccs)
            den = self._update(images, net_c, net_n, _lib.EdmScalars(c_skip, c_out, scale, 1.0, 0.0, 0.0, int(bool(clamp))),
                               want_denoised=True)[2]
            t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
            hh = t_next - t
            if old is None or sigmas[i + 1] == 0:
                w_new, w_old, old_arg = 1.0, 0.0, None
            else:
                r = (t - t_fn(sigmas[i - 1])) / hh
                g = -1 / (2 * r)
                w_new, w_old, old_arg = float(1 - g), float(g), old
            a, b = float(sigma_fn(t_next) / sigma_fn(t)), float((-hh).expm1())
            last = float(sigmas[i + 1]) == 0.0
            nxt = torch.empty_like(images)
            x_in = None if last else torch.empty_like(images)
            with torch.cuda.device(dev):
                rc = lib.srgd_edm_dpmpp(_lib.ptr(images), _lib.ptr(den), _lib.ptr(old_arg), a, b, w_new, w_old,
                                        0.0 if last else self._coeffs(sigmas[i + 1].item())[0], _lib.ptr(nxt),
                                        _lib.ptr(x_in), images.numel(), _lib.current_stream())
            _lib.check(rc, "srgd_edm_dpmpp")
            if with_x0_images:
                dd = den if old_arg is None else w_new * den + w_old * old
                x0_list.append(dd.clone().cpu())
            images, old = nxt, den
            if with_images:
                image_list.append(images.clone().cpu())
        out = self._finalize(images)
        if with_images:
            return (out, image_list, x0_list) if with_x0_images else (out, image_list)
        return out

    @torch.inference_mode()
    def tiled_sample(self, batch_size=4, tile_size=256, tile_stride=256, condition_x=None, class_label=None,
                     cond_scale=1.0, guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
                     generation_start_steps=0, num_sample_steps=None, clamp=True, zero_init=False, with_images=False,
                     with_x0_images=False, start_white_noise=True, amp=False):
        """Large images with the Heun sampler, model.py:2309-2462: the whole canvas is perturbed first, the tiles of the
        step's grid are denoised from images_hat in minibatches of `batch_size`, and after odd steps everything outside
        the hull of the shifted grid is replaced by sigma_i * noise (get_noised_images(zeros, i))."""
        from .tiled import CudaTiledOps
        num_sample_steps = self.num_sample_steps if num_sample_steps is None else num_sample_steps
        _lib.require_cuda(condition_x, "tiled_sample")
        condition_x = condition_x * 2 - 1
        batch, ch, h, w = condition_x.shape
        assert batch == 1, "the reference's tiled_sample works on one image (model.py:2391)"
        plan = TilePlan(h, w, tile_size, tile_stride)
        condition_x = F.pad(condition_x, plan.canvas_pad, mode='reflect').contiguous().float()
        shape, dev = tuple(condition_x.shape), condition_x.device
        sigmas, gammas = self._schedule(num_sample_steps)
        images = self._init_images(shape, condition_x, sigmas, generation_start_steps, zero_init, dev)
        top, bottom, left, right = plan.crop
        image_list = [images[:, :, top:bottom, left:right].clone().cpu()] if with_images else None
        x0_list = [images[:, :, top:bottom, left:right].clone().cpu()] if with_x0_images else None
        it, ib, il, ir = plan.inner
        cond_canvas = torch.zeros_like(condition_x)
        cond_canvas[:, :, it:ib, il:ir] = condition_x[:, :, it:ib, il:ir]
        ops = CudaTiledOps(self)
        x_start = images.clone() if with_x0_images else None
        cond_tiles = {}
        for i in self._iter(num_sample_steps):
            if i < generation_start_steps:
                continue
            cs = 1.0 if i < guidance_start_steps else cond_scale
            ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
            sigma, sigma_next, gamma = sigmas[i].item(), sigmas[i + 1].item(), gammas[i].item()
            self.last_step_launches = 0
            noise = self._randn(shape, dev)
            sigma_hat = sigma + gamma * sigma
            c_in = self._coeffs(sigma_hat)[0]
            images_hat, xin_canvas = self._perturb(images, noise, math.sqrt(sigma_hat ** 2 - sigma ** 2), c_in)
            tiles = plan.grids[i % 2]
            for s0 in range(0, len(tiles), batch_size):
                chunk = tiles[s0:s0 + batch_size]
                key = (i % 2, s0)
                if key not in cond_tiles:
                    cond_tiles[key] = ops.gather(cond_canvas, chunk, tile_size)
                x_hat = ops.gather(images_hat, chunk, tile_size)
                x_in = ops.gather(xin_canvas, chunk, tile_size)
                nxt, x0 = self._heun(x_hat, x_in, sigma_hat, sigma_next, cond_tiles[key], class_label, cs, ccs, clamp)
                if plan.disjoint:
                    ops.scatter(images, chunk, nxt, tile_size)
                    if x_start is not None:
                        ops.scatter(x_start, chunk, x0, tile_size)
                else:                # overlapping tiles (tile_stride < tile_size): the later tile wins (model.py:2436)
                    for j, c in enumerate(chunk):
                        ops.scatter(images, [c], nxt[j:j + 1], tile_size)
                        if x_start is not None:
                            ops.scatter(x_start, [c], x0[j:j + 1], tile_size)
            if i % 2 == 1:
                # get_noised_images(zeros, i), model.py:2448: sigma_i of the DEFAULT schedule (self.num_sample_steps)
                fresh = self._randn(shape, dev)
                ops.renoise_outside(images, fresh, float(self.sample_schedule()[i]), plan.inner)
            if with_images:
                image_list.append(images.clone().cpu())
            if with_x0_images:
                x0_list.append(x_start.clone().cpu())
        out = self._finalize(images[:, :, top:bottom, left:right].contiguous())
        if with_images:
            return (out, image_list, x0_list) if with_x0_images else (out, image_list)
        return out

    def forward(self, *args, **kwargs):
        raise NotImplementedError("srgd_b200 implements the sampling path only (the reference ships no trainer)")
