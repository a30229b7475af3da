"""CPU oracle for the Real-SRGD sampling hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-PyTorch fp32 *functional* restatement (no nn.Module, no CUDA) of the reference's
conditional U-Net denoiser, classifier-free-guidance dispatch, continuous-time linear-logSNR
posterior update and the two sampling drivers.  Every function cites the reference lines it
follows (paths relative to /root/reference).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this module; the product package
`srgd_b200` never does and fails loudly when its CUDA library is missing.

Parity pinning: the reference ships no tests or golden vectors, so this file is
pinned against outputs of the UNMODIFIED reference `model.py` imported in the build container
behind `oracle/_shim` (three absent pip packages).  The script `tests/golden/make_golden.py`
generated `tests/golden/*.npz`; `tests/test_oracle_golden.py` re-runs this oracle on the same
seeded inputs and compares.  One sub-function is "parity unpinned" at the third-party boundary:
`Attend.forward` of denoising-diffusion-pytorch==1.8.15 (requirements.txt:1) is not under
/root/reference; `_attend` restates its published non-flash algorithm (softmax(q k^T d^-1/2) v).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# architecture description (what get_model builds for conf.model == 'conditional_continuous',
# model.py:3500-3515, with the defaults of config.py)
# --------------------------------------------------------------------------------------------

@dataclass(frozen=True)
class UnetSpec:
    dim: int = 128
    dim_mults: Tuple[int, ...] = (1, 2, 4, 8)
    channels: int = 3
    groups: int = 8
    learned_sinusoidal_dim: int = 32
    heads: int = 4
    dim_head: int = 32
    full_attn: Tuple[bool, ...] = (False, False, False, True)
    num_classes: Optional[int] = 3
    learned_sinusoidal_cond: bool = True     # False: fixed SinusoidalPosEmb(dim) (model.py:600; discrete-time family)

    @property
    def dims(self):
        return [self.dim] + [self.dim * m for m in self.dim_mults]      # model.py:585

    @property
    def in_out(self):
        d = self.dims
        return list(zip(d[:-1], d[1:]))                                  # model.py:586

    @property
    def time_dim(self):
        return self.dim * 4                                              # model.py:592

    @property
    def hidden(self):
        return self.heads * self.dim_head                                # model.py:297, 336


def param_shapes(spec: UnetSpec, prefix: str = "model.") -> "OrderedDict[str, Tuple[int, ...]]":
    """State-dict keys and shapes in registration order (model.py:583-675)."""
    S: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    td, hid = spec.time_dim, spec.hidden

    def conv(name, cout, cin, k):
        S[f"{prefix}{name}.weight"] = (cout, cin, k, k)
        S[f"{prefix}{name}.bias"] = (cout,)

    def linear(name, cout, cin):
        S[f"{prefix}{name}.weight"] = (cout, cin)
        S[f"{prefix}{name}.bias"] = (cout,)

    def resblock(name, cin, cout):                                        # model.py:261-271
        linear(f"{name}.mlp.1", 2 * cout, td)
        for b, ci in (("block1", cin), ("block2", cout)):
            conv(f"{name}.{b}.proj", cout, ci, 3)
            S[f"{prefix}{name}.{b}.norm.weight"] = (cout,)
            S[f"{prefix}{name}.{b}.norm.bias"] = (cout,)
        if cin != cout:
            conv(f"{name}.res_conv", cout, cin, 1)

    def attn(name, c, full):                                              # model.py:287-305, 326-342
        S[f"{prefix}{name}.norm.g"] = (1, c, 1, 1)
        S[f"{prefix}{name}.to_qkv.weight"] = (3 * hid, c, 1, 1)
        if full:
            conv(f"{name}.to_out", c, hid, 1)
        else:
            conv(f"{name}.to_out.0", c, hid, 1)
            S[f"{prefix}{name}.to_out.1.g"] = (1, c, 1, 1)

    conv("init_conv", spec.dim, 2 * spec.channels, 7)                     # model.py:583
    if spec.learned_sinusoidal_cond:
        S[f"{prefix}time_mlp.0.weights"] = (spec.learned_sinusoidal_dim // 2,)  # model.py:231
        linear("time_mlp.1", td, spec.learned_sinusoidal_dim + 1)         # model.py:598, 605
    else:
        linear("time_mlp.1", td, spec.dim)                                # model.py:600-601, 605
    linear("time_mlp.3", td, td)
    if spec.num_classes is not None:                                      # model.py:612-619
        S[f"{prefix}class_mlp.0.weight"] = (spec.num_classes, spec.dim)
        linear("class_mlp.1", td, spec.dim)
        linear("class_mlp.3", td, td)
    n = len(spec.in_out)
    for i, ((di, do), full) in enumerate(zip(spec.in_out, spec.full_attn)):   # model.py:638-648
        resblock(f"downs.{i}.0", di, di)
        resblock(f"downs.{i}.1", di, di)
        attn(f"downs.{i}.2", di, full)
        if i < n - 1:
            conv(f"downs.{i}.3.1", do, 4 * di, 1)                         # Downsample, model.py:106-110
        else:
            conv(f"downs.{i}.3", do, di, 3)
    for i, ((di, do), full) in enumerate(zip(reversed(spec.in_out), reversed(spec.full_attn))):  # model.py:659-669
        resblock(f"ups.{i}.0", do + di, do)
        resblock(f"ups.{i}.1", do + di, do)
        attn(f"ups.{i}.2", do, full)
        if i < n - 1:
            conv(f"ups.{i}.3.net.0", 4 * di, do, 1)                       # PixelShuffleUpsample, model.py:78
        else:
            conv(f"ups.{i}.3", di, do, 3)
    # nn.ModuleList registration order: `downs` and `ups` are both registered (model.py:634-635)
    # before the mid blocks (model.py:651-653), so their keys come first in the state dict.
    mid = spec.dims[-1]
    resblock("mid_block1", mid, mid)
    attn("mid_attn", mid, True)
    resblock("mid_block2", mid, mid)
    resblock("final_res_block", 2 * spec.dim, spec.dim)                   # model.py:674
    conv("final_conv", spec.channels, spec.dim, 1)                        # model.py:675
    return S


def make_state_dict(spec: UnetSpec, seed: int = 1234, prefix: str = "model.", init: str = "unit") -> Dict[str, torch.Tensor]:
    """Deterministic random-init weights of the reference architecture (the shipped .pth is a
    Git-LFS pointer).

    init="unit"  : stress init -- unit-gain uniform weights (var 1/fan_in), norm gains 1 +- 0.1,
                   non-zero norm/conv biases, so every gain/bias path is exercised and activations
                   stay O(1..3) through the net.  Used by the golden fixtures.
    init="torch" : what the reference constructors produce: nn.Conv2d / nn.Linear
                   defaults (kaiming_uniform(a=sqrt 5): U(+-1/sqrt(fan_in)) for weight and bias),
                   GroupNorm / RMSNorm gains 1, biases 0, nn.Embedding N(0,1), sinusoidal weights
                   randn (model.py:231), PixelShuffle conv = kaiming-uniform rows repeated x4 with
                   zero bias (model.py:88-95).  Same distribution, not the same draws as torch's
                   own constructors (those cannot run without the reference's pip dependencies)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = OrderedDict()
    for name, shape in param_shapes(spec, prefix).items():
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        if name.endswith("time_mlp.0.weights") or name.endswith("class_mlp.0.weight"):
            t = torch.randn(shape, generator=g)
        elif name.endswith(".g") or name.endswith("norm.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g) if init == "unit" else torch.ones(shape)
        elif name.endswith("norm.bias"):
            t = 0.1 * torch.randn(shape, generator=g) if init == "unit" else torch.zeros(shape)
        elif name.endswith(".bias"):
            if init == "unit":
                t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
            elif ".net.0." in name:
                t = torch.zeros(shape)
            else:
                wshape = param_shapes(spec, prefix)[name[:-4] + "weight"]
                fi = 1
                for s in wshape[1:]:
                    fi *= s
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fi)
        elif init == "torch" and ".net.0.weight" in name:
            o, i, kh, kw = shape
            base = (torch.rand((o // 4, i, kh, kw), generator=g) * 2 - 1) * math.sqrt(6.0 / (i * kh * kw))
            t = base.repeat_interleave(4, dim=0)                           # 'o ... -> (o 4) ...'
        else:
            bound = math.sqrt(3.0 / fan_in) if init == "unit" else 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name] = t.float().contiguous()
    return sd


# --------------------------------------------------------------------------------------------
# U-Net building blocks
# --------------------------------------------------------------------------------------------

def _rmsnorm(x, g):
    """model.py:201-207  F.normalize(x, dim=1) * g * sqrt(C)  (eps 1e-12 on the L2 norm)."""
    return F.normalize(x, dim=1) * g * (x.shape[1] ** 0.5)


def _block(sd, p, x, groups, scale_shift=None):
    """model.py:243-259  conv3x3 -> GroupNorm -> (scale+1)*x+shift -> SiLU."""
    x = F.conv2d(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"], padding=1)
    x = F.group_norm(x, groups, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def _resblock(sd, p, x, t, groups):
    """model.py:261-285."""
    ss = None
    if t is not None:
        e = F.linear(F.silu(t), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])
        e = e[:, :, None, None]
        ss = e.chunk(2, dim=1)
    h = _block(sd, p + ".block1", x, groups, ss)
    h = _block(sd, p + ".block2", h, groups)
    if (p + ".res_conv.weight") in sd:
        x = F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    return h + x


def _linear_attention(sd, p, x, heads, dim_head):
    """model.py:287-324."""
    b, c, h, w = x.shape
    x = _rmsnorm(x, sd[p + ".norm.g"])
    qkv = F.conv2d(x, sd[p + ".to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, dim_head, h * w) for t in qkv)         # b h c (x y)
    q = q.softmax(dim=-2)
    k = k.softmax(dim=-1)
    q = q * dim_head ** -0.5
    context = torch.einsum("bhdn,bhen->bhde", k, v)
    out = torch.einsum("bhde,bhdn->bhen", context, q)
    out = out.reshape(b, heads * dim_head, h, w)
    out = F.conv2d(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return _rmsnorm(out, sd[p + ".to_out.1.g"])


def _attend(q, k, v):
    """denoising-diffusion-pytorch==1.8.15 Attend.forward, flash=False (call site model.py:352).
    PARITY UNPINNED at this third-party boundary (see module docstring)."""
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * q.shape[-1] ** -0.5
    return torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)


def _attention(sd, p, x, heads, dim_head):
    """model.py:326-355."""
    b, c, h, w = x.shape
    x = _rmsnorm(x, sd[p + ".norm.g"])
    qkv = F.conv2d(x, sd[p + ".to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, dim_head, h * w).transpose(-1, -2) for t in qkv)  # b h (x y) c
    out = _attend(q, k, v)
    out = out.transpose(-1, -2).reshape(b, heads * dim_head, h, w)
    return F.conv2d(out, sd[p + ".to_out.weight"], sd[p + ".to_out.bias"])


def _downsample(sd, p, x):
    """model.py:106-110  'b c (h p1) (w p2) -> b (c p1 p2) h w' then 1x1 conv."""
    return F.conv2d(F.pixel_unshuffle(x, 2), sd[p + ".1.weight"], sd[p + ".1.bias"])


def _pixel_shuffle_upsample(sd, p, x):
    """model.py:70-98  1x1 conv (C -> 4C') -> SiLU -> PixelShuffle(2)."""
    return F.pixel_shuffle(F.silu(F.conv2d(x, sd[p + ".net.0.weight"], sd[p + ".net.0.bias"])), 2)


def time_embedding(sd, spec: UnetSpec, time, class_label=None, prefix="model."):
    """model.py:223-238 + 603-608 (time_mlp) and 612-619 + 692-694 (class_mlp)."""
    if spec.learned_sinusoidal_cond:
        x = time.reshape(-1, 1).float()
        freqs = x * sd[prefix + "time_mlp.0.weights"][None, :] * 2 * math.pi
        f = torch.cat((x, freqs.sin(), freqs.cos()), dim=-1)
    else:                                                                 # SinusoidalPosEmb, model.py:209-221
        half_dim = spec.dim // 2
        emb = math.log(10000) / (half_dim - 1)
        emb = torch.exp(torch.arange(half_dim, device=time.device) * -emb)
        emb = time.reshape(-1)[:, None] * emb[None, :]
        f = torch.cat((emb.sin(), emb.cos()), dim=-1)
    t = F.linear(f, sd[prefix + "time_mlp.1.weight"], sd[prefix + "time_mlp.1.bias"])
    t = F.linear(F.gelu(t), sd[prefix + "time_mlp.3.weight"], sd[prefix + "time_mlp.3.bias"])
    if class_label is not None:
        c = sd[prefix + "class_mlp.0.weight"][class_label]
        c = F.linear(c, sd[prefix + "class_mlp.1.weight"], sd[prefix + "class_mlp.1.bias"])
        c = F.linear(F.gelu(c), sd[prefix + "class_mlp.3.weight"], sd[prefix + "class_mlp.3.bias"])
        t = t + c
    return t


def unet_forward(sd, spec: UnetSpec, x, time, class_label=None, x_self_cond=None, prefix="model.",
                 taps: Optional[dict] = None):
    """ConditionalSRUnet.forward, model.py:678-725.  `taps` (optional dict) receives named
    intermediate activations for layer-level parity tests."""
    factor = 2 ** (len(spec.dim_mults) - 1)
    assert all(d % factor == 0 for d in x.shape[-2:]), \
        f"your input dimensions {tuple(x.shape[-2:])} need to be divisible by {factor}, given the unet"
    g, nh, dh = spec.groups, spec.heads, spec.dim_head
    P = prefix

    def tap(name, v):
        if taps is not None:
            taps[name] = v
        return v

    if x_self_cond is None:
        x_self_cond = torch.zeros_like(x)                                 # model.py:682
    x = torch.cat((x, x_self_cond), dim=1)                                # model.py:684
    x = F.conv2d(x, sd[P + "init_conv.weight"], sd[P + "init_conv.bias"], padding=3)
    tap("init_conv", x)
    r = x
    t = tap("t_emb", time_embedding(sd, spec, time, class_label, prefix))
    hs = []
    n = len(spec.in_out)
    for i, full in enumerate(spec.full_attn):                             # model.py:698-706
        x = tap(f"downs.{i}.0", _resblock(sd, f"{P}downs.{i}.0", x, t, g))
        hs.append(x)
        x = tap(f"downs.{i}.1", _resblock(sd, f"{P}downs.{i}.1", x, t, g))
        att = _attention if full else _linear_attention
        x = tap(f"downs.{i}.2", att(sd, f"{P}downs.{i}.2", x, nh, dh) + x)
        hs.append(x)
        if i < n - 1:
            x = _downsample(sd, f"{P}downs.{i}.3", x)
        else:
            x = F.conv2d(x, sd[f"{P}downs.{i}.3.weight"], sd[f"{P}downs.{i}.3.bias"], padding=1)
        tap(f"downs.{i}.3", x)
    x = tap("mid_block1", _resblock(sd, P + "mid_block1", x, t, g))       # model.py:708-710
    x = tap("mid_attn", _attention(sd, P + "mid_attn", x, nh, dh) + x)
    x = tap("mid_block2", _resblock(sd, P + "mid_block2", x, t, g))
    for i, full in enumerate(reversed(spec.full_attn)):                   # model.py:712-720
        x = torch.cat((x, hs.pop()), dim=1)
        x = tap(f"ups.{i}.0", _resblock(sd, f"{P}ups.{i}.0", x, t, g))
        x = torch.cat((x, hs.pop()), dim=1)
        x = tap(f"ups.{i}.1", _resblock(sd, f"{P}ups.{i}.1", x, t, g))
        att = _attention if full else _linear_attention
        x = tap(f"ups.{i}.2", att(sd, f"{P}ups.{i}.2", x, nh, dh) + x)
        if i < n - 1:
            x = _pixel_shuffle_upsample(sd, f"{P}ups.{i}.3", x)
        else:
            x = F.conv2d(x, sd[f"{P}ups.{i}.3.weight"], sd[f"{P}ups.{i}.3.bias"], padding=1)
        tap(f"ups.{i}.3", x)
    x = torch.cat((x, r), dim=1)                                          # model.py:722
    x = tap("final_res_block", _resblock(sd, P + "final_res_block", x, t, g))
    return F.conv2d(x, sd[P + "final_conv.weight"], sd[P + "final_conv.bias"])


# --------------------------------------------------------------------------------------------
# continuous-time schedule and sampler
# --------------------------------------------------------------------------------------------

def log_snr_linear(t: torch.Tensor) -> torch.Tensor:
    """beta_linear_log_snr, model.py:2629-2633:  -log(clamp(expm1(1e-4 + 10 t^2), 1e-20))."""
    return -torch.log(torch.special.expm1(1e-4 + 10 * (t ** 2)).clamp(min=1e-20))


def step_scalars(time: torch.Tensor, time_next: torch.Tensor):
    """The 0-dim fp32 tensor arithmetic of p_mean_variance, model.py:3127-3134, 3168."""
    log_snr, log_snr_next = log_snr_linear(time), log_snr_linear(time_next)
    c = -torch.special.expm1(log_snr - log_snr_next)
    sq_alpha, sq_alpha_next = log_snr.sigmoid(), log_snr_next.sigmoid()
    sq_sigma, sq_sigma_next = (-log_snr).sigmoid(), (-log_snr_next).sigmoid()
    alpha, sigma, alpha_next = sq_alpha.sqrt(), sq_sigma.sqrt(), sq_alpha_next.sqrt()
    return dict(log_snr=log_snr, log_snr_next=log_snr_next, c=c, alpha=alpha, sigma=sigma,
                alpha_next=alpha_next, var=sq_sigma_next * c)


def guided_noise(sd, spec, x, batch_log_snr, condition_x, class_label, cond_scale, class_cond_scale):
    """CFG dispatch + combine, model.py:3138-3158.  Returns (pred_noise, cond_out, null_out)."""
    if (cond_scale != 1.0) and (class_cond_scale != 1.0):
        raise NotImplementedError(
            "Currently, you cannot specify both cond_scale and class_cond_scale at the same time.")
    if cond_scale != 1.0:
        cond = unet_forward(sd, spec, x, batch_log_snr, class_label, condition_x)
        null = unet_forward(sd, spec, x, batch_log_snr, class_label, None)
        return null + (cond - null) * cond_scale, cond, null
    if class_cond_scale != 1.0:
        cond = unet_forward(sd, spec, x, batch_log_snr, class_label, condition_x)
        null = unet_forward(sd, spec, x, batch_log_snr, None, condition_x)
        return null + (cond - null) * class_cond_scale, cond, null
    out = unet_forward(sd, spec, x, batch_log_snr, class_label, condition_x)
    return out, out, None


def posterior_update(x, pred_noise, s, clip=True):
    """x0 / clamp / posterior mean, model.py:3160-3168."""
    x_start = (x - s["sigma"] * pred_noise) / s["alpha"]
    if clip:
        x_start = x_start.clamp(-1., 1.)
        mean = s["alpha_next"] * (x * (1 - s["c"]) / s["alpha"] + s["c"] * x_start)
    else:
        mean = s["alpha_next"] / s["alpha"] * (x - s["c"] * s["sigma"] * pred_noise)
    return mean, s["var"], x_start


def p_mean_variance(sd, spec, x, time, condition_x, class_label, cond_scale, class_cond_scale,
                    time_next, clip=True):
    """model.py:3122-3170."""
    s = step_scalars(time, time_next)
    batch_log_snr = s["log_snr"].reshape(1).expand(x.shape[0])            # model.py:3136
    eps, _, _ = guided_noise(sd, spec, x, batch_log_snr, condition_x, class_label,
                             cond_scale, class_cond_scale)
    return posterior_update(x, eps, s, clip)


def p_sample(sd, spec, x, time, condition_x, class_label, cond_scale, class_cond_scale, time_next,
             noise=None, generator=None, clip=True):
    """model.py:3174-3188.  `noise` teacher-forces the randn_like draw."""
    mean, var, x_start = p_mean_variance(sd, spec, x, time, condition_x, class_label,
                                         cond_scale, class_cond_scale, time_next, clip)
    if time_next == 0:
        return mean, x_start
    if noise is None:
        noise = torch.randn(x.shape, generator=generator, device=x.device)      # randn_like(x)
    return mean + var.sqrt() * noise, x_start


def q_sample(x_start, times, noise=None, generator=None):
    """model.py:3434-3447."""
    if noise is None:
        noise = torch.randn(x_start.shape, generator=generator, device=x_start.device)
    log_snr = log_snr_linear(times)
    pad = log_snr.reshape(*log_snr.shape, *((1,) * max(0, x_start.ndim - log_snr.ndim)))
    alpha, sigma = pad.sigmoid().sqrt(), (-pad).sigmoid().sqrt()
    return x_start * alpha + noise * sigma, log_snr


def sample(sd, spec, batch_size, condition_x, class_label=None, cond_scale=1.0,
           guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
           generation_start_steps=0, num_sample_steps=250, image_size=256, generator=None,
           clip=True, progress=None):
    """sample -> p_sample_loop, model.py:3417-3430 -> 3191-3246 (condition_x in [0,1])."""
    condition_x = condition_x * 2 - 1
    shape = (batch_size, spec.channels, image_size, image_size)
    if generation_start_steps > 0:
        st = 1. - torch.tensor(generation_start_steps / num_sample_steps, device=condition_x.device)
        img, _ = q_sample(condition_x, st.reshape(1).expand(batch_size), generator=generator)
    else:
        img = torch.randn(shape, generator=generator, device=condition_x.device)
    steps = torch.linspace(1., 0., num_sample_steps + 1, device=condition_x.device)
    for i in range(num_sample_steps):
        if i < generation_start_steps:
            continue
        cs = 1.0 if i < guidance_start_steps else cond_scale
        ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
        img, _x0 = p_sample(sd, spec, img, steps[i], condition_x, class_label, cs, ccs,
                            steps[i + 1], generator=generator, clip=clip)
        if progress is not None:
            progress(i, img, _x0)
    return (img.clamp(-1., 1.) + 1) * 0.5


# --------------------------------------------------------------------------------------------
# tile geometry and tiled sampling (the direct caller of the hot path)
# --------------------------------------------------------------------------------------------

def get_coord_and_pad(height, width, tile_size=256):
    """model.py:116-135."""
    if height <= tile_size and width <= tile_size:
        nh, nw = tile_size, tile_size
    else:
        nh = ((height - 1) // tile_size + 1) * tile_size + tile_size
        nw = ((width - 1) // tile_size + 1) * tile_size + tile_size
    left, top = (nw - width) // 2, (nh - height) // 2
    coord = (left, top, left + width, top + height)
    pad = (left, nw - left - width, top, nh - top - height)
    return coord, pad


def get_coords(h, w, tile_size, tile_stride, diff=0):
    """model.py:137-150."""
    hi = list(range(0, h - tile_size + 1, tile_stride))
    if (h - tile_size) % tile_stride != 0:
        hi.append(h - tile_size)
    wi = list(range(0, w - tile_size + 1, tile_stride))
    if (w - tile_size) % tile_stride != 0:
        wi.append(w - tile_size)
    return [(a + diff, a + tile_size + diff, b + diff, b + tile_size + diff) for a in hi for b in wi]


def get_area(coords, height, width):
    """model.py:152-179."""
    top, bottom, left, right = height, 0, width, 0
    for hs, he, ws, we in coords:
        top, bottom = min(top, hs), max(bottom, he)
        left, right = min(left, ws), max(right, we)
    coord = (left, top, right, bottom)
    pad = (left, width - left - (right - left), top, height - top - (bottom - top))
    return coord, pad


def tiled_setup(condition_x, tile_size=256, tile_stride=256):
    """Canvas set-up of tiled_sample, model.py:3296-3342: reflect-padded condition in [-1,1], the two tile
    grids, the hull of the shifted grid, the hull-masked condition canvas and the crop window."""
    condition_x = condition_x * 2 - 1
    batch, c, h, w = condition_x.shape
    (left, top, right, bottom), pad = get_coord_and_pad(h, w)       # model.py:3301 (default 256!)
    condition_x = F.pad(condition_x, pad, mode="reflect")
    _, _, height, width = condition_x.shape
    coords0 = get_coords(height, width, tile_size, tile_size, 0)
    if height <= tile_size and width <= tile_size:
        coords1 = get_coords(height, width, tile_size, tile_stride, 0)
    else:
        coords1 = get_coords(height - tile_size, width - tile_size, tile_size, tile_stride, tile_size // 2)
    (sleft, stop, sright, sbottom), small_pad = get_area(coords1, height, width)
    masked = F.pad(condition_x[:, :, stop:sbottom, sleft:sright], small_pad, mode="constant", value=0)
    return dict(padded=condition_x, masked=masked, coord_list=[coords0, coords1],
                hull=(stop, sbottom, sleft, sright), crop=(top, bottom, left, right))


def tiled_steps(sd, spec, img, x_start, masked_condition, coord_list, hull, steps, first, last, batch_size,
                class_label=None, cond_scale=1.0, guidance_start_steps=0, class_cond_scale=1.0,
                class_guidance_start_steps=0, generator=None, clip=True):
    """Steps [first, last) of the sampling loop of tiled_sample, model.py:3345-3401, on the canvas `img`
    (updated in place where the reference does; the re-noised canvas is a new tensor, returned)."""
    stop, sbottom, sleft, sright = hull
    for i in range(first, last):
        cs = 1.0 if i < guidance_start_steps else cond_scale
        ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
        cur = coord_list[i % 2]
        for s0 in range(0, len(cur), batch_size):
            chunk = cur[s0:s0 + batch_size]
            mb = torch.cat([img[:, :, hs:he, ws:we] for hs, he, ws, we in chunk], 0)
            mc = torch.cat([masked_condition[:, :, hs:he, ws:we] for hs, he, ws, we in chunk], 0)
            out, x0 = p_sample(sd, spec, mb, steps[i], mc, class_label, cs, ccs, steps[i + 1],
                               generator=generator, clip=clip)
            for k, (hs, he, ws, we) in enumerate(chunk):
                img[:, :, hs:he, ws:we] = out[k]
                x_start[:, :, hs:he, ws:we] = x0[k]
        if i % 2 == 1:
            cropped = img[:, :, stop:sbottom, sleft:sright]
            img, _ = q_sample(torch.zeros_like(masked_condition), steps[i + 1], generator=generator)
            img[:, :, stop:sbottom, sleft:sright] = cropped
    return img, x_start


def tiled_sample(sd, spec, batch_size, condition_x, class_label=None, cond_scale=1.0,
                 guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
                 generation_start_steps=0, num_sample_steps=250, tile_size=256, tile_stride=256,
                 generator=None, clip=True):
    """model.py:3288-3413 (start_white_noise=True path; RNG call order preserved)."""
    batch = condition_x.shape[0]
    ts = tiled_setup(condition_x, tile_size, tile_stride)
    padded = ts["padded"]
    if generation_start_steps > 0:
        st = 1. - torch.tensor(generation_start_steps / num_sample_steps, device=padded.device)
        img, _ = q_sample(padded, st.reshape(1).expand(batch), generator=generator)
    else:
        img = torch.randn(padded.shape, generator=generator, device=padded.device)
    steps = torch.linspace(1., 0., num_sample_steps + 1, device=padded.device)
    x_start = img.clone()
    img, _ = tiled_steps(sd, spec, img, x_start, ts["masked"], ts["coord_list"], ts["hull"], steps,
                         max(0, generation_start_steps), num_sample_steps, batch_size, class_label, cond_scale,
                         guidance_start_steps, class_cond_scale, class_guidance_start_steps, generator, clip)
    top, bottom, left, right = ts["crop"]
    img = img[:, :, top:bottom, left:right].clamp(-1., 1.)
    return (img + 1) * 0.5


# --------------------------------------------------------------------------------------------
# EDM sampler family on the same U-Net: ConditionalElucidatedDiffusionSR,
# model.py:2059-2560.  The preconditioning coefficients and the sigma schedule come from the pip
# package's `ElucidatedDiffusion` base class, which is not under /root/reference: PARITY UNPINNED
# for `edm_schedule` and `edm_coeffs` (restated from denoising-diffusion-pytorch 1.8.15 / Karras et
# al. 2022, Table 1; the goldens were produced with the same restatement in oracle/_shim).
# Everything else below follows the reference lines cited.  State-dict prefix of this family: "net.".
# --------------------------------------------------------------------------------------------

@dataclass(frozen=True)
class EdmParams:
    sigma_min: float = 0.002
    sigma_max: float = 80
    sigma_data: float = 0.5
    rho: float = 7
    S_churn: float = 80
    S_tmin: float = 0.05
    S_tmax: float = 50
    S_noise: float = 1.003


def edm_schedule(p: EdmParams, n: int, device=None) -> torch.Tensor:
    """ElucidatedDiffusion.sample_schedule (pip, unpinned): n sigmas on the rho-grid + a trailing 0."""
    inv_rho = 1 / p.rho
    steps = torch.arange(n, device=device, dtype=torch.float32)
    sigmas = (p.sigma_max ** inv_rho + steps / (n - 1) * (p.sigma_min ** inv_rho - p.sigma_max ** inv_rho)) ** p.rho
    return F.pad(sigmas, (0, 1), value=0.)


def edm_coeffs(p: EdmParams, sigma: torch.Tensor):
    """c_in, c_out, c_skip, c_noise of the pip base class (unpinned), fp32 tensor arithmetic."""
    sd2 = p.sigma_data ** 2
    c_in = 1 * (sigma ** 2 + sd2) ** -0.5
    c_out = sigma * p.sigma_data * (sd2 + sigma ** 2) ** -0.5
    c_skip = sd2 / (sigma ** 2 + sd2)
    c_noise = torch.log(sigma.clamp(min=1e-20)) * 0.25
    return c_in, c_out, c_skip, c_noise


def edm_denoise(sd, spec, p: EdmParams, x, sigma, condition_x, class_label, cond_scale=1.0, class_cond_scale=1.0,
                clamp=False, prefix="net."):
    """preconditioned_network_forward, model.py:2128-2183 (sigma: python float or [B] tensor)."""
    b = x.shape[0]
    if isinstance(sigma, float):
        sigma = torch.full((b,), sigma, device=x.device)
    pad = sigma.reshape(b, 1, 1, 1)
    c_in, c_out, c_skip, _ = edm_coeffs(p, pad)
    c_noise = edm_coeffs(p, sigma)[3]
    out = c_skip * x + c_out * unet_forward(sd, spec, c_in * x, c_noise, class_label, condition_x, prefix=prefix)
    if (cond_scale != 1.0) and (class_cond_scale != 1.0):
        raise NotImplementedError(
            "Currently, you cannot specify both cond_scale and class_cond_scale at the same time.")
    if cond_scale != 1.0:
        null = c_skip * x + c_out * unet_forward(sd, spec, c_in * x, c_noise, class_label, None, prefix=prefix)
        out = null + (out - null) * cond_scale
    if class_cond_scale != 1.0:
        null = c_skip * x + c_out * unet_forward(sd, spec, c_in * x, c_noise, None, condition_x, prefix=prefix)
        out = null + (out - null) * class_cond_scale
    return out.clamp(-1., 1.) if clamp else out


def _edm_gammas(p: EdmParams, sigmas, n):
    """model.py:2234-2238."""
    return torch.where((sigmas >= p.S_tmin) & (sigmas <= p.S_tmax), min(p.S_churn / n, math.sqrt(2) - 1), 0.)


def _edm_heun(sd, spec, p, x_hat, sigma_hat, sigma_next, cond, label, cs, ccs, clamp):
    """One Heun step on (a batch of tiles of) images_hat, model.py:2276-2289 / 2400-2410.
    Returns (images_next, the x0-side quantity the reference records for with_x0_images / x_start)."""
    out = edm_denoise(sd, spec, p, x_hat, sigma_hat, cond, label, cs, ccs, clamp)
    d = (x_hat - out) / sigma_hat
    nxt = x_hat + (sigma_next - sigma_hat) * d
    if sigma_next != 0:
        out2 = edm_denoise(sd, spec, p, nxt, sigma_next, cond, label, cs, ccs, clamp)
        d2 = (nxt - out2) / sigma_next
        nxt = x_hat + 0.5 * (sigma_next - sigma_hat) * (d + d2)
        return nxt, d2
    return nxt, d


def edm_sample_heun(sd, spec, p: EdmParams, batch_size, condition_x, class_label=None, cond_scale=1.0,
                    guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
                    generation_start_steps=0, num_sample_steps=32, clamp=True, zero_init=False, generator=None,
                    model_steps=None):
    """sample_org, model.py:2213-2307 (condition_x in [0,1]; RNG call order preserved).  `model_steps` = the object's
    own num_sample_steps, whose schedule get_noised_images uses (model.py:2189, 2241); default: num_sample_steps."""
    _n, _c, h, w = condition_x.shape
    shape = (batch_size, spec.channels, h, w)
    dev = condition_x.device
    condition_x = condition_x * 2 - 1
    sigmas = edm_schedule(p, num_sample_steps, dev)
    gammas = _edm_gammas(p, sigmas, num_sample_steps)
    if generation_start_steps > 0:                                        # get_noised_images, model.py:2186-2195
        images = condition_x + edm_schedule(p, model_steps or num_sample_steps, dev)[generation_start_steps] * \
            torch.randn(condition_x.shape, generator=generator, device=dev)
    elif zero_init:
        images = torch.zeros(shape, device=dev)
    else:
        images = sigmas[0] * torch.randn(shape, generator=generator, device=dev)
    for i in range(num_sample_steps):
        if i < generation_start_steps:
            continue
        cs = 1.0 if i < guidance_start_steps else cond_scale
        ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
        sigma, sigma_next, gamma = sigmas[i].item(), sigmas[i + 1].item(), gammas[i].item()
        eps = p.S_noise * torch.randn(shape, generator=generator, device=dev)
        sigma_hat = sigma + gamma * sigma
        images_hat = images + math.sqrt(sigma_hat ** 2 - sigma ** 2) * eps
        images, _ = _edm_heun(sd, spec, p, images_hat, sigma_hat, sigma_next, condition_x, class_label, cs, ccs, clamp)
    return (images.clamp(-1., 1.) + 1) * 0.5


def edm_sample_dpmpp(sd, spec, p: EdmParams, batch_size, condition_x, class_label=None, cond_scale=1.0,
                     guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
                     generation_start_steps=0, num_sample_steps=32, clamp=True, zero_init=False, generator=None):
    """sample_using_dpmpp (DPM-Solver++ 2M), model.py:2466-2544."""
    _n, _c, h, w = condition_x.shape
    shape = (batch_size, spec.channels, h, w)
    dev = condition_x.device
    condition_x = condition_x * 2 - 1
    sigmas = edm_schedule(p, num_sample_steps, dev)
    if generation_start_steps > 0:
        images = condition_x + sigmas[generation_start_steps] * torch.randn(condition_x.shape, generator=generator, device=dev)
    elif zero_init:
        images = torch.zeros(shape, device=dev)
    else:
        images = sigmas[0] * torch.randn(shape, generator=generator, device=dev)
    t_fn = lambda s: s.log().neg()
    sigma_fn = lambda t: t.neg().exp()
    old = None
    for i in range(len(sigmas) - 1):
        if i < generation_start_steps:
            continue
        cs = 1.0 if i < guidance_start_steps else cond_scale
        ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
        den = edm_denoise(sd, spec, p, images, sigmas[i].item(), condition_x, class_label, cs, ccs, clamp)
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        hh = t_next - t
        if old is None or sigmas[i + 1] == 0:
            den_d = den
        else:
            r = (t - t_fn(sigmas[i - 1])) / hh
            g = -1 / (2 * r)
            den_d = (1 - g) * den + g * old
        images = (sigma_fn(t_next) / sigma_fn(t)) * images - (-hh).expm1() * den_d
        old = den
    return (images.clamp(-1., 1.) + 1) * 0.5


def edm_tiled_sample(sd, spec, p: EdmParams, batch_size, condition_x, class_label=None, cond_scale=1.0,
                     guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0,
                     generation_start_steps=0, num_sample_steps=32, tile_size=256, tile_stride=256, clamp=True,
                     zero_init=False, generator=None, model_steps=None):
    """tiled_sample of the EDM class (Heun), model.py:2309-2462: the whole canvas is perturbed first, the tiles of the
    step's grid are denoised from images_hat in minibatches, and after odd steps everything outside the hull of the
    shifted grid is replaced by sigma_i * noise (get_noised_images(zeros, i): the CURRENT step's sigma)."""
    dev = condition_x.device
    ts = tiled_setup(condition_x, tile_size, tile_stride)
    padded, masked = ts["padded"], ts["masked"]
    shape = padded.shape
    sigmas = edm_schedule(p, num_sample_steps, dev)
    gammas = _edm_gammas(p, sigmas, num_sample_steps)
    if generation_start_steps > 0:
        images = padded + sigmas[generation_start_steps] * torch.randn(shape, generator=generator, device=dev)
    elif zero_init:
        images = torch.zeros(shape, device=dev)
    else:
        images = sigmas[0] * torch.randn(shape, generator=generator, device=dev)
    stop, sbottom, sleft, sright = ts["hull"]
    for i in range(num_sample_steps):
        if i < generation_start_steps:
            continue
        cs = 1.0 if i < guidance_start_steps else cond_scale
        ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
        sigma, sigma_next, gamma = sigmas[i].item(), sigmas[i + 1].item(), gammas[i].item()
        eps = p.S_noise * torch.randn(shape, generator=generator, device=dev)
        sigma_hat = sigma + gamma * sigma
        images_hat = images + math.sqrt(sigma_hat ** 2 - sigma ** 2) * eps
        cur = ts["coord_list"][i % 2]
        for s0 in range(0, len(cur), batch_size):
            chunk = cur[s0:s0 + batch_size]
            mb = torch.cat([images_hat[:, :, hs:he, ws:we] for hs, he, ws, we in chunk], 0)
            mc = torch.cat([masked[:, :, hs:he, ws:we] for hs, he, ws, we in chunk], 0)
            nxt, _ = _edm_heun(sd, spec, p, mb, sigma_hat, sigma_next, mc, class_label, cs, ccs, clamp)
            for k, (hs, he, ws, we) in enumerate(chunk):
                images[:, :, hs:he, ws:we] = nxt[k]
        if i % 2 == 1:
            cropped = images[:, :, stop:sbottom, sleft:sright]
            images = edm_schedule(p, model_steps or num_sample_steps, dev)[i] * \
                torch.randn(shape, generator=generator, device=dev)                       # zeros + sigma_i * noise
            images[:, :, stop:sbottom, sleft:sright] = cropped
    top, bottom, left, right = ts["crop"]
    return (images[:, :, top:bottom, left:right].clamp(-1., 1.) + 1) * 0.5


# --------------------------------------------------------------------------------------------
# Discrete-time sampler family on the same U-Net: ConditionalGaussianDiffusionSR,
# model.py:1311-1660 -- DDPM ancestral sampling and DDIM.  The schedules and buffers are the reference's own code
# (model.py:744-777, 1362-1424: pinned by the goldens); `predict_start_from_noise`, `predict_noise_from_start`,
# `predict_start_from_v`, `q_posterior` and `q_sample` come from the absent pip base class `GaussianDiffusion`
# (denoising-diffusion-pytorch==1.8.15) and are restated from the published algebra: PARITY UNPINNED at that
# boundary.  tests/golden/gauss_tiny.npz (tests/golden/make_golden_gauss.py) pins the rest against the unmodified
# reference class running on the same restated base (oracle/_shim).
# --------------------------------------------------------------------------------------------

def gauss_betas(schedule: str, timesteps: int) -> torch.Tensor:
    """model.py:744-777 (float64)."""
    if schedule == "linear":
        scale = 1000 / timesteps
        return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    if schedule == "cosine":
        ac = torch.cos((t + 0.008) / (1 + 0.008) * math.pi * 0.5) ** 2
    elif schedule == "sigmoid":
        start, end, tau = -3, 3, 1
        v_start, v_end = torch.tensor(start / tau).sigmoid(), torch.tensor(end / tau).sigmoid()
        ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    else:
        raise ValueError(f"unknown beta schedule {schedule}")
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


@dataclass(frozen=True)
class GaussParams:
    timesteps: int = 1000
    sampling_timesteps: int = 250
    objective: str = "pred_noise"
    beta_schedule: str = "linear"
    ddim_sampling_eta: float = 0.


def gauss_tables(p: GaussParams) -> Dict[str, torch.Tensor]:
    """The registered fp32 buffers, model.py:1371-1424."""
    betas = gauss_betas(p.beta_schedule, p.timesteps)
    alphas = 1. - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.)
    pv = betas * (1. - ac_prev) / (1. - ac)
    T = dict(betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev, sqrt_alphas_cumprod=torch.sqrt(ac),
             sqrt_one_minus_alphas_cumprod=torch.sqrt(1. - ac), sqrt_recip_alphas_cumprod=torch.sqrt(1. / ac),
             sqrt_recipm1_alphas_cumprod=torch.sqrt(1. / ac - 1), posterior_variance=pv,
             posterior_log_variance_clipped=torch.log(pv.clamp(min=1e-20)),
             posterior_mean_coef1=betas * torch.sqrt(ac_prev) / (1. - ac),
             posterior_mean_coef2=(1. - ac_prev) * torch.sqrt(alphas) / (1. - ac))
    return {k: v.to(torch.float32) for k, v in T.items()}


def _gx(a, t, x):
    """extract, model.py:730-733"""
    return a.to(x.device).gather(-1, t).reshape(t.shape[0], *((1,) * (x.ndim - 1)))


def gauss_model_predictions(sd, spec, p: GaussParams, T, x, t, condition_x=None, class_label=None, cond_scale=1.0,
                            class_cond_scale=1.0, clip_x_start=False, rederive_pred_noise=False):
    """model.py:1449-1489 -> (pred_noise, x_start)."""
    if (cond_scale != 1.0) and (class_cond_scale != 1.0):
        raise NotImplementedError("Currently, you cannot specify both cond_scale and class_cond_scale at the same time.")
    net = lambda lab, cond: unet_forward(sd, spec, x, t, lab, cond)
    if cond_scale == 1.0 and class_cond_scale == 1.0:
        out = net(class_label, condition_x)
    elif cond_scale != 1.0:
        cond_out, null_out = net(class_label, condition_x), net(class_label, None)
        out = null_out + (cond_out - null_out) * cond_scale
    else:
        cond_out, null_out = net(class_label, condition_x), net(None, condition_x)
        out = null_out + (cond_out - null_out) * class_cond_scale
    clip = (lambda v: torch.clamp(v, min=-1., max=1.)) if clip_x_start else (lambda v: v)
    from_start = lambda x0: (_gx(T["sqrt_recip_alphas_cumprod"], t, x) * x - x0) / _gx(T["sqrt_recipm1_alphas_cumprod"], t, x)
    if p.objective == "pred_noise":
        pred_noise = out
        x_start = clip(_gx(T["sqrt_recip_alphas_cumprod"], t, x) * x - _gx(T["sqrt_recipm1_alphas_cumprod"], t, x) * out)
        if clip_x_start and rederive_pred_noise:
            pred_noise = from_start(x_start)
    elif p.objective == "pred_x0":
        x_start = clip(out)
        pred_noise = from_start(x_start)
    else:
        x_start = clip(_gx(T["sqrt_alphas_cumprod"], t, x) * x - _gx(T["sqrt_one_minus_alphas_cumprod"], t, x) * out)
        pred_noise = from_start(x_start)
    return pred_noise, x_start


def gauss_p_sample(sd, spec, p: GaussParams, T, x, t: int, condition_x, class_label, cond_scale=1.0,
                   class_cond_scale=1.0, noise=None, generator=None):
    """model.py:1491-1514 -> (pred_img, x_start).  `noise` overrides the randn_like draw (teacher-forced tests)."""
    bt = torch.full((x.shape[0],), t, device=x.device, dtype=torch.long)
    _, x_start = gauss_model_predictions(sd, spec, p, T, x, bt, condition_x, class_label, cond_scale, class_cond_scale)
    x_start = x_start.clamp(-1., 1.)
    mean = _gx(T["posterior_mean_coef1"], bt, x) * x_start + _gx(T["posterior_mean_coef2"], bt, x) * x
    log_var = _gx(T["posterior_log_variance_clipped"], bt, x)
    if noise is None:
        noise = torch.randn(x.shape, generator=generator, device=x.device) if t > 0 else 0.
    elif t == 0:
        noise = 0.
    return mean + (0.5 * log_var).exp() * noise, x_start


def gauss_ddim_step(sd, spec, p: GaussParams, T, img, time: int, time_next: int, condition_x, class_label,
                    cond_scale=1.0, class_cond_scale=1.0, noise=None, generator=None):
    """One iteration of ddim_sample's loop, model.py:1599-1622 -> (img, x_start)."""
    tc = torch.full((img.shape[0],), time, device=img.device, dtype=torch.long)
    pred_noise, x_start = gauss_model_predictions(sd, spec, p, T, img, tc, condition_x, class_label, cond_scale,
                                                  class_cond_scale, clip_x_start=True, rederive_pred_noise=True)
    if time_next < 0:
        return x_start, x_start
    alpha, alpha_next = T["alphas_cumprod"][time], T["alphas_cumprod"][time_next]
    sigma = p.ddim_sampling_eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
    c = (1 - alpha_next - sigma ** 2).sqrt()
    if noise is None:
        noise = torch.randn(img.shape, generator=generator, device=img.device)
    return x_start * alpha_next.sqrt() + c * pred_noise + sigma * noise, x_start


def gauss_sample(sd, spec, p: GaussParams, batch_size, condition_x, class_label=None, cond_scale=1.0,
                 guidance_start_steps=0, class_cond_scale=1.0, class_guidance_start_steps=0, generation_start_steps=0,
                 num_sample_steps=None, model_steps=None, generator=None):
    """sample -> p_sample_loop (sampling_timesteps == timesteps) or ddim_sample, model.py:1517-1660.  `model_steps`
    (optional list) receives the state after every step; draws come from `generator` (default: torch's global CPU
    generator) in the reference's order."""
    T = gauss_tables(p)
    S = p.sampling_timesteps if num_sample_steps is None else num_sample_steps
    _n, _c, h, w = condition_x.shape
    condition_x = condition_x * 2 - 1
    shape = (batch_size, spec.channels, h, w)

    def start(target_time):
        if generation_start_steps > 0:
            t = torch.tensor([target_time] * batch_size, device=condition_x.device).long()
            noise = torch.randn(condition_x.shape, generator=generator, device=condition_x.device)
            return (_gx(T["sqrt_alphas_cumprod"], t, condition_x) * condition_x +
                    _gx(T["sqrt_one_minus_alphas_cumprod"], t, condition_x) * noise)            # q_sample (pip, unpinned)
        return torch.randn(shape, generator=generator, device=condition_x.device)

    def scales(i):
        return (1.0 if i < guidance_start_steps else cond_scale,
                1.0 if i < class_guidance_start_steps else class_cond_scale)

    if p.sampling_timesteps >= p.timesteps:                                   # is_ddim_sampling False, model.py:1388
        img = start(p.timesteps - generation_start_steps)
        for i, t in enumerate(reversed(range(0, p.timesteps))):
            if i < generation_start_steps:
                continue
            cs, ccs = scales(i)
            img, _ = gauss_p_sample(sd, spec, p, T, img, t, condition_x, class_label, cs, ccs, generator=generator)
            if model_steps is not None:
                model_steps.append(img.clone())
        return (img + 1) * 0.5
    times = torch.linspace(-1, p.timesteps - 1, steps=S + 1)
    times = list(reversed(times.int().tolist()))
    pairs = list(zip(times[:-1], times[1:]))
    img = start(pairs[generation_start_steps][0] if generation_start_steps > 0 else 0)
    for i, (time, time_next) in enumerate(pairs):
        if i < generation_start_steps:
            continue
        cs, ccs = scales(i)
        img, _ = gauss_ddim_step(sd, spec, p, T, img, time, time_next, condition_x, class_label, cs, ccs,
                                 generator=generator)
        if model_steps is not None:
            model_steps.append(img.clone())
    return (img + 1) * 0.5
