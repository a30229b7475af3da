"""Architecture description of the denoiser get_model() builds for `conditional_continuous`
(reference: model.py:3500-3515 with config.py defaults) and its checkpoint key layout
(model.py:583-675; 280 tensors for the shipped conf)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple


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
    # True: RandomOrLearnedSinusoidalPosEmb on log-SNR / c_noise (continuous-time and EDM families, model.py:596-598);
    # False: the fixed SinusoidalPosEmb(dim) on integer timesteps (discrete-time family, model.py:600)
    learned_sinusoidal_cond: bool = True

    @property
    def widths(self) -> List[int]:
        return [self.dim] + [self.dim * m for m in self.dim_mults]

    @property
    def time_dim(self) -> int:
        return 4 * self.dim

    @property
    def hidden(self) -> int:
        return self.heads * self.dim_head

    @property
    def downsample_factor(self) -> int:
        return 2 ** (len(self.dim_mults) - 1)


def _res(keys: Dict, name: str, cin: int, cout: int, tdim: int):
    keys[f"{name}.mlp.1.weight"] = (2 * cout, tdim)
    keys[f"{name}.mlp.1.bias"] = (2 * cout,)
    for blk, ci in (("block1", cin), ("block2", cout)):
        keys[f"{name}.{blk}.proj.weight"] = (cout, ci, 3, 3)
        keys[f"{name}.{blk}.proj.bias"] = (cout,)
        keys[f"{name}.{blk}.norm.weight"] = (cout,)
        keys[f"{name}.{blk}.norm.bias"] = (cout,)
    if cin != cout:
        keys[f"{name}.res_conv.weight"] = (cout, cin, 1, 1)
        keys[f"{name}.res_conv.bias"] = (cout,)


def _attn(keys: Dict, name: str, c: int, hidden: int, full: bool):
    keys[f"{name}.norm.g"] = (1, c, 1, 1)
    keys[f"{name}.to_qkv.weight"] = (3 * hidden, c, 1, 1)
    out = f"{name}.to_out" if full else f"{name}.to_out.0"
    keys[f"{out}.weight"] = (c, hidden, 1, 1)
    keys[f"{out}.bias"] = (c,)
    if not full:
        keys[f"{name}.to_out.1.g"] = (1, c, 1, 1)


def unet_keys(spec: UnetSpec) -> Dict[str, Tuple[int, ...]]:
    """U-Net state-dict entries (without the diffusion wrapper's `model.` prefix), in the
    registration order of the reference module tree."""
    k: Dict[str, Tuple[int, ...]] = {}
    w, td, hid = spec.widths, spec.time_dim, spec.hidden
    n = len(spec.dim_mults)
    k["init_conv.weight"] = (spec.dim, 2 * spec.channels, 7, 7)
    k["init_conv.bias"] = (spec.dim,)
    if spec.learned_sinusoidal_cond:
        k["time_mlp.0.weights"] = (spec.learned_sinusoidal_dim // 2,)
        k["time_mlp.1.weight"] = (td, spec.learned_sinusoidal_dim + 1)
    else:                                            # SinusoidalPosEmb has no parameters; fourier_dim = dim
        k["time_mlp.1.weight"] = (td, spec.dim)
    k["time_mlp.1.bias"] = (td,)
    k["time_mlp.3.weight"] = (td, td)
    k["time_mlp.3.bias"] = (td,)
    if spec.num_classes is not None:
        k["class_mlp.0.weight"] = (spec.num_classes, spec.dim)
        k["class_mlp.1.weight"] = (td, spec.dim)
        k["class_mlp.1.bias"] = (td,)
        k["class_mlp.3.weight"] = (td, td)
        k["class_mlp.3.bias"] = (td,)
    for i in range(n):
        cin, cout = w[i], w[i + 1]
        _res(k, f"downs.{i}.0", cin, cin, td)
        _res(k, f"downs.{i}.1", cin, cin, td)
        _attn(k, f"downs.{i}.2", cin, hid, spec.full_attn[i])
        if i < n - 1:
            k[f"downs.{i}.3.1.weight"] = (cout, 4 * cin, 1, 1)
            k[f"downs.{i}.3.1.bias"] = (cout,)
        else:
            k[f"downs.{i}.3.weight"] = (cout, cin, 3, 3)
            k[f"downs.{i}.3.bias"] = (cout,)
    for i in range(n):
        cin, cout = w[n - 1 - i], w[n - i]
        _res(k, f"ups.{i}.0", cout + cin, cout, td)
        _res(k, f"ups.{i}.1", cout + cin, cout, td)
        _attn(k, f"ups.{i}.2", cout, hid, spec.full_attn[n - 1 - i])
        if i < n - 1:
            k[f"ups.{i}.3.net.0.weight"] = (4 * cin, cout, 1, 1)
            k[f"ups.{i}.3.net.0.bias"] = (4 * cin,)
        else:
            k[f"ups.{i}.3.weight"] = (cin, cout, 3, 3)
            k[f"ups.{i}.3.bias"] = (cin,)
    _res(k, "mid_block1", w[-1], w[-1], td)
    _attn(k, "mid_attn", w[-1], hid, True)
    _res(k, "mid_block2", w[-1], w[-1], td)
    _res(k, "final_res_block", 2 * spec.dim, spec.dim, td)
    k["final_conv.weight"] = (spec.channels, spec.dim, 1, 1)
    k["final_conv.bias"] = (spec.channels,)
    return k


def seeded_state_dict(spec: UnetSpec, seed: int = 1234, prefix: str = "model.", init: str = "unit"):
    """Deterministic random-init checkpoint contents `{prefix + key: fp32 tensor}` of the reference architecture, for
    benchmarks and smoke runs when the shipped .pth is only a Git-LFS pointer.

    init="torch": the distributions the reference constructors produce (nn.Conv2d / nn.Linear defaults U(+-1/sqrt(fan_in))
    for weight and bias, norm gains 1 / biases 0, nn.Embedding and the sinusoidal frequencies N(0,1), the pixel-shuffle
    conv = kaiming-uniform rows repeated x4 with zero bias, model.py:88-95).
    init="unit": a harsher stress init (unit-gain weights, jittered norm gains, non-zero biases) that exercises every
    gain / bias path with O(1..3) activations; this is what bench.py loads.
    Draw order = checkpoint key order, so the tensors equal the ones the parity tests load, bit for bit
    (tests/test_host_logic.py checks that)."""
    import math

    import torch
    g = torch.Generator().manual_seed(seed)
    keys = unet_keys(spec)
    sd = {}
    for name, shape in keys.items():
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
                fi = 1
                for s in keys[name[:-4] + "weight"][1:]:
                    fi *= s
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fi)
        elif init == "torch" and ".net.0.weight" in name:
            o, i, kh, kw = shape
            base = (torch.rand((o // 4, i, kh, kw), generator=g) * 2 - 1) * math.sqrt(6.0 / (i * kh * kw))
            t = base.repeat_interleave(4, dim=0)
        else:
            bound = math.sqrt(3.0 / fan_in) if init == "unit" else 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[prefix + name] = t.float().contiguous()
    return sd
