"""Host-side mirror of the reference's `ConditionalSRUnet` (model.py:536-725).

Same constructor keywords, attributes (`channels`, `self_condition`, `num_classes`,
`random_or_learned_sinusoidal_cond`, `downsample_factor`), state-dict keys and
`forward(x, time, class_label=None, x_self_cond=None) -> eps` contract.  The module owns fp32
parameters only so that `load_state_dict(ckpt['ema_model'])`, `.to(device)` and `.eval()` behave
like the reference's nn.Module; every FLOP of forward() runs in libsrgd_b200.so through
`srgd_unet_forward` (tcgen05 implicit-GEMM convolutions, fused norm / attention kernels).
"""
from __future__ import annotations

import copy
import ctypes as C
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, weights
from .arch import UnetSpec, unet_keys


class _Node(nn.Module):
    """Anonymous container used to reproduce the reference's dotted parameter names."""


def _attach(root: nn.Module, dotted: str, param: nn.Parameter) -> None:
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, _Node())
        node = node._modules[p]
    node.register_parameter(parts[-1], param)


def _init_like_torch(name: str, shape: Tuple[int, ...], gen: torch.Generator) -> torch.Tensor:
    """Random init in the spirit of torch's defaults (the reference relies on nn.Conv2d / nn.Linear
    defaults); real use always loads a checkpoint over it."""
    if name.endswith("time_mlp.0.weights") or name.endswith("class_mlp.0.weight"):
        return torch.randn(shape, generator=gen)
    if name.endswith(".g") or name.endswith("norm.weight"):
        return torch.ones(shape)
    if name.endswith("norm.bias"):
        return torch.zeros(shape)
    fan_in = 1
    for s in (shape[1:] if len(shape) > 1 else shape):
        fan_in *= s
    if name.endswith(".bias"):
        return torch.zeros(shape)
    bound = (1.0 / fan_in) ** 0.5
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


class ConditionalSRUnet(nn.Module):
    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3,
                 self_condition=True, resnet_block_groups=8, learned_variance=False,
                 learned_sinusoidal_cond=False, random_fourier_features=False, learned_sinusoidal_dim=16,
                 attn_dim_head=32, attn_heads=4, full_attn=(False, False, False, True), flash_attn=False,
                 pixel_shuffle_upsample=True, num_classes=None, _init_weights=True):
        super().__init__()
        unsupported = []
        if init_dim not in (None, dim): unsupported.append("init_dim != dim")
        if out_dim not in (None, channels): unsupported.append("out_dim")
        if not self_condition: unsupported.append("self_condition=False")
        if learned_variance: unsupported.append("learned_variance=True")
        if not pixel_shuffle_upsample: unsupported.append("pixel_shuffle_upsample=False")
        if unsupported:
            raise NotImplementedError("srgd_b200 builds the class-conditional U-Net of the shipped configuration; unsupported: "
                                      + ", ".join(unsupported))
        if isinstance(full_attn, bool):
            full_attn = (full_attn,) * len(dim_mults)
        assert len(full_attn) == len(dim_mults)
        self.spec = UnetSpec(dim=dim, dim_mults=tuple(int(m) for m in dim_mults), channels=channels,
                             groups=resnet_block_groups, learned_sinusoidal_dim=learned_sinusoidal_dim,
                             heads=attn_heads, dim_head=attn_dim_head, full_attn=tuple(bool(f) for f in full_attn),
                             num_classes=num_classes,
                             learned_sinusoidal_cond=bool(learned_sinusoidal_cond or random_fourier_features))
        self.channels = channels
        self.self_condition = self_condition
        self.num_classes = num_classes
        self.out_dim = channels
        self.random_or_learned_sinusoidal_cond = self.spec.learned_sinusoidal_cond         # model.py:594
        self.downsample_factor = self.spec.downsample_factor
        gen = torch.Generator().manual_seed(0)
        for name, shape in unet_keys(self.spec).items():
            # _init_weights=False (get_model with a checkpoint to load): uninitialised storage, no 1.5 s of random init
            t = _init_like_torch(name, shape, gen) if _init_weights else torch.empty(shape)
            _attach(self, name, nn.Parameter(t, requires_grad=False))
        # Ingest cache (weights.load_pack_cache): the packed tensors of an earlier start.  While `_deferred_ckpt` is
        # set, the fp32 parameters are uninitialised placeholders -- the forward only ever reads the pack -- and are
        # filled from the checkpoint the first time anybody asks for them (state_dict()).
        self._cached_pack: Optional[Dict[str, torch.Tensor]] = None
        self._deferred_ckpt: Optional[str] = None
        self._save_pack_for: Optional[str] = None
        self._handle = None
        self._packed: Optional[Dict[str, torch.Tensor]] = None
        self._packed_key = None
        # A load through a PARENT module (`diffusion.load_state_dict(ckpt['ema_model'])`, the documented path) never
        # calls this module's load_state_dict, only its _load_from_state_dict + post hooks: drop the packed device
        # weights there too.  In-place edits of a parameter are caught by the version key in _ensure_handle.
        self.register_load_state_dict_post_hook(ConditionalSRUnet._after_load)
        self._workspaces: Dict[Tuple[int, int, int], torch.Tensor] = {}
        self.conv_impl = 0          # debug knob: 1 = CUDA-core direct conv, 2 = stand-alone GN statistics
        self.last_launches = 0

    _RUNTIME_FIELDS = ("_handle", "_packed", "_packed_key", "_workspaces", "_plist", "_labels_ok", "_save_pack_for")

    def __deepcopy__(self, memo):
        # device handles / packed weights are per-instance runtime state: a copy re-packs lazily
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in self._RUNTIME_FIELDS else copy.deepcopy(v, memo)
        new.__dict__["_workspaces"] = {}
        return new

    # -- parameter packing ---------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._drop_handle()
        return r

    def _apply(self, fn, *a, **kw):
        if self._deferred_ckpt is not None:
            # placeholders: re-create them where fn would put them instead of copying 550 MB of nothing
            probe = fn(torch.empty(1, device=next(self.parameters()).device))
            with torch.no_grad():
                for p in self.parameters():
                    p.data = torch.empty(p.shape, device=probe.device, dtype=probe.dtype)
            r = self
        else:
            r = super()._apply(fn, *a, **kw)
        self._drop_handle()
        return r

    @staticmethod
    def _after_load(module, incompatible_keys) -> None:
        module._cached_pack = None               # new weights: an attached ingest cache no longer describes them
        module._deferred_ckpt = None
        module._drop_handle()

    def attach_pack_cache(self, pack: Dict[str, torch.Tensor], ckpt_path: str, prefix: str = "model.") -> None:
        """Start from the ingest cache of `ckpt_path` (weights.load_pack_cache) instead of its fp32 state dict;
        `prefix` = the U-Net's key prefix inside ckpt['ema_model'] ("model." / "net." by sampler family)."""
        self._drop_handle()
        self._cached_pack, self._deferred_ckpt = pack, ckpt_path
        self.__dict__["_ckpt_prefix"] = prefix

    def save_pack_cache_after_first_pack(self, ckpt_path: str) -> None:
        self._save_pack_for = ckpt_path

    def _materialize(self) -> None:
        """Fill the placeholder parameters from the checkpoint the ingest cache was made from."""
        if self._deferred_ckpt is None:
            return
        path, self._deferred_ckpt = self._deferred_ckpt, None
        sd = torch.load(path, map_location="cpu", weights_only=True)["ema_model"]
        prefix = self.__dict__.get("_ckpt_prefix", "model.")
        own = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        pack, handle_state = self._cached_pack, (self._handle, self._packed, self._packed_key)
        with torch.no_grad():
            for k, p in super().state_dict(keep_vars=True).items():
                p.data.copy_(own[k])
        # same weights as the pack: keep using it (copy_ bumped the version counters, so refresh the key)
        self._cached_pack = pack
        self._handle, self._packed, _ = handle_state
        if self._handle is not None:
            self._packed_key = self._weights_key(self._packed_key[0])

    def state_dict(self, *a, **kw):
        self._materialize()
        return super().state_dict(*a, **kw)

    def _drop_handle(self):
        if getattr(self, "_handle", None) is not None:
            _lib.load().srgd_unet_destroy(self._handle)
        self._handle = None
        self._packed = None
        self._workspaces = {}
        self.__dict__["_plist"] = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _weights_key(self, device):
        """(device, sum of the parameters' in-place version counters): any optimizer-style or manual in-place
        change of a weight bumps its `_version`, so a stale pack is never reused."""
        v = 0
        plist = self.__dict__.get("_plist")
        if plist is None:                  # flat parameter list (walking the module tree costs ~0.7 ms per call)
            plist = self.__dict__["_plist"] = list(self.parameters())
        try:
            for p in plist:
                v += p._version
        except RuntimeError:               # inference tensors carry no version counter (and cannot be edited in place)
            v = -1
        return (device, v)

    def _ensure_handle(self, device: torch.device):
        key = self._weights_key(device)
        if self._handle is not None and self._packed_key == key:
            return
        if self._handle is not None:
            self._drop_handle()
        lib = _lib.load()
        _lib.check(lib.srgd_device_check(device.index if device.index is not None else torch.cuda.current_device()),
                   "srgd_device_check")
        with torch.cuda.device(device):
            if self._cached_pack is not None:
                packed = {k: v.to(device, non_blocking=False) for k, v in self._cached_pack.items()}
            else:
                sd = {k: v for k, v in self.state_dict().items()}
                packed = weights.pack(self.spec, sd, device)
                if self._save_pack_for is not None:
                    path, self._save_pack_for = self._save_pack_for, None
                    weights.save_pack_cache(path, self.spec, packed)
            names = weights.param_names(self.spec)
            missing = [n for n in names if n not in packed]
            if missing:
                raise _lib.SrgdError(f"weight packer did not produce {missing[:4]}...")
            arr = (C.c_void_p * len(names))(*[packed[n].data_ptr() for n in names])
            cfg = weights.make_config(self.spec)
            handle = C.c_void_p()
            _lib.check(lib.srgd_unet_create(C.byref(cfg), arr, len(names), C.byref(handle)), "srgd_unet_create")
        self._handle, self._packed, self._packed_key = handle, packed, key
        self._workspaces = {}

    def _workspace(self, B: int, H: int, W: int, device) -> torch.Tensor:
        key = (B, H, W)
        ws = self._workspaces.get(key)
        if ws is None:
            nbytes = _lib.load().srgd_unet_workspace_bytes(self._handle, B, H, W)
            if nbytes == 0:
                raise _lib.SrgdError(f"unsupported U-Net input shape: {_lib.last_error()}")
            # keep only the largest workspace alive; smaller shapes reuse it
            big = max(self._workspaces.values(), key=lambda t: t.numel(), default=None)
            if big is not None and big.numel() >= nbytes:
                ws = big
            else:
                self._workspaces = {}
                ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._workspaces[key] = ws
        return ws

    # -- the hot call ----------------------------------------------------------------------------
    def run(self, x: torch.Tensor, log_snr: torch.Tensor, labels_i32: Optional[torch.Tensor],
            cond: Optional[torch.Tensor], rows: int, n_cond_rows: int, out: Optional[torch.Tensor] = None):
        """eps[rows,3,H,W]; row b reads x[b % Bx] / cond[b % Bx] (cond only for b < n_cond_rows) with
        label labels_i32[b] (<0 = null) and log_snr[b].  rows = Bx, or 2*Bx for the CFG pair."""
        # the reference fails in torch.cat((x, x_self_cond), dim=1) / the init conv for these (model.py:684-686); here
        # they would be out-of-bounds reads of the raw pointers
        if x.ndim != 4 or x.shape[1] != self.channels:
            raise RuntimeError(f"expected x of shape [B, {self.channels}, H, W], got {tuple(x.shape)}")
        assert all(d % self.downsample_factor == 0 for d in x.shape[-2:]), \
            f'your input dimensions {x.shape[-2:]} need to be divisible by {self.downsample_factor}, given the unet'
        if cond is not None and tuple(cond.shape) != tuple(x.shape):
            raise RuntimeError(f"Sizes of tensors must match: x {tuple(x.shape)} vs condition {tuple(cond.shape)}")
        if cond is not None and cond.device != x.device:
            raise RuntimeError(f"Expected all tensors to be on the same device, got {x.device} and {cond.device}")
        if rows not in (x.shape[0], 2 * x.shape[0]):
            raise RuntimeError(f"rows={rows} must be the batch of x ({x.shape[0]}) or twice that (guidance pair)")
        if log_snr.numel() != rows or (labels_i32 is not None and labels_i32.numel() != rows):
            raise RuntimeError(f"time has {log_snr.numel()} entries"
                               + ("" if labels_i32 is None else f", class_label {labels_i32.numel()}")
                               + f" for a batch of {rows}")
        _lib.require_cuda(x, "ConditionalSRUnet")
        Bx, _, H, W = x.shape
        self._ensure_handle(x.device)
        x = x.contiguous().float()
        if cond is not None:
            cond = cond.contiguous().float()
        log_snr = log_snr.to(x.device).contiguous().float()
        if out is None:
            out = torch.empty(rows, self.channels, H, W, device=x.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(x.device):
            ws = self._workspace(rows, H, W, x.device)
            rc = lib.srgd_unet_forward(self._handle, _lib.ptr(x), _lib.ptr(cond), _lib.ptr(log_snr),
                                       _lib.ptr(labels_i32), n_cond_rows, Bx, _lib.ptr(out), rows, H, W,
                                       _lib.ptr(ws), ws.numel(), int(self.conv_impl), _lib.current_stream())
        _lib.check(rc, "srgd_unet_forward")
        self.last_launches = lib.srgd_unet_last_launch_count(self._handle)
        return out

    def labels_for(self, class_label: Optional[torch.Tensor], rows: int, device) -> Optional[torch.Tensor]:
        """int32 device labels for `rows` rows from the reference-style class_label ([B] or [1] int64)."""
        if class_label is None or self.num_classes is None:
            return None
        lab = class_label.reshape(-1)
        # nn.Embedding raises IndexError for an out-of-range label (model.py:612, 693); negative values are reserved
        # for the internally generated null rows of the guidance batch and never come from a caller
        self._check_labels(lab)
        lab = lab.to(device=device).to(torch.int32)
        if lab.numel() == 1 and rows != 1:
            lab = lab.expand(rows)
        if lab.numel() != rows:
            raise RuntimeError(f"class_label has {lab.numel()} entries for a batch of {rows}")
        return lab.contiguous()

    def _check_labels(self, lab: torch.Tensor) -> None:
        """Host labels are validated on every call; a device label costs one sync, so it is validated once per
        distinct tensor (same storage, same version) -- sampling loops pass the same label tensor every step."""
        key = None
        if lab.is_cuda:
            try:
                ver = lab._version
            except RuntimeError:           # created under inference_mode: immutable
                ver = -1
            key = (lab.data_ptr(), ver, lab.numel(), lab.device.index)
            if key == getattr(self, "_labels_ok", None):
                return
        if lab.numel() and (int(lab.min()) < 0 or int(lab.max()) >= self.num_classes):
            raise IndexError(f"class_label {lab.tolist()[:8]} out of range for num_classes={self.num_classes}")
        self._labels_ok = key

    def forward(self, x, time, class_label=None, x_self_cond=None):
        assert all(d % self.downsample_factor == 0 for d in x.shape[-2:]), \
            f'your input dimensions {x.shape[-2:]} need to be divisible by {self.downsample_factor}, given the unet'
        B = x.shape[0]
        labels = self.labels_for(class_label, B, x.device)
        return self.run(x, time.to(x.device).reshape(-1).expand(B) if time.numel() == 1 else time.to(x.device),
                        labels, x_self_cond, B, B if x_self_cond is not None else 0)
