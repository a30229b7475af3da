"""Checkpoint ingest: the reference's fp32 state dict ({'ema_model': ...}, model.py:3659-3664) ->
the packed device tensors `srgd_unet_create` expects (names from `srgd_unet_param_name`).

Pure data movement (permutes, casts, concatenations) done once per load with torch as plumbing:
  * conv weights OIHW fp32 -> [O][kh][kw][I] bf16 (K-major rows for the implicit GEMM);
  * init conv 7x7 -> [O][ky][64] with column kx*6+c (row-im2col layout of srgd_pack_input);
  * Downsample 1x1: input channel (c p1 p2) -> (p1 p2 c)  (model.py:108);
  * PixelShuffle 1x1: output channel (c' i j) -> (i j c') (model.py:83);
  * the pre-attention RMSNorm gain g*sqrt(C) folded into the to_qkv weight columns (model.py:207, 310);
  * all ResnetBlock time-MLP matrices concatenated into one [sum 2C][4 dim] matrix (model.py:264-267);
  * class_mlp evaluated for every class once (3 rows) with srgd_dense_rows (model.py:612-619).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from .arch import UnetSpec


def make_config(spec: UnetSpec) -> _lib.UnetConfig:
    cfg = _lib.UnetConfig()
    cfg.dim = spec.dim
    cfg.n_stages = len(spec.dim_mults)
    for i, m in enumerate(spec.dim_mults):
        cfg.dim_mults[i] = int(m)
        cfg.full_attn[i] = int(bool(spec.full_attn[i]))
    cfg.heads, cfg.dim_head, cfg.groups = spec.heads, spec.dim_head, spec.groups
    cfg.channels, cfg.sinu_dim = spec.channels, spec.learned_sinusoidal_dim
    cfg.num_classes = int(spec.num_classes or 0)
    cfg.fixed_sinusoidal = 0 if spec.learned_sinusoidal_cond else 1
    return cfg


def param_names(spec: UnetSpec) -> List[str]:
    lib = _lib.load()
    cfg = make_config(spec)
    n = lib.srgd_unet_param_count(C.byref(cfg))
    if n <= 0:
        raise _lib.SrgdError(f"unsupported U-Net configuration: {_lib.last_error()}")
    return [lib.srgd_unet_param_name(C.byref(cfg), i).decode() for i in range(n)]


def _conv_kmajor(w: torch.Tensor) -> torch.Tensor:
    o, i, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(o, kh * kw * i)


def _res_order(spec: UnetSpec) -> List[str]:
    n = len(spec.dim_mults)
    names = []
    for i in range(n):
        names += [f"downs.{i}.0", f"downs.{i}.1"]
    names += ["mid_block1", "mid_block2"]
    for i in range(n):
        names += [f"ups.{i}.0", f"ups.{i}.1"]
    names.append("final_res_block")
    return names


def pack(spec: UnetSpec, sd: Dict[str, torch.Tensor], device: torch.device) -> Dict[str, torch.Tensor]:
    """sd: U-Net state dict (no `model.` prefix), fp32, any device.  Returns name -> device tensor."""
    lib = _lib.load()
    n = len(spec.dim_mults)
    g = lambda k: sd[k].detach().to(device=device, dtype=torch.float32)
    bf = lambda t: t.to(torch.bfloat16).contiguous()
    f32 = lambda t: t.to(torch.float32).contiguous()
    out: Dict[str, torch.Tensor] = {}

    w = g("init_conv.weight")                                   # [O, 6, 7(ky), 7(kx)]
    o = w.shape[0]
    wi = torch.zeros(o, 7, 64, device=device)
    wi[:, :, :42] = w.permute(0, 2, 3, 1).reshape(o, 7, 42)     # (ky, kx, c) -> column kx*6 + c
    out["init.w"] = bf(wi.reshape(o, 7 * 64))
    out["init.b"] = f32(g("init_conv.bias"))
    if spec.learned_sinusoidal_cond:
        out["time.freq"] = f32(g("time_mlp.0.weights"))
    else:
        # SinusoidalPosEmb's frequency table with the reference's own ops (model.py:216-218), on the host
        half = spec.dim // 2
        emb = math.log(10000) / (half - 1)
        out["time.freq"] = torch.exp(torch.arange(half) * -emb).to(device=device, dtype=torch.float32).contiguous()
    out["time.w1"], out["time.b1"] = f32(g("time_mlp.1.weight")), f32(g("time_mlp.1.bias"))
    out["time.w2"], out["time.b2"] = f32(g("time_mlp.3.weight")), f32(g("time_mlp.3.bias"))

    if spec.num_classes:
        emb = f32(g("class_mlp.0.weight"))
        w1, b1 = f32(g("class_mlp.1.weight")), f32(g("class_mlp.1.bias"))
        w3, b3 = f32(g("class_mlp.3.weight")), f32(g("class_mlp.3.bias"))
        td = spec.time_dim
        h1 = torch.empty(spec.num_classes, td, device=device)
        table = torch.empty(spec.num_classes, td, device=device)
        st = _lib.current_stream()
        _lib.check(lib.srgd_dense_rows(_lib.ptr(emb), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(h1), spec.num_classes, td,
                                       spec.dim, 0, 0, st), "class_mlp.1")
        _lib.check(lib.srgd_dense_rows(_lib.ptr(h1), _lib.ptr(w3), _lib.ptr(b3), _lib.ptr(table), spec.num_classes,
                                       td, td, 2, 0, st), "class_mlp.3")
        torch.cuda.current_stream().synchronize()               # emb/w1/... may be freed after return
        out["class.table"] = table

    order = _res_order(spec)
    out["ss.w"] = f32(torch.cat([g(f"{r}.mlp.1.weight") for r in order], 0))
    out["ss.b"] = f32(torch.cat([g(f"{r}.mlp.1.bias") for r in order], 0))

    def res(name):
        for short, blk in (("1", "block1"), ("2", "block2")):
            out[f"{name}.c{short}.w"] = bf(_conv_kmajor(g(f"{name}.{blk}.proj.weight")))
            out[f"{name}.c{short}.b"] = f32(g(f"{name}.{blk}.proj.bias"))
            out[f"{name}.n{short}.g"] = f32(g(f"{name}.{blk}.norm.weight"))
            out[f"{name}.n{short}.b"] = f32(g(f"{name}.{blk}.norm.bias"))
        if f"{name}.res_conv.weight" in sd:
            out[f"{name}.res.w"] = bf(_conv_kmajor(g(f"{name}.res_conv.weight")))
            out[f"{name}.res.b"] = f32(g(f"{name}.res_conv.bias"))

    def attn(name, full):
        gain = g(f"{name}.norm.g").reshape(-1)
        c = gain.numel()
        wq = g(f"{name}.to_qkv.weight").reshape(-1, c)
        out[f"{name}.qkv.w"] = bf(wq * (gain * math.sqrt(c))[None, :])
        key = f"{name}.to_out" if full else f"{name}.to_out.0"
        out[f"{name}.out.w"] = bf(g(f"{key}.weight").reshape(c, -1))
        out[f"{name}.out.b"] = f32(g(f"{key}.bias"))
        if not full:
            out[f"{name}.out.g"] = f32(g(f"{name}.to_out.1.g").reshape(-1))

    for i in range(n):
        res(f"downs.{i}.0"); res(f"downs.{i}.1"); attn(f"downs.{i}.2", spec.full_attn[i])
        if i < n - 1:
            wd = g(f"downs.{i}.3.1.weight")                     # [O, (c p1 p2), 1, 1]
            o_, k4 = wd.shape[0], wd.shape[1]
            out[f"downs.{i}.3.w"] = bf(wd.reshape(o_, k4 // 4, 2, 2).permute(0, 2, 3, 1).reshape(o_, k4))
            out[f"downs.{i}.3.b"] = f32(g(f"downs.{i}.3.1.bias"))
        else:
            out[f"downs.{i}.3.w"] = bf(_conv_kmajor(g(f"downs.{i}.3.weight")))
            out[f"downs.{i}.3.b"] = f32(g(f"downs.{i}.3.bias"))
    res("mid_block1"); attn("mid_attn", True); res("mid_block2")
    for i in range(n):
        res(f"ups.{i}.0"); res(f"ups.{i}.1"); attn(f"ups.{i}.2", spec.full_attn[n - 1 - i])
        if i < n - 1:
            wu = g(f"ups.{i}.3.net.0.weight")                   # [(c' i j), C, 1, 1]
            o4, c = wu.shape[0], wu.shape[1]
            out[f"ups.{i}.3.w"] = bf(wu.reshape(o4 // 4, 4, c).permute(1, 0, 2).reshape(o4, c))
            out[f"ups.{i}.3.b"] = f32(g(f"ups.{i}.3.net.0.bias").reshape(o4 // 4, 4).permute(1, 0).reshape(o4))
        else:
            out[f"ups.{i}.3.w"] = bf(_conv_kmajor(g(f"ups.{i}.3.weight")))
            out[f"ups.{i}.3.b"] = f32(g(f"ups.{i}.3.bias"))
    res("final_res_block")
    out["final.w"] = f32(g("final_conv.weight").reshape(spec.channels, spec.dim))
    out["final.b"] = f32(g("final_conv.bias"))
    return out


# ---------------------------------------------------------------------------------------------------
# Ingest cache.  The reference pays `torch.load` of the 550 MB fp32 checkpoint plus module
# construction on every start (model.py:3657-3664); here the packed tensors of pack() -- ~276 MB, bf16 K-major conv
# weights, fused embedding matrices, the evaluated class table -- are written next to the checkpoint after the first
# load and memory-mapped on later starts, which then skip torch.load, load_state_dict and the repack.
# ---------------------------------------------------------------------------------------------------
PACK_FORMAT = 2                      # bump whenever pack()'s output layout changes


def pack_cache_path(ckpt_path: str) -> str:
    return ckpt_path + ".srgd_b200_pack"


def file_sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 24), b""):
            h.update(chunk)
    return h.hexdigest()


def _ckpt_identity(ckpt_path: str) -> Dict[str, int]:
    st = os.stat(ckpt_path)
    return dict(size=int(st.st_size), mtime_ns=int(st.st_mtime_ns))


def _spec_key(spec: UnetSpec) -> str:
    return repr((spec.dim, tuple(spec.dim_mults), spec.channels, spec.groups, spec.learned_sinusoidal_dim, spec.heads,
                 spec.dim_head, tuple(spec.full_attn), spec.num_classes) +
                (() if spec.learned_sinusoidal_cond else ("fixed_sinusoidal",)))


def save_pack_cache(ckpt_path: str, spec: UnetSpec, packed: Dict[str, torch.Tensor]) -> Optional[str]:
    """Writes <ckpt>.srgd_b200_pack (atomically).  The cache is keyed by the checkpoint's sha256, recorded together
    with its size / mtime so that later starts can validate it without re-reading 550 MB.  Returns the path, or None
    if the directory is not writable."""
    path = pack_cache_path(ckpt_path)
    blob = dict(format=PACK_FORMAT, spec=_spec_key(spec), sha256=file_sha256(ckpt_path), **_ckpt_identity(ckpt_path),
                tensors={k: v.detach().to("cpu").contiguous() for k, v in packed.items()})
    tmp = f"{path}.tmp{os.getpid()}"
    try:
        torch.save(blob, tmp)
        os.replace(tmp, path)
    except OSError:
        try:
            os.unlink(tmp)
        except OSError:
            pass
        return None
    return path


def load_pack_cache(ckpt_path: str, spec: UnetSpec, verify_sha256: Optional[bool] = None):
    """The cached pack {name: CPU tensor (memory-mapped)} of `ckpt_path`, or None when there is no valid cache: missing
    file, other pack format / architecture, or a checkpoint whose size or mtime changed (then the sha256 decides: an
    identical file that was merely copied or touched keeps its cache).  SRGD_B200_VERIFY_CACHE=1 re-hashes always."""
    path = pack_cache_path(ckpt_path)
    if not os.path.exists(path) or not os.path.exists(ckpt_path):
        return None
    try:
        blob = torch.load(path, map_location="cpu", weights_only=True, mmap=True)
    except Exception:
        return None
    if not isinstance(blob, dict) or blob.get("format") != PACK_FORMAT or blob.get("spec") != _spec_key(spec):
        return None
    if verify_sha256 is None:
        verify_sha256 = os.environ.get("SRGD_B200_VERIFY_CACHE", "0") == "1"
    ident = _ckpt_identity(ckpt_path)
    same_stat = blob.get("size") == ident["size"] and blob.get("mtime_ns") == ident["mtime_ns"]
    if blob.get("size") != ident["size"]:
        return None
    if (verify_sha256 or not same_stat) and file_sha256(ckpt_path) != blob.get("sha256"):
        return None
    return blob["tensors"]
