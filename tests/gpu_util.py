"""Helpers shared by the GPU parity tests (call the C-ABI directly through ctypes)."""
import ctypes as C

import torch

from srgd_b200 import _lib


_alive = []


def P(t):
    """Device pointer of `t`, keeping `t` alive until release() (ctypes pointers do not own storage)."""
    if t is None:
        return None
    _alive.append(t)
    return _lib.ptr(t)


def release():
    torch.cuda.synchronize()
    _alive.clear()


def stream():
    return _lib.current_stream()


def nhwc_bf16(x_nchw: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW (cpu or cuda) -> bf16 NHWC contiguous on cuda."""
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(device="cuda", dtype=torch.bfloat16)


def to_nchw_f32(x_nhwc: torch.Tensor) -> torch.Tensor:
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous().cpu()


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float()


def splitk_workspace():
    """A split-K workspace for srgd_conv_desc.splitk_ws (flags zeroed once; the kernels re-arm them)."""
    n = _lib.load().srgd_conv_splitk_workspace_bytes()
    return torch.zeros(n, dtype=torch.uint8, device="cuda")


def conv_desc(srcs, phases, weight, Ktot, B, Ho, Wo, Cout, out, bias=None, row_scale=None, residual=None, act=0,
              out_mode=0, gn_partials=None, splitk_ws=None):
    """srcs: list of (tensor_or_ptr, sb, sy, sx, H, W, C); phases: list of (src, dy, dx, k_start)."""
    d = _lib.ConvDesc()
    d.B, d.Ho, d.Wo, d.Cout = B, Ho, Wo, Cout
    d.n_src, d.n_phase = len(srcs), len(phases)
    for i, (t, sb, sy, sx, H, W, Cc) in enumerate(srcs):
        p = t if isinstance(t, int) else t.data_ptr()
        d.srcs[i] = _lib.ConvSrc(p, sb, sy, sx, H, W, Cc)
    for i, (s, dy, dx, k) in enumerate(phases):
        d.phases[i] = _lib.ConvPhase(s, dy, dx, k)
    d.weight = weight.data_ptr()
    d.Ktot = Ktot
    d.bias = bias.data_ptr() if bias is not None else None
    d.row_scale = row_scale.data_ptr() if row_scale is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.act, d.out_mode = act, out_mode
    d.out = out.data_ptr()
    d.gn_partials = gn_partials.data_ptr() if gn_partials is not None else None
    d.splitk_ws = splitk_ws.data_ptr() if splitk_ws is not None else None
    d.splitk_ws_bytes = splitk_ws.numel() if splitk_ws is not None else 0
    d._keep = [srcs, weight, out, bias, row_scale, residual, gn_partials, splitk_ws]     # keep the storages alive
    return d


def plain_conv_desc(x_list, w_packed, B, H, W, Cout, ksize, out, **kw):
    """k x k 'same' conv over concatenated NHWC sources (list of bf16 [B,H,W,C] cuda tensors)."""
    ctot = sum(t.shape[-1] for t in x_list)
    srcs = [(t, H * W * t.shape[-1], W * t.shape[-1], t.shape[-1], H, W, t.shape[-1]) for t in x_list]
    phases = []
    r = ksize // 2
    for ky in range(ksize):
        for kx in range(ksize):
            off = 0
            for si, t in enumerate(x_list):
                phases.append((si, ky - r, kx - r, (ky * ksize + kx) * ctot + off))
                off += t.shape[-1]
    return conv_desc(srcs, phases, w_packed, ksize * ksize * ctot, B, H, W, Cout, out, **kw)


def pack_conv_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    o, i, kh, kw = w_oihw.shape
    return w_oihw.permute(0, 2, 3, 1).reshape(o, kh * kw * i).contiguous().to(device="cuda", dtype=torch.bfloat16)


def run_conv(desc, direct=False):
    lib = _lib.load()
    fn = lib.srgd_conv_direct if direct else lib.srgd_conv_igemm
    _lib.check(fn(C.byref(desc), stream()), "conv")
    torch.cuda.synchronize()


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else float(10 * torch.log10(torch.tensor(1.0 / mse)))
