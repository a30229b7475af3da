"""ctypes binding of libsrgd_b200.so (C-ABI: include/srgd_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not an
sm_100 GPU, every entry point raises.  PyTorch is only used by callers for device memory and
streams; all arithmetic on the hot path happens inside the library's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SRGD_B200_LIB: developer A/B knob (an alternative build of the same library, srgd_b200/build.py `variant`)
LIB_PATH = os.environ.get("SRGD_B200_LIB") or os.path.join(HERE, "libsrgd_b200.so")

SRGD_CONV_MAX_SRC = 4
SRGD_CONV_MAX_PHASE = 20
OUT_BF16_NHWC = 0
OUT_PIXEL_SHUFFLE = 1


class SrgdError(RuntimeError):
    pass


class StepScalars(C.Structure):
    _fields_ = [("alpha", C.c_float), ("sigma", C.c_float), ("alpha_next", C.c_float), ("c", C.c_float),
                ("noise_scale", C.c_float), ("guidance_scale", C.c_float), ("clip", C.c_int32)]


class EdmScalars(C.Structure):
    _fields_ = [("c_skip", C.c_float), ("c_out", C.c_float), ("guidance_scale", C.c_float), ("sigma_eval", C.c_float),
                ("step", C.c_float), ("c_in_next", C.c_float), ("clip", C.c_int32)]


class GaussScalars(C.Structure):
    _fields_ = [("objective", C.c_int32), ("mode", C.c_int32), ("clip", C.c_int32), ("rederive", C.c_int32),
                ("guidance_scale", C.c_float), ("sqrt_recip_ac", C.c_float), ("sqrt_recipm1_ac", C.c_float),
                ("sqrt_ac", C.c_float), ("sqrt_1m_ac", C.c_float), ("coef1", C.c_float), ("coef2", C.c_float),
                ("noise_scale", C.c_float), ("sqrt_ac_next", C.c_float), ("c", C.c_float)]


OBJ_PRED_NOISE, OBJ_PRED_X0, OBJ_PRED_V = 0, 1, 2
GAUSS_DDPM, GAUSS_DDIM, GAUSS_DDIM_LAST = 0, 1, 2

SRGD_MAX_TILES_PER_CALL = 64


class TileCoords(C.Structure):
    _fields_ = [("n", C.c_int32), ("yx", (C.c_int32 * 2) * SRGD_MAX_TILES_PER_CALL)]


class ConvSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sb", C.c_int64), ("sy", C.c_int64), ("sx", C.c_int64),
                ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32)]


class ConvPhase(C.Structure):
    _fields_ = [("src", C.c_int32), ("dy", C.c_int32), ("dx", C.c_int32), ("k_start", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("B", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32), ("Cout", C.c_int32),
                ("n_src", C.c_int32), ("n_phase", C.c_int32),
                ("srcs", ConvSrc * SRGD_CONV_MAX_SRC), ("phases", ConvPhase * SRGD_CONV_MAX_PHASE),
                ("weight", C.c_void_p), ("Ktot", C.c_int64), ("bias", C.c_void_p), ("row_scale", C.c_void_p),
                ("residual", C.c_void_p), ("act", C.c_int32), ("out_mode", C.c_int32), ("out", C.c_void_p),
                ("gn_partials", C.c_void_p), ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_int64)]


class UnetConfig(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_stages", C.c_int32), ("dim_mults", C.c_int32 * 6),
                ("full_attn", C.c_int32 * 6), ("heads", C.c_int32), ("dim_head", C.c_int32),
                ("groups", C.c_int32), ("channels", C.c_int32), ("sinu_dim", C.c_int32),
                ("num_classes", C.c_int32), ("fixed_sinusoidal", C.c_int32)]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_SZ = C.c_size_t

# name -> (restype, argtypes); every symbol declared in include/srgd_b200.h
SIGNATURES = {
    "srgd_version": (C.c_int, []),
    "srgd_last_error": (C.c_char_p, []),
    "srgd_device_check": (C.c_int, [C.c_int]),
    "srgd_sampler_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _I64, C.POINTER(StepScalars), _P]),
    "srgd_edm_perturb": (C.c_int, [_P, _P, C.c_float, C.c_float, C.c_float, _P, _P, _I64, _P]),
    "srgd_edm_update": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, C.POINTER(EdmScalars), _P]),
    "srgd_edm_dpmpp": (C.c_int, [_P, _P, _P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P, _I64, _P]),
    "srgd_gauss_update": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, C.POINTER(GaussScalars), _P]),
    "srgd_q_sample": (C.c_int, [_P, _P, _P, _I64, C.c_float, C.c_float, _P]),
    "srgd_finalize_image": (C.c_int, [_P, _P, _I64, _P]),
    "srgd_gather_tiles": (C.c_int, [_P, _P, C.POINTER(TileCoords), _I32, _I32, _I32, _I32, _P]),
    "srgd_scatter_tiles": (C.c_int, [_P, _P, C.POINTER(TileCoords), _I32, _I32, _I32, _I32, _P]),
    "srgd_renoise_outside": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, C.c_float, _P]),
    "srgd_conv_m_tiles": (C.c_int, [_I32, _I32, _I32]),
    "srgd_conv_splitk_workspace_bytes": (_SZ, []),
    "srgd_conv_igemm": (C.c_int, [C.POINTER(ConvDesc), _P]),
    "srgd_conv_direct": (C.c_int, [C.POINTER(ConvDesc), _P]),
    "srgd_groupnorm_finalize": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P]),
    "srgd_groupnorm_stats": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P]),
    "srgd_groupnorm_apply": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "srgd_groupnorm_apply_ex": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _I64, _P, _I32, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "srgd_groupnorm_apply_final": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P]),
    "srgd_pixel_inv_norm": (C.c_int, [_P, _P, _I64, _I32, _P]),
    "srgd_rmsnorm_residual": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P]),
    "srgd_linear_attention_workspace": (_SZ, [_I32, _I32, _I32]),
    "srgd_linear_attention": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _SZ, _P]),
    "srgd_linear_attention_block_supported": (C.c_int, [_I32, _I32, _I32]),
    "srgd_linear_attention_block_workspace": (_SZ, [_I32, _I32, _I32, _I32]),
    "srgd_linear_attention_block": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _SZ, _P]),
    "srgd_attention": (C.c_int, [_P, _P, _I32, _I32, _I32, _P]),
    "srgd_attention_tc_supported": (C.c_int, [_I32, _I32]),
    "srgd_attention_tc": (C.c_int, [_P, _P, _I32, _I32, _I32, _P]),
    "srgd_pack_input": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, _I32, _I32, _P]),
    "srgd_final_conv": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "srgd_dense_rows": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _P]),
    "srgd_fourier_features": (C.c_int, [_P, _P, _P, _I32, _I32, _P]),
    "srgd_sinusoidal_pos_emb": (C.c_int, [_P, _P, _P, _I32, _I32, _P]),
    "srgd_add_class_rows": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P]),
    "srgd_unet_param_count": (C.c_int, [C.POINTER(UnetConfig)]),
    "srgd_unet_param_name": (C.c_char_p, [C.POINTER(UnetConfig), C.c_int]),
    "srgd_unet_create": (C.c_int, [C.POINTER(UnetConfig), C.POINTER(_P), C.c_int, C.POINTER(_P)]),
    "srgd_unet_destroy": (None, [_P]),
    "srgd_unet_workspace_bytes": (_SZ, [_P, _I32, _I32, _I32]),
    "srgd_unet_forward": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _P, _I32, _I32, _I32, _P, _SZ, _I32, _P]),
    "srgd_unet_set_tap": (C.c_int, [_P, C.c_char_p, _P, _SZ]),
    "srgd_unet_last_launch_count": (C.c_int, [_P]),
    "srgd_launch_count": (C.c_longlong, []),
    "srgd_set_batch_invariant": (C.c_int, [C.c_int]),
    "srgd_profile_begin": (C.c_int, []),
    "srgd_profile_end": (C.c_int, []),
    "srgd_profile_get": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                   C.POINTER(C.c_int)]),
    "srgd_profile_record_count": (C.c_int, []),
    "srgd_profile_record": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double)]),
}

PROFILE_KINDS = ["conv_igemm", "gn_apply", "sampler_step", "linear_attention", "full_attention", "norm_misc", "other"]


def profile_report():
    """{kind: dict(ms, flops, bytes, launches)} accumulated since srgd_profile_begin (after _end)."""
    lib = load()
    out = {}
    for k, name in enumerate(PROFILE_KINDS):
        ms, fl, by, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        check(lib.srgd_profile_get(k, C.byref(ms), C.byref(fl), C.byref(by), C.byref(n)), "srgd_profile_get")
        out[name] = dict(ms=ms.value, flops=fl.value, bytes=by.value, launches=n.value)
    return out

def profile_records():
    """[(kind name, ms, flops, bytes)] per API-level launch, in launch order (after srgd_profile_end)."""
    lib = load()
    out = []
    for i in range(lib.srgd_profile_record_count()):
        k, ms, fl, by = C.c_int(), C.c_double(), C.c_double(), C.c_double()
        check(lib.srgd_profile_record(i, C.byref(k), C.byref(ms), C.byref(fl), C.byref(by)), "srgd_profile_record")
        out.append((PROFILE_KINDS[k.value], ms.value, fl.value, by.value))
    return out


_lib = None


def load():
    """Load the library (once).  Raises SrgdError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SrgdError(
            f"{LIB_PATH} is missing: build it with `python -m srgd_b200.build` (nvcc, sm_100a). "
            "srgd_b200 has no CPU or PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().srgd_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SrgdError(f"{what or 'srgd_b200'} failed ({rc}): {last_error()}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_like(t, ref, what: str):
    """A caller-supplied tensor that a kernel reads element for element next to `ref` (torch would raise a broadcast
    / size error in the reference's elementwise op; a raw pointer would be read out of bounds)."""
    if tuple(t.shape) != tuple(ref.shape) or t.device != ref.device:
        raise RuntimeError(f"{what}: expected a tensor of shape {tuple(ref.shape)} on {ref.device}, "
                           f"got {tuple(t.shape)} on {t.device}")


def require_cuda(t, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: tensor is on {t.device}; srgd_b200 runs on sm_100 CUDA devices only "
                           "(there is no CPU fallback)")
