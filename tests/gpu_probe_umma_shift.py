"""Probe (not a test): which descriptor base_offset makes row-shifted SWIZZLE_128B operands work?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from srgd_b200 import _lib

lib = _lib.load()
g = torch.Generator().manual_seed(0)
w = (torch.randn(128, 64, generator=g)).bfloat16().cuda()
x = (torch.randn(136, 64, generator=g)).bfloat16().cuda()
out = torch.empty(128, 128, device="cuda")
st = _lib.current_stream()
for shift in range(0, 9):
    ref = w.float() @ x.float()[shift:shift + 128].T
    res = []
    for bo in sorted({0, shift & 7, (8 - shift) & 7}):
        out.fill_(float("nan"))
        _lib.check(lib.srgd_debug_umma_shift(_lib.ptr(w), _lib.ptr(x), _lib.ptr(out), shift, bo, st))
        torch.cuda.synchronize()
        err = float((out - ref).abs().max())
        res.append(f"base_off={bo}: max err {err:.4f}")
    print(f"shift {shift}: " + " | ".join(res))
