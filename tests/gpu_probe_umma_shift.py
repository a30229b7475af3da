"""Probe (not a test): which descriptor base_offset makes row-shifted SWIZZLE_128B operands work?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import subprocess
from srgd_b200 import _lib

lib = _lib.load()
# the probe kernel lives outside the product library: tests/csrc/debug_umma.cu, linked against libsrgd_b200.so for
# the tensor-map helper
PROBE = os.path.join(ROOT, "tests", "csrc", "libsrgd_probe.so")
subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared",
                "-Xcompiler", "-fPIC", os.path.join(ROOT, "tests", "csrc", "debug_umma.cu"), "-o", PROBE,
                "-L" + os.path.dirname(_lib.LIB_PATH), "-l:libsrgd_b200.so",
                "-Xlinker", "-rpath=" + os.path.dirname(_lib.LIB_PATH)], check=True)
probe = C.CDLL(PROBE)
probe.srgd_debug_umma_shift.restype = C.c_int
probe.srgd_debug_umma_shift.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
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
        _lib.check(probe.srgd_debug_umma_shift(_lib.ptr(w), _lib.ptr(x), _lib.ptr(out), shift, bo, st))
        torch.cuda.synchronize()
        err = float((out - ref).abs().max())
        res.append(f"base_off={bo}: max err {err:.4f}")
    print(f"shift {shift}: " + " | ".join(res))
