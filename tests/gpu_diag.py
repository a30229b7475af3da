"""Diagnostic dump for GPU bring-up (not a test): per-layer errors of the CUDA U-Net vs the CPU
oracle, for both conv implementations.  Usage: python tests/gpu_diag.py > gpurun_out/diag.log"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import srgd_oracle as O  # noqa: E402
import model as M  # noqa: E402
from srgd_b200 import _lib  # noqa: E402

NAMES = ["init_conv", "t_emb", "downs.0.0", "downs.0.1", "downs.0.2", "downs.0.3", "downs.1.0", "downs.1.2",
         "downs.1.3", "downs.2.0", "downs.2.2", "downs.2.3", "downs.3.0", "downs.3.1", "downs.3.2", "downs.3.3",
         "mid_block1", "mid_attn", "mid_block2", "ups.0.0", "ups.0.1", "ups.0.2", "ups.0.3", "ups.1.0", "ups.1.2",
         "ups.1.3", "ups.2.2", "ups.2.3", "ups.3.0", "ups.3.1", "ups.3.2", "ups.3.3", "final_res_block"]


def main():
    dim = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    spec = O.UnetSpec(dim=dim)
    sd = O.make_state_dict(spec, 1234)
    unet = M.ConditionalSRUnet(dim=dim, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=size)
    diff.load_state_dict(sd)
    diff = diff.eval().to("cuda")
    g = torch.Generator().manual_seed(0)
    B = 2
    x = torch.randn(B, 3, size, size, generator=g)
    cond = torch.rand(B, 3, size, size, generator=g) * 2 - 1
    lsnr = torch.tensor([-3.7, 2.1])
    lab = torch.tensor([0, 2])
    taps = {}
    ref = O.unet_forward(sd, spec, x, lsnr, lab, cond, taps=taps)
    lib = _lib.load()
    for impl in (0, 3):
        unet.conv_impl = impl
        unet._ensure_handle(torch.device("cuda", 0))
        bufs = {}
        lib.srgd_unet_set_tap(unet._handle, None, None, 0)
        for n in NAMES:
            t = taps[n]
            if n == "t_emb":
                buf = torch.zeros(t.shape, device="cuda", dtype=torch.float32)
            else:
                b, c, h, w = t.shape
                buf = torch.zeros(b, h, w, c, device="cuda", dtype=torch.bfloat16)
            bufs[n] = buf
            _lib.check(lib.srgd_unet_set_tap(unet._handle, n.encode(), _lib.ptr(buf), buf.numel() * buf.element_size()))
        try:
            eps = unet(x.cuda(), lsnr.cuda(), lab.cuda(), cond.cuda()).cpu()
        except Exception as e:  # noqa: BLE001
            print(f"impl {impl}: forward failed: {e}")
            continue
        torch.cuda.synchronize()
        print(f"==== conv_impl={impl} dim={dim} size={size}: eps max-abs err {float((eps - ref).abs().max()):.5f} "
              f"rms err {float((eps - ref).pow(2).mean().sqrt()):.5f} (ref rms {float(ref.pow(2).mean().sqrt()):.4f})")
        for n in NAMES:
            t = taps[n]
            got = bufs[n].float().cpu()
            if n != "t_emb":
                got = got.permute(0, 3, 1, 2)
            err = (got - t).abs()
            print(f"  {n:18s} shape {tuple(t.shape)!s:22s} max-abs {float(err.max()):9.5f} rms {float(err.pow(2).mean().sqrt()):9.5f} "
                  f"ref rms {float(t.pow(2).mean().sqrt()):8.4f} nan {int(torch.isnan(got).sum())}")
        lib.srgd_unet_set_tap(unet._handle, None, None, 0)
    unet.conv_impl = 0


if __name__ == "__main__":
    main()
