"""Experiment (not a test): does running two half-batches on two CUDA streams overlap the HBM-bound kernels
(GroupNorm apply, linear attention) of one with the tensor-bound convs of the other?

    python tests/gpu_two_stream.py [--batch 16] [--steps 20]
"""
import argparse
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import srgd_oracle as O  # noqa: E402  (deterministic random-init weights only)
import model as M  # noqa: E402


def make(dev):
    spec = O.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=256, num_sample_steps=250)
    diff.load_state_dict(O.make_state_dict(spec, 1234), strict=True)
    diff = diff.eval().to(dev)
    diff.progress = False
    return diff


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--streams", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    ns = a.streams
    diffs = [make(dev) for _ in range(ns)]
    streams = [torch.cuda.Stream() for _ in range(ns)]
    steps = torch.linspace(1., 0., 251)
    label = torch.tensor([0], device=dev)
    B = a.batch
    g = torch.Generator().manual_seed(71)
    cond = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).to(dev)
    img0 = torch.randn(B, 3, 256, 256, generator=g).to(dev)
    noise = torch.randn(B, 3, 256, 256, generator=g).to(dev)
    hb = B // ns

    def one_stream(n):
        img = img0
        for k in range(n):
            img, _ = diffs[0].p_sample(img, steps[100 + k], cond, label, 1.0, 1.0, steps[101 + k], noise=noise)
        return img

    def multi_stream(n):
        parts = [img0[i * hb:(i + 1) * hb] for i in range(ns)]
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for k in range(n):
            for i, s in enumerate(streams):
                with torch.cuda.stream(s):
                    parts[i], _ = diffs[i].p_sample(parts[i], steps[100 + k], cond[i * hb:(i + 1) * hb], label, 1.0, 1.0,
                                                    steps[101 + k], noise=noise[i * hb:(i + 1) * hb])
        for s in streams:
            cur.wait_stream(s)
        return torch.cat(parts)

    with torch.inference_mode():
        for fn, name in ((one_stream, f"1 stream  x B={B}"), (multi_stream, f"{ns} streams x B={hb}")):
            fn(3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(a.steps)
            e1.record()
            torch.cuda.synchronize()
            print(f"{name}: {e0.elapsed_time(e1) / a.steps:.3f} ms/step, checksum {float(out.double().sum()):.6f}", flush=True)
        a1, a2 = one_stream(2), multi_stream(2)
        torch.cuda.synchronize()
        print("outputs equal:", torch.equal(a1, a2), float((a1 - a2).abs().max()))


if __name__ == "__main__":
    main()
