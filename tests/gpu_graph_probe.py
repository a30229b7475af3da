"""Experiment (not a test): U-Net forward replayed from a CUDA graph vs launched kernel by kernel, small tile batches.

    python tests/gpu_graph_probe.py [--batches 1,4,9]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import srgd_oracle as O  # noqa: E402  (deterministic random-init weights only)
import model as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=str, default="1,4,9")
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    spec = O.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=256, num_sample_steps=250)
    diff.load_state_dict(O.make_state_dict(spec, 1234), strict=True)
    diff = diff.eval().to("cuda:0")
    un = diff.model
    for B in [int(b) for b in a.batches.split(",")]:
        g = torch.Generator().manual_seed(3)
        x = torch.randn(B, 3, 256, 256, generator=g).cuda()
        cond = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).cuda()
        lsnr = torch.full((B,), -1.3, device="cuda")
        labels = torch.zeros(B, dtype=torch.int32, device="cuda")
        out = torch.empty(B, 3, 256, 256, device="cuda")
        with torch.inference_mode():
            for _ in range(3):
                un.run(x, lsnr, labels, cond, B, B, out=out)
            torch.cuda.synchronize()
            ref = out.clone()

            def timed(fn):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / a.iters

            t_plain = timed(lambda: un.run(x, lsnr, labels, cond, B, B, out=out))
            graph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                un.run(x, lsnr, labels, cond, B, B, out=out)
                torch.cuda.synchronize()
                with torch.cuda.graph(graph, stream=s):
                    un.run(x, lsnr, labels, cond, B, B, out=out)
            torch.cuda.current_stream().wait_stream(s)
            out.zero_()
            graph.replay()
            torch.cuda.synchronize()
            same = torch.equal(out, ref)
            t_graph = timed(graph.replay)
        print(f"B={B}: {t_plain:.3f} ms per forward launched, {t_graph:.3f} ms replayed from a CUDA graph "
              f"({un.last_launches} kernels), identical output: {same}", flush=True)


if __name__ == "__main__":
    main()
