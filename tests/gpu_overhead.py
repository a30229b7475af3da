"""Host-overhead probe (not a pytest test): per tile batch, the CPU time to ISSUE one p_sample step (GPU idle at the
start, no sync inside) next to the device time of the same step, plus a run-to-run bit-exactness check of the step
(a race between programmatically overlapped launches would show up here).

    python tests/gpu_overhead.py [--batches 1,4,16] [--steps 20]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import srgd_oracle as O  # noqa: E402  (deterministic random-init weights only)
import model as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=str, default="1,4,16")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--class_cond_scale", type=float, default=1.0)
    a = ap.parse_args()
    spec = O.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=256, num_sample_steps=250)
    diff.load_state_dict(O.make_state_dict(spec, 1234), strict=True)
    diff = diff.eval().to("cuda:0")
    diff.progress = False
    steps = torch.linspace(1., 0., 251)
    label = torch.tensor([0], device="cuda")
    for B in [int(b) for b in a.batches.split(",")]:
        g = torch.Generator().manual_seed(71)
        cond = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).cuda()
        img0 = torch.randn(B, 3, 256, 256, generator=g).cuda()
        noise = torch.randn(B, 3, 256, 256, generator=g).cuda()
        with torch.inference_mode():
            img = img0
            for k in range(3):
                img, _ = diff.p_sample(img, steps[100 + k], cond, label, 1.0, a.class_cond_scale, steps[101 + k])
            torch.cuda.synchronize()
            cpu_ms = []
            for k in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                img, _ = diff.p_sample(img, steps[110 + k], cond, label, 1.0, a.class_cond_scale, steps[111 + k])
                cpu_ms.append((time.perf_counter() - t0) * 1e3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(a.steps):
                img, _ = diff.p_sample(img, steps[120 + k], cond, label, 1.0, a.class_cond_scale, steps[121 + k])
            e1.record()
            torch.cuda.synchronize()
            gpu_ms = e0.elapsed_time(e1) / a.steps
            # bit-exactness of one teacher-forced step across repeats
            outs = []
            for rep in range(4):
                o, _ = diff.p_sample(img0, steps[100], cond, label, 1.0, a.class_cond_scale, steps[101], noise=noise)
                outs.append(o.clone())
            torch.cuda.synchronize()
            same = all(torch.equal(outs[0], o) for o in outs[1:])
        print(f"B={B:3d}: issue {sorted(cpu_ms)[len(cpu_ms) // 2]:.3f} ms/step on the CPU, {gpu_ms:.3f} ms/step on the "
              f"device ({diff.last_step_launches} launches), repeats bit-identical: {same}", flush=True)


if __name__ == "__main__":
    main()
