"""ncu driver (not a test): runs `p_sample` steps of the bench workload and brackets the measured
ones with cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly those launches.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tests/gpu_profile_step.py --batch 16 --steps 1
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_igemm \
        -o gpurun_out/conv python tests/gpu_profile_step.py --batch 16 --steps 1

Numbers printed by a run under ncu are never bench values (B200_PROFILING.md)."""
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
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--class_cond_scale", type=float, default=1.0)
    ap.add_argument("--first_step", type=int, default=100)
    a = ap.parse_args()
    spec = O.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=a.tile, num_sample_steps=250)
    diff.load_state_dict(O.make_state_dict(spec, 1234), strict=True)
    diff = diff.eval().to("cuda:0")
    diff.progress = False
    g = torch.Generator().manual_seed(71)
    cond = (torch.rand(a.batch, 3, a.tile, a.tile, generator=g) * 2 - 1).cuda()
    img = torch.randn(a.batch, 3, a.tile, a.tile, generator=g).cuda()
    label = torch.tensor([0], device="cuda")
    steps = torch.linspace(1., 0., 251)
    rt = torch.cuda.cudart()
    with torch.inference_mode():
        for k in range(a.warmup + a.steps):
            if k == a.warmup:
                torch.cuda.synchronize()
                rt.cudaProfilerStart()
            i = a.first_step + k
            img, _ = diff.p_sample(img, steps[i], cond, label, 1.0, a.class_cond_scale, steps[i + 1])
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
    print(f"profiled {a.steps} step(s), {diff.last_step_launches} launches per step")


if __name__ == "__main__":
    main()
