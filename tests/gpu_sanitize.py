"""compute-sanitizer driver (not a pytest test): two p_sample steps of a dim-64 U-Net on a 128x64 tile batch of 2 with
classifier-free guidance -- every kernel family of the hot path (both conv kernels incl. the halo variant is skipped at
this width, fused linear attention, tcgen05 flash attention, GroupNorm apply with folded statistics, sampler update; round 2: split-K conv tiles, class-guidance sharing, staged LinearAttention stores, the EDM kernels, the discrete-time family's embedding + update kernels).

    compute-sanitizer --tool memcheck  python tests/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tests/gpu_sanitize.py
    compute-sanitizer --tool synccheck python tests/gpu_sanitize.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import srgd_oracle as O  # noqa: E402  (deterministic random-init weights only)
import model as M  # noqa: E402


def main():
    dim = int(os.environ.get("SAN_DIM", "128"))
    H, W = int(os.environ.get("SAN_H", "64")), int(os.environ.get("SAN_W", "256"))
    spec = O.UnetSpec(dim=dim)
    unet = M.ConditionalSRUnet(dim=dim, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=H, num_sample_steps=250)
    diff.load_state_dict(O.make_state_dict(spec, 5), strict=True)
    diff = diff.eval().to("cuda:0")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, H, W, generator=g).cuda()
    cond = (torch.rand(2, 3, H, W, generator=g) * 2 - 1).cuda()
    steps = torch.linspace(1., 0., 251)
    with torch.inference_mode():
        for i in (() if os.environ.get("SAN_ONLY_EDM") == "1" else (100, 101)):
            x, _ = diff.p_sample(x, steps[i], cond, torch.tensor([1], device="cuda"), 1.0, 3.0, steps[i + 1])
        torch.cuda.synchronize()
        print("continuous-time steps done", flush=True)
        y = x
        if os.environ.get("SAN_EDM", "1") == "1":
            # the EDM family's fused kernels (perturb / update incl. the Heun correction) around the same U-Net
            edm = M.ConditionalElucidatedDiffusionSR(unet, image_size=H, num_sample_steps=4).eval().to("cuda:0")
            edm.progress = False
            y = edm.sample(batch_size=2, condition_x=(cond + 1) * 0.5, class_label=torch.tensor([1], device="cuda"),
                           class_cond_scale=2.0, num_sample_steps=2)
    torch.cuda.synchronize()
    print("edm:", float(y.mean()), flush=True)
    if os.environ.get("SAN_GAUSS", "1") == "1":
        # the discrete-time family: fixed sinusoidal time embedding + the fused DDPM / DDIM update (small U-Net)
        spec_g = O.UnetSpec(dim=64, learned_sinusoidal_cond=False)
        unet_g = M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=False, num_classes=3)
        gd = M.ConditionalGaussianDiffusionSR(unet_g, image_size=64, timesteps=1000, sampling_timesteps=2,
                                              objective="pred_v", ddim_sampling_eta=0.5)
        gd.load_state_dict(O.make_state_dict(spec_g, 5), strict=False)
        gd = gd.eval().to("cuda:0")
        gd.progress = False
        with torch.inference_mode():
            c64 = torch.rand(2, 3, 64, 64, generator=g).cuda()
            z = gd.sample(batch_size=2, condition_x=c64, class_label=torch.tensor([1], device="cuda"), class_cond_scale=2.0)
            z2, _ = gd.p_sample(z * 2 - 1, 500, c64 * 2 - 1, torch.tensor([1], device="cuda"), 1.5, 1.0)
        torch.cuda.synchronize()
        print("gauss:", float(z.mean()), float(z2.mean()), flush=True)
    print("sanitize run finished:", float(x.abs().mean()), diff.last_step_launches, "launches per step")


if __name__ == "__main__":
    main()
