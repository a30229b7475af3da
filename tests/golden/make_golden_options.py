"""Goldens for the step-gating options of the sampling loops, from the UNMODIFIED reference (CPU, tiny U-Net, seconds):
`generation_start_steps` (start from q_sample(condition) and skip the first steps, model.py:3196-3203, 3218-3219),
`class_guidance_start_steps` / `guidance_start_steps` (scale forced to 1 before that step, model.py:3221-3228), for both
`sample()` and `tiled_sample()`.

    python tests/golden/make_golden_options.py
"""
import os, sys, warnings
import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.environ.get("GOLDEN_OUT", HERE)
sys.path.insert(0, "/root/reference")                 # the reference's model.py must win over the repo-root drop-in
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.append(ROOT)
import model as ref
assert os.path.realpath(ref.__file__).startswith("/root/reference/"), ref.__file__
from oracle import srgd_oracle as O

torch.set_num_threads(os.cpu_count())


def build(spec, seed, image_size, steps):
    unet = ref.ConditionalSRUnet(dim=spec.dim, dim_mults=spec.dim_mults, full_attn=spec.full_attn,
                                 learned_variance=False, learned_sinusoidal_cond=True,
                                 learned_sinusoidal_dim=spec.learned_sinusoidal_dim, flash_attn=False,
                                 pixel_shuffle_upsample=True, num_classes=spec.num_classes)
    diff = ref.ConditionalContinuousTimeGaussianDiffusionSR(
        model=unet, image_size=image_size, noise_schedule="linear", num_sample_steps=steps,
        clip_sample_denoised=True).eval()
    diff.load_state_dict(O.make_state_dict(spec, seed), strict=True)
    return diff


@torch.inference_mode()
def main():
    spec = O.UnetSpec(dim=16)
    out = {}
    g = torch.Generator().manual_seed(21)
    cond01 = torch.rand(2, 3, 64, 64, generator=g)
    diff = build(spec, 11, 64, 8)
    torch.manual_seed(71)
    out["sample_a"] = diff.sample(batch_size=2, condition_x=cond01, class_label=torch.tensor([1]), class_cond_scale=2.5,
                                  class_guidance_start_steps=3, generation_start_steps=2, num_sample_steps=8)
    torch.manual_seed(71)
    out["sample_b"] = diff.sample(batch_size=2, condition_x=cond01, class_label=torch.tensor([0, 2]), cond_scale=1.7,
                                  guidance_start_steps=5, num_sample_steps=8)
    diff = build(spec, 11, 256, 4)
    cond_t = torch.rand(1, 3, 272, 264, generator=g)
    torch.manual_seed(71)
    out["tiled_a"] = diff.tiled_sample(batch_size=4, condition_x=cond_t, class_label=torch.tensor([2]),
                                       class_cond_scale=2.0, class_guidance_start_steps=2, generation_start_steps=1,
                                       num_sample_steps=4)
    path = os.path.join(OUT, "options_tiny.npz")
    # inputs are re-drawn by the test from torch.Generator().manual_seed(21) (cond01 [2,3,64,64], then cond_t
    # [1,3,272,264]); the tiled output keeps a stride-2 sub-sample to stay small
    out["tiled_a_sub2"] = out.pop("tiled_a")[..., ::2, ::2].contiguous()
    np.savez_compressed(path, cond01_checksum=float(cond01.double().sum()), cond_t_checksum=float(cond_t.double().sum()),
                        **{k: v.numpy().astype(np.float32) for k, v in out.items()})
    print("options_tiny:", os.path.getsize(path) / 1024, "KiB")


if __name__ == "__main__":
    main()
