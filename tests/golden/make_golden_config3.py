"""Classifier-free guidance end to end on the UNMODIFIED reference (CPU, fp32): the batched `sample()` entry of
BASELINE.json configs[2] at batch 1 -- shipped conf (dim 128, 250 steps), one synthetic 64x64 LR image (PIL bicubic
x4), test_label 2, class_cond_scale 3.0 (two U-Net calls per step, model.py:3151-3154), seed 71.  About 25 minutes on
8 cores.

    python tests/golden/make_golden_config3.py

Weights: oracle.make_state_dict(UnetSpec(), 1234, init="torch").  The product replays the reference's noise stream with
`rng_device = "cpu"` (tests/test_gpu_unet.py::test_config3_cfg_vs_reference_golden).
"""
import os, sys, time, warnings, random
import numpy as np
import torch
from PIL import Image

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")                 # the reference's model.py must win over the repo-root drop-in
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.append(ROOT)
import model as ref
assert os.path.realpath(ref.__file__).startswith("/root/reference/"), ref.__file__
from oracle import srgd_oracle as O

torch.set_num_threads(os.cpu_count())


@torch.inference_mode()
def main():
    steps = int(os.environ.get("GOLDEN_STEPS", "250"))
    spec = O.UnetSpec()
    unet = ref.ConditionalSRUnet(dim=spec.dim, dim_mults=spec.dim_mults, full_attn=spec.full_attn,
                                 learned_variance=False, learned_sinusoidal_cond=True,
                                 learned_sinusoidal_dim=spec.learned_sinusoidal_dim, flash_attn=False,
                                 pixel_shuffle_upsample=True, num_classes=spec.num_classes)
    diff = ref.ConditionalContinuousTimeGaussianDiffusionSR(
        model=unet, image_size=256, noise_schedule="linear", num_sample_steps=steps,
        clip_sample_denoised=True).eval()
    diff.load_state_dict(O.make_state_dict(spec, 1234, init="torch"), strict=True)
    lr = np.random.RandomState(73).randint(0, 256, (64, 64, 3), dtype=np.uint8)
    hr = Image.fromarray(lr, mode="RGB").resize((256, 256), resample=Image.BICUBIC)
    cond01 = torch.from_numpy(np.array(hr, dtype=np.uint8)).permute(2, 0, 1).float().div(255.)[None]
    label = torch.LongTensor([2])
    random.seed(71); np.random.seed(71); torch.manual_seed(71)
    t0 = time.time()
    out = diff.sample(batch_size=1, condition_x=cond01, class_label=label, cond_scale=1.0, guidance_start_steps=0,
                      class_cond_scale=3.0, class_guidance_start_steps=0, generation_start_steps=0,
                      num_sample_steps=steps)
    print(f"reference sample() with CFG 3.0: {time.time() - t0:.1f} s for {steps} steps on {os.cpu_count()} threads")
    name = "config3_full" if steps == 250 else f"config3_{steps}"
    np.savez_compressed(os.path.join(HERE, name + ".npz"), lr=lr, img=out.numpy().astype(np.float32), steps=steps,
                        seed=71, label=2, ccs=3.0)
    print(name, os.path.getsize(os.path.join(HERE, name + ".npz")) / 1024, "KiB")


if __name__ == "__main__":
    main()
