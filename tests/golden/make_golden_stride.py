"""Golden for `tiled_sample` with OVERLAPPING tiles, from the UNMODIFIED reference (CPU, tiny U-Net, seconds):
tile_stride < tile_size (the shifted grid's tiles overlap and the canvas advances in place, minibatch by minibatch,
model.py:3374-3385) and a tile size / stride that does not divide the 256-aligned canvas (the last tile of an axis is
pulled back to the border, model.py:137-150).

    python tests/golden/make_golden_stride.py
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
ref.tqdm = lambda it, **kw: it
from oracle import srgd_oracle as O

torch.set_num_threads(os.cpu_count())
CASES = {"t32_s16": (32, 16), "t64_s48": (64, 48), "t48_s48": (48, 48)}


@torch.inference_mode()
def main():
    spec = O.UnetSpec(dim=16)
    unet = ref.ConditionalSRUnet(dim=spec.dim, dim_mults=spec.dim_mults, full_attn=spec.full_attn, learned_variance=False,
                                 learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, flash_attn=False,
                                 pixel_shuffle_upsample=True, num_classes=3)
    diff = ref.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=32, noise_schedule="linear",
                                                            num_sample_steps=3, clip_sample_denoised=True).eval()
    diff.load_state_dict(O.make_state_dict(spec, 11), strict=True)
    cond01 = torch.rand(1, 3, 104, 120, generator=torch.Generator().manual_seed(7))
    out = {}
    for name, (tile, stride) in CASES.items():
        torch.manual_seed(71)
        out[name] = diff.tiled_sample(batch_size=5, tile_size=tile, tile_stride=stride, condition_x=cond01,
                                      class_label=torch.tensor([1]), class_cond_scale=2.0, num_sample_steps=3)
    path = os.path.join(OUT, "tiled_stride_tiny.npz")
    np.savez_compressed(path, cond01_checksum=float(cond01.double().sum()),
                        **{k: v.numpy().astype(np.float32) for k, v in out.items()})
    print("tiled_stride_tiny:", os.path.getsize(path) / 1024, "KiB")


if __name__ == "__main__":
    main()
