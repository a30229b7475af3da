"""Generate tests/golden/edm_tiny.npz from the UNMODIFIED reference's ConditionalElucidatedDiffusionSR
(/root/reference/model.py:2059-2560) behind oracle/_shim (whose `ElucidatedDiffusion` base restates the pip package's
coefficients and schedule: parity unpinned for those).  Run in the build container only:

    python tests/golden/make_golden_edm.py

Cases (dim-16 U-Net, seeded CPU generator re-seeded per case like inference.py:47-51): Heun `sample_org` with class
guidance 2.0 (two U-Net calls per evaluation there), Heun with LR-condition guidance and generation_start_steps,
DPM-Solver++ 2M `sample_using_dpmpp`, Heun `tiled_sample` (104x120 image -> 256x256 canvas of 32-pixel tiles), and one
`preconditioned_network_forward` evaluation per guidance kind."""
import os, sys, warnings
import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.environ.get("GOLDEN_OUT", HERE)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.append(ROOT)
import model as ref
assert os.path.realpath(ref.__file__).startswith("/root/reference/"), ref.__file__
import tqdm as _tqdm
ref.tqdm = lambda it, **kw: it                       # silence the progress bars
from oracle import srgd_oracle as O

torch.set_num_threads(os.cpu_count())
SPEC, SEED, STEPS = O.UnetSpec(dim=16), 11, 6


def build(use_dpmpp):
    unet = ref.ConditionalSRUnet(dim=SPEC.dim, dim_mults=SPEC.dim_mults, full_attn=SPEC.full_attn, learned_variance=False,
                                 learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, flash_attn=False,
                                 pixel_shuffle_upsample=True, num_classes=3)
    edm = ref.ConditionalElucidatedDiffusionSR(unet, image_size=32, num_sample_steps=STEPS, use_dpmpp_solver=use_dpmpp).eval()
    sd = O.make_state_dict(SPEC, SEED, prefix="net.")
    assert list(edm.state_dict().keys()) == list(sd.keys())
    edm.load_state_dict(sd, strict=True)
    return edm


@torch.inference_mode()
def main():
    g = torch.Generator().manual_seed(3)
    cond = torch.rand(2, 3, 32, 32, generator=g)
    big = torch.rand(1, 3, 104, 120, generator=g)
    x = torch.randn(2, 3, 32, 32, generator=g)
    label = torch.tensor([1])
    out = dict(cond=cond, big=big, x=x, label=label, steps=STEPS)
    edm = build(False)
    for name, sig, cs, ccs in (("fwd_plain", 1.7, 1.0, 1.0), ("fwd_class", 0.3, 1.0, 2.0), ("fwd_cond", 12.0, 1.5, 1.0)):
        out[name] = edm.preconditioned_network_forward(x, sig, cond * 2 - 1, label, cs, ccs, clamp=True)
        out[name + "_meta"] = np.array([sig, cs, ccs])
    torch.manual_seed(71)
    out["heun_class"] = edm.sample(batch_size=2, condition_x=cond, class_label=label, class_cond_scale=2.0,
                                   num_sample_steps=STEPS)
    torch.manual_seed(71)
    # (generation_start_steps > 0 only works for one image in the reference: get_noised_images broadcasts a [b] sigma
    # against [b,3,h,w], model.py:2192-2194)
    out["heun_cond_start2"] = edm.sample(batch_size=1, condition_x=cond[:1], class_label=label, cond_scale=1.5,
                                         guidance_start_steps=3, generation_start_steps=2, num_sample_steps=STEPS)
    torch.manual_seed(71)
    out["tiled_heun"] = edm.tiled_sample(batch_size=5, tile_size=32, tile_stride=32, condition_x=big, class_label=label,
                                         class_cond_scale=2.0, num_sample_steps=STEPS)
    edm2 = build(True)
    torch.manual_seed(71)
    out["dpmpp_class"] = edm2.sample(batch_size=2, condition_x=cond, class_label=label, class_cond_scale=2.0,
                                     num_sample_steps=STEPS)
    path = os.path.join(OUT, "edm_tiny.npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print(f"edm_tiny: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
