"""Generate tests/golden/*.npz from the UNMODIFIED reference (/root/reference/model.py).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py          # GOLDEN_OUT=/tmp/g to write elsewhere

Weights come from oracle.srgd_oracle.make_state_dict(spec, seed) and are loaded into the
reference modules with load_state_dict(strict=True) -- which also pins the 280-key checkpoint
layout.  Inputs are seeded torch CPU tensors; every fixture stores its inputs so the tests do
not depend on RNG reproducibility across machines (except the free-running sample()/
tiled_sample() cases, which re-seed torch's global CPU generator exactly like inference.py:47-51).
"""
import os, sys, logging, warnings
import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.environ.get("GOLDEN_OUT", HERE)              # write somewhere else to check the committed fixtures reproduce
sys.path.insert(0, "/root/reference")                 # the reference's model.py must win over the repo-root drop-in
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.append(ROOT)
import model as ref
assert os.path.realpath(ref.__file__).startswith("/root/reference/"), ref.__file__
from oracle import srgd_oracle as O

torch.set_num_threads(os.cpu_count())
log = logging.getLogger("golden")

SPECS = {
    "tiny": O.UnetSpec(dim=16),
    "mid": O.UnetSpec(dim=64),
    "full": O.UnetSpec(dim=128),
}
SEEDS = {"tiny": 11, "mid": 22, "full": 1234}


def build_ref(spec, seed, image_size, steps=250):
    unet = ref.ConditionalSRUnet(dim=spec.dim, dim_mults=spec.dim_mults, full_attn=spec.full_attn,
                                 learned_variance=False, learned_sinusoidal_cond=True,
                                 learned_sinusoidal_dim=spec.learned_sinusoidal_dim, flash_attn=False,
                                 pixel_shuffle_upsample=True, num_classes=spec.num_classes)
    diff = ref.ConditionalContinuousTimeGaussianDiffusionSR(
        model=unet, image_size=image_size, noise_schedule="linear", num_sample_steps=steps,
        clip_sample_denoised=True).eval()
    sd = O.make_state_dict(spec, seed)
    assert list(diff.state_dict().keys()) == list(sd.keys()), "checkpoint key order differs"
    diff.load_state_dict(sd, strict=True)
    return diff


def save(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


@torch.inference_mode()
def main():
    # ---- schedule scalars for the shipped 250-step schedule (model.py:3127-3134, 3325) ----
    steps = torch.linspace(1., 0., 251)
    rows = []
    for i in range(250):
        t, tn = steps[i], steps[i + 1]
        l, ln = ref.beta_linear_log_snr(t), ref.beta_linear_log_snr(tn)
        c = -torch.special.expm1(l - ln)
        rows.append([l, ln, c, l.sigmoid().sqrt(), (-l).sigmoid().sqrt(), ln.sigmoid().sqrt(),
                     (-ln).sigmoid() * c])
    save("scalars_250", steps=steps, table=torch.tensor(rows, dtype=torch.float32))

    # ---- tile geometry (model.py:116-179) ----
    geo = []
    for (h, w) in [(256, 256), (100, 180), (272, 264), (512, 512), (2048, 2048), (300, 700), (257, 256)]:
        coord, pad = ref.get_coord_and_pad(h, w)
        H, W = h + pad[2] + pad[3], w + pad[0] + pad[1]
        c0 = ref.get_coords(H, W, 256, 256, 0)
        c1 = ref.get_coords(H, W, 256, 256, 0) if (H <= 256 and W <= 256) else \
            ref.get_coords(H - 256, W - 256, 256, 256, 128)
        area, apad = ref.get_area(c1, H, W)
        geo.append(dict(hw=(h, w), coord=coord, pad=pad, n0=len(c0), n1=len(c1), c0=c0, c1=c1,
                        area=area, apad=apad))
    import json
    with open(os.path.join(OUT, "geometry.json"), "w") as f:
        json.dump(geo, f)

    # ---- U-Net forward goldens ----
    for tag, size, B in (("tiny", 64, 2), ("mid", 64, 2), ("full", 64, 1)):
        spec = SPECS[tag]
        diff = build_ref(spec, SEEDS[tag], size)
        g = torch.Generator().manual_seed(100 + B)
        x = torch.randn(B, 3, size, size, generator=g)
        cond = torch.rand(B, 3, size, size, generator=g) * 2 - 1
        lsnr = torch.tensor([-3.7, 2.1][:B])
        labels = torch.tensor([0, 2][:B])
        out = {}
        out["eps_label_cond"] = diff.model(x, lsnr, labels, cond)
        out["eps_nolabel_cond"] = diff.model(x, lsnr, None, cond)
        out["eps_label_nocond"] = diff.model(x, lsnr, labels, None)
        out["eps_label1_cond"] = diff.model(x, lsnr, labels[:1], cond)      # [1]-label broadcast
        # a few intermediate activations via hooks (layer-level pinning)
        acts = {}
        hooks = []
        names = {"init_conv": diff.model.init_conv, "downs.0.0": diff.model.downs[0][0],
                 "downs.0.2": diff.model.downs[0][2], "downs.3.2": diff.model.downs[3][2],
                 "mid_block1": diff.model.mid_block1, "ups.0.3": diff.model.ups[0][3],
                 "final_res_block": diff.model.final_res_block}
        for n, m in names.items():
            hooks.append(m.register_forward_hook(lambda mod, i, o, n=n: acts.__setitem__(n, o.clone())))
        diff.model(x, lsnr, labels, cond)
        for h in hooks:
            h.remove()
        # keep fixtures small: store a strided sub-sample of big activations
        for n, a in acts.items():
            out["act_" + n] = a[:, ::max(1, a.shape[1] // 8), ::4, ::4].contiguous()
        save(f"unet_{tag}", x=x, cond=cond, log_snr=lsnr, labels=labels, **out)

    # ---- teacher-forced p_sample (model.py:3174-3188) ----
    spec = SPECS["mid"]
    diff = build_ref(spec, SEEDS["mid"], 64)
    g = torch.Generator().manual_seed(7)
    B = 2
    cond = torch.rand(B, 3, 64, 64, generator=g) * 2 - 1
    label = torch.tensor([1])
    cases = {}
    for ci, (i, cs, ccs) in enumerate([(0, 1.0, 1.0), (0, 1.0, 3.0), (37, 1.0, 3.0), (124, 2.0, 1.0),
                                       (200, 1.0, 1.0), (249, 1.0, 3.0), (249, 1.0, 1.0)]):
        t, tn = steps[i], steps[i + 1]
        a, s = ref.beta_linear_log_snr(t).sigmoid().sqrt(), (-ref.beta_linear_log_snr(t)).sigmoid().sqrt()
        x0 = torch.rand(B, 3, 64, 64, generator=g) * 2 - 1
        x = a * x0 + s * torch.randn(B, 3, 64, 64, generator=g)
        noise = torch.randn(B, 3, 64, 64, generator=g)
        # teacher-force the randn_like draw: re-seed so we can reproduce it -> instead draw here
        torch.manual_seed(1000 + ci)
        img, xs = diff.p_sample(x, t, cond, label, cs, ccs, tn)
        torch.manual_seed(1000 + ci)
        drawn = torch.randn_like(x) if tn != 0 else torch.zeros_like(x)
        mean, var, xs2 = diff.p_mean_variance(x, t, cond, label, cs, ccs, tn)
        cases[f"c{ci}_x"] = x; cases[f"c{ci}_noise"] = drawn
        cases[f"c{ci}_meta"] = np.array([i, cs, ccs], dtype=np.float64)
        cases[f"c{ci}_img"] = img; cases[f"c{ci}_x0"] = xs; cases[f"c{ci}_mean"] = mean
        cases[f"c{ci}_var"] = var
    save("p_sample_mid", cond=cond, label=label, ncases=7, **cases)

    # ---- free-running sample() (model.py:3417-3430), global-RNG seeded like inference.py ----
    for tag, ccs, nsteps in (("tiny", 1.0, 8), ("mid", 3.0, 6)):
        spec = SPECS[tag]
        diff = build_ref(spec, SEEDS[tag], 64, steps=nsteps)
        g = torch.Generator().manual_seed(5)
        cond01 = torch.rand(2, 3, 64, 64, generator=g)
        torch.manual_seed(71)
        img = diff.sample(batch_size=2, condition_x=cond01, class_label=torch.tensor([2]),
                          class_cond_scale=ccs, num_sample_steps=nsteps)
        save(f"sample_{tag}", cond01=cond01, label=np.array([2]), ccs=ccs, nsteps=nsteps, seed=71, img=img)

    # ---- tiled_sample() (model.py:3288-3413): 272x264 HR -> 768x768 canvas, 9/4 tiles ----
    spec = SPECS["tiny"]
    diff = build_ref(spec, SEEDS["tiny"], 256, steps=4)
    g = torch.Generator().manual_seed(9)
    cond01 = torch.rand(1, 3, 272, 264, generator=g)
    torch.manual_seed(71)
    img = diff.tiled_sample(batch_size=4, condition_x=cond01, class_label=torch.tensor([0]),
                            class_cond_scale=2.0, num_sample_steps=4)
    save("tiled_tiny", cond01=cond01, label=np.array([0]), ccs=2.0, nsteps=4, seed=71, batch_size=4,
         img=img)
    # single-tile case (<=256): 96x128 HR -> one 256x256 tile
    cond01 = torch.rand(1, 3, 96, 128, generator=g)
    torch.manual_seed(71)
    img = diff.tiled_sample(batch_size=8, condition_x=cond01, class_label=None,
                            num_sample_steps=4)
    save("tiled_tiny_single", cond01=cond01, nsteps=4, seed=71, batch_size=8, img=img)


if __name__ == "__main__":
    main()
