"""Precision calibration (not a test): how much of the CUDA path's deviation from the fp32 oracle is
inherent to bf16 tensor-core operands?  Runs the oracle ON THE GPU (plain torch) in four numeric
modes plus the srgd_b200 kernels, for both weight inits:
   fp32      : strict fp32 (TF32 off)                         -- the reference
   autocast  : torch.autocast(bfloat16)                       -- what the reference would do with amp
   operands  : conv/linear-attn operands rounded to bf16, everything else fp32 (floor for any
               bf16 tensor-core implementation)
   op+out    : operands and conv outputs rounded to bf16
   srgd_b200 : this repo's kernels
Reports single-forward eps error and 250-step free-running PSNR vs fp32 (B=2, 64x64 U-Net input)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import srgd_oracle as O  # noqa: E402
import model as M  # noqa: E402
import gpu_util as G  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
_conv2d = F.conv2d
MODE = {"m": "fp32"}


def patched_conv2d(x, w, b=None, *a, **k):
    m = MODE["m"]
    if m in ("operands", "op+out"):
        y = _conv2d(x.bfloat16().float(), w.bfloat16().float(), b, *a, **k)
        return y.bfloat16().float() if m == "op+out" else y
    return _conv2d(x, w, b, *a, **k)


def oracle_psample(sd, spec, x, t, tn, c, label, ccs, nz):
    if MODE["m"] == "autocast":
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, _ = O.p_sample(sd, spec, x, t, c, label, 1.0, ccs, tn, noise=nz)
        return out.float()
    out, _ = O.p_sample(sd, spec, x, t, c, label, 1.0, ccs, tn, noise=nz)
    return out


def main():
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 250
    F.conv2d = patched_conv2d
    O.F.conv2d = patched_conv2d
    spec = O.UnetSpec()
    B = 2
    g = torch.Generator().manual_seed(4)
    cond01 = torch.rand(B, 3, 64, 64, generator=g)
    c = (cond01 * 2 - 1).cuda()
    label = torch.tensor([1]).cuda()
    steps = torch.linspace(1., 0., nsteps + 1)
    for init in ("unit", "torch"):
        sd = O.make_state_dict(spec, 1234, init=init)
        gsd = {k: v.cuda() for k, v in sd.items()}
        unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
        diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=64, num_sample_steps=nsteps)
        diff.load_state_dict(sd)
        diff = diff.eval().to("cuda")
        diff.progress = False
        # single forward
        x = torch.randn(B, 3, 64, 64, generator=g).cuda()
        lsnr = torch.tensor([-3.7, 2.1]).cuda()
        lab2 = torch.tensor([0, 2]).cuda()
        with torch.inference_mode():
            MODE["m"] = "fp32"
            ref = O.unet_forward(gsd, spec, x, lsnr, lab2, c)
            ref_null = O.unet_forward(gsd, spec, x, lsnr, None, c)
            print(f"[{init}] eps rms {float(ref.pow(2).mean().sqrt()):.4f}; cond-null rms diff {float((ref - ref_null).pow(2).mean().sqrt()):.4f}")
            for m in ("autocast", "operands", "op+out"):
                MODE["m"] = m
                if m == "autocast":
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        e = O.unet_forward(gsd, spec, x, lsnr, lab2, c).float()
                else:
                    e = O.unet_forward(gsd, spec, x, lsnr, lab2, c)
                d = e - ref
                print(f"[{init}] forward {m:9s}: eps max-abs {float(d.abs().max()):.4f} rms {float(d.pow(2).mean().sqrt()):.5f}")
            e = diff.model(x, lsnr, lab2, c)
            d = e - ref
            print(f"[{init}] forward srgd_b200: eps max-abs {float(d.abs().max()):.4f} rms {float(d.pow(2).mean().sqrt()):.5f}")
            # free running
            for ccs in (1.0, 3.0):
                torch.manual_seed(71)
                noises = [torch.randn(B, 3, 64, 64, device="cuda") for _ in range(nsteps)]
                finals = {}
                for m in ("fp32", "autocast", "operands", "op+out"):
                    MODE["m"] = m
                    xx = noises[0]
                    for i in range(nsteps):
                        nz = noises[i + 1] if i + 1 < nsteps else None
                        xx = oracle_psample(gsd, spec, xx, steps[i].cuda(), steps[i + 1].cuda(), c, label, ccs, nz)
                    finals[m] = ((xx.clamp(-1, 1) + 1) * 0.5).cpu()
                torch.manual_seed(71)
                finals["srgd_b200"] = diff.sample(batch_size=B, condition_x=cond01.cuda(), class_label=label,
                                                  class_cond_scale=ccs, num_sample_steps=nsteps).cpu()
                for m in ("autocast", "operands", "op+out", "srgd_b200"):
                    print(f"[{init}] {nsteps}-step free-running ccs={ccs}: {m:9s} PSNR vs fp32 {G.psnr(finals[m], finals['fp32']):.2f} dB")
        MODE["m"] = "fp32"
        del diff, unet


if __name__ == "__main__":
    main()
