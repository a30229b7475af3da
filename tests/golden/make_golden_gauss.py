"""Generate tests/golden/gauss_tiny.npz from the UNMODIFIED reference's ConditionalGaussianDiffusionSR
(/root/reference/model.py:1311-1660) behind oracle/_shim (whose `GaussianDiffusion` base restates the pip package's five
helper formulas: parity unpinned for those; schedules, buffers, guidance, DDPM and DDIM loops are the reference's own
code).  Run in the build container only:

    python tests/golden/make_golden_gauss.py

Cases (dim-16 U-Net with the fixed SinusoidalPosEmb, seeded CPU generator re-seeded per case like inference.py:47-51):
  unet            one U-Net forward at integer timesteps [7, 500]
  pred_*          model_predictions at t = 400 for every objective (clip + rederive on), with class guidance
  p_sample        one ancestral step at t = 500 (pred_noise, linear betas) with LR-condition guidance, and at t = 0
  ddpm_x0_cosine  full p_sample_loop of an 8-timestep model (pred_x0, cosine betas), class guidance from step 3
  ddim_v_sigmoid  DDIM, 6 of 1000 steps (pred_v, sigmoid betas, eta 0), class guidance 2.0
  ddim_eps_eta    DDIM, 6 of 1000 steps (pred_noise, linear betas, eta 0.7), LR-condition guidance 1.5,
                  generation_start_steps 2
"""
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
ref.tqdm = lambda it, **kw: it                       # silence the progress bars
from oracle import srgd_oracle as O

torch.set_num_threads(os.cpu_count())
SPEC, SEED = O.UnetSpec(dim=16, learned_sinusoidal_cond=False), 11


def build(**kw):
    unet = ref.ConditionalSRUnet(dim=SPEC.dim, dim_mults=SPEC.dim_mults, full_attn=SPEC.full_attn, learned_variance=False,
                                 learned_sinusoidal_cond=False, flash_attn=False, pixel_shuffle_upsample=True,
                                 num_classes=3)
    m = ref.ConditionalGaussianDiffusionSR(unet, image_size=32, **kw).eval()
    sd = O.make_state_dict(SPEC, SEED, prefix="model.")
    own = m.state_dict()
    assert [k for k in own if k.startswith("model.")] == list(sd.keys())
    m.load_state_dict(sd, strict=False)              # the buffers keep their constructor values
    return m


@torch.inference_mode()
def main():
    g = torch.Generator().manual_seed(5)
    cond = torch.rand(2, 3, 32, 32, generator=g)
    x = torch.randn(2, 3, 32, 32, generator=g)
    noise = torch.randn(2, 3, 32, 32, generator=g)
    label = torch.tensor([2])
    out = dict(cond=cond, x=x, noise=noise, label=label)
    m = build(timesteps=1000, sampling_timesteps=6, objective='pred_noise', beta_schedule='linear')
    out["unet"] = m.model(x, torch.tensor([7, 500]), label, cond * 2 - 1)
    for name in ("betas", "alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_log_variance_clipped",
                 "posterior_mean_coef1", "posterior_mean_coef2"):
        out["linear_" + name] = getattr(m, name)
    t400 = torch.full((2,), 400, dtype=torch.long)
    for obj, sched in (("pred_noise", "linear"), ("pred_x0", "cosine"), ("pred_v", "sigmoid")):
        mm = build(timesteps=1000, sampling_timesteps=6, objective=obj, beta_schedule=sched)
        pn, x0 = mm.model_predictions(x, t400, cond * 2 - 1, label, 1.0, 2.0, clip_x_start=True, rederive_pred_noise=True)
        out[obj + "_noise"], out[obj + "_x0"] = pn, x0
        out[sched + "_alphas_cumprod"] = mm.alphas_cumprod
    torch.manual_seed(71)
    img, x0 = m.p_sample(x, 500, cond * 2 - 1, label, 1.5, 1.0)
    out["p_sample_500"], out["p_sample_500_x0"] = img, x0
    img, x0 = m.p_sample(x, 0, cond * 2 - 1, label, 1.0, 1.0)
    out["p_sample_0"] = img
    m8 = build(timesteps=8, sampling_timesteps=8, objective='pred_x0', beta_schedule='cosine')
    assert not m8.is_ddim_sampling
    torch.manual_seed(71)
    out["ddpm_x0_cosine"] = m8.sample(batch_size=2, condition_x=cond, class_label=label, class_cond_scale=2.0,
                                      class_guidance_start_steps=3)
    mv = build(timesteps=1000, sampling_timesteps=6, objective='pred_v', beta_schedule='sigmoid')
    torch.manual_seed(71)
    out["ddim_v_sigmoid"] = mv.sample(batch_size=2, condition_x=cond, class_label=label, class_cond_scale=2.0)
    me = build(timesteps=1000, sampling_timesteps=6, objective='pred_noise', beta_schedule='linear', ddim_sampling_eta=0.7)
    torch.manual_seed(71)
    out["ddim_eps_eta"] = me.sample(batch_size=2, condition_x=cond, class_label=label, cond_scale=1.5,
                                    generation_start_steps=2)
    path = os.path.join(OUT, "gauss_tiny.npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()})
    print(f"gauss_tiny: {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
