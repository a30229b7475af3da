"""GPU parity of the discrete-time sampler family (srgd_b200/gaussian.py; reference ConditionalGaussianDiffusionSR,
model.py:1311-1660: DDPM ancestral sampling and DDIM) against the fp32 oracle
(oracle/srgd_oracle.py gauss_*, pinned bit-exactly to the unmodified reference class by tests/golden/gauss_tiny.npz; the
five helper formulas of its pip base class are a restatement: parity unpinned, like `Attend`).
Run with `pytest -m gpu` on a B200.

The oracle runs on the same GPU in strict fp32 and draws from torch's CUDA generator with the reference's shapes and
order, the stream the product path consumes after the same torch.manual_seed."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_util as G  # noqa: E402
from oracle import srgd_oracle as O  # noqa: E402  (checker only)
import model as M  # noqa: E402
from test_gpu_unet import _oracle_on_gpu  # noqa: E402

_cache = {}


def build(p: O.GaussParams, dim=64, image_size=64):
    key = (p, dim, image_size)
    if key not in _cache:
        spec = O.UnetSpec(dim=dim, learned_sinusoidal_cond=False)
        sd = O.make_state_dict(spec, 22, prefix="model.", init="torch")
        unet = M.ConditionalSRUnet(dim=dim, learned_sinusoidal_cond=False, num_classes=3)
        m = M.ConditionalGaussianDiffusionSR(unet, image_size=image_size, timesteps=p.timesteps,
                                             sampling_timesteps=p.sampling_timesteps, objective=p.objective,
                                             beta_schedule=p.beta_schedule, ddim_sampling_eta=p.ddim_sampling_eta)
        assert not m.load_state_dict(sd, strict=False).unexpected_keys
        m = m.eval().to("cuda")
        m.progress = False
        _cache[key] = (m, _oracle_on_gpu(sd), spec)
    return _cache[key]


def test_unet_with_the_fixed_sinusoidal_embedding_vs_oracle():
    """ConditionalSRUnet(learned_sinusoidal_cond=False): SinusoidalPosEmb(dim) on integer timesteps (model.py:209-221,
    600) feeding the same time MLP; eps against the fp32 oracle at small / middle / large timesteps in one batch."""
    m, gsd, spec = build(O.GaussParams())
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 3, 64, 64, generator=g).cuda()
    cond = (torch.rand(3, 3, 64, 64, generator=g) * 2 - 1).cuda()
    t = torch.tensor([3, 500, 999], device="cuda")
    label = torch.tensor([0, 2, 1], device="cuda")
    got = m.model(x, t, label, cond)
    with torch.inference_mode():
        ref = O.unet_forward(gsd, spec, x, t, label, cond)
    err, rms = float((got - ref).abs().max()), float((got - ref).pow(2).mean().sqrt())
    print(f"fixed-sinusoidal U-Net: eps max-abs {err:.5f} rms {rms:.5f} (|eps| max {float(ref.abs().max()):.3f})")
    assert err <= 3e-2 and rms <= 5e-3
    # the embedding itself distinguishes neighbouring timesteps: t and t + 1 give different eps
    assert float((m.model(x, t + 0, label, cond) - m.model(x, (t - 1).clamp(min=0), label, cond)).abs().max()) > 0


@pytest.mark.parametrize("objective,schedule", [("pred_noise", "linear"), ("pred_x0", "cosine"), ("pred_v", "sigmoid")])
def test_predictions_and_p_sample_teacher_forced(objective, schedule):
    """model_predictions (clip + rederive, the DDIM call) and p_sample (the DDPM call) at t = 700 / 250 / 0 with each
    guidance kind, teacher-forced with the same noise: img max-abs <= 1e-2 (the bf16 tolerance of SURVEY section 8d)."""
    p = O.GaussParams(1000, 1000, objective, schedule)
    m, gsd, spec = build(p)
    tab = O.gauss_tables(p)
    g = torch.Generator().manual_seed(8)
    cond = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).cuda()
    label = torch.tensor([1]).cuda()
    for t, cs, ccs in ((700, 1.0, 1.0), (250, 1.0, 2.0), (40, 1.5, 1.0), (0, 1.0, 1.0)):
        k = torch.full((2,), t, device="cuda", dtype=torch.long)
        a = float(tab["sqrt_alphas_cumprod"][t])
        x0_true = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
        x = (a * x0_true + (1 - a * a) ** 0.5 * torch.randn(2, 3, 64, 64, generator=g)).cuda()
        noise = torch.randn(2, 3, 64, 64, generator=g).cuda()
        img, x0 = m.p_sample(x, t, cond, label, cs, ccs, noise=noise)
        pn, x0c = m.model_predictions(x, k, cond, label, cs, ccs, clip_x_start=True, rederive_pred_noise=True)
        with torch.inference_mode():
            r_img, r_x0 = O.gauss_p_sample(gsd, spec, p, tab, x, t, cond, label, cs, ccs, noise=noise)
            r_pn, r_x0c = O.gauss_model_predictions(gsd, spec, p, tab, x, k, cond, label, cs, ccs, True, True)
        e_img, e_x0 = float((img - r_img).abs().max()), float((x0 - r_x0).abs().max())
        # pred_noise = (x / sqrt(ac) - x0) / sqrt(1/ac - 1): at small t the division amplifies the x0 error
        amp = 1.0 / float(tab["sqrt_recipm1_alphas_cumprod"][t])
        e_pn = float((pn - r_pn).abs().max())
        print(f"{objective}/{schedule} t={t} cs={cs} ccs={ccs}: img {e_img:.5f} x0 {e_x0:.5f} pred_noise {e_pn:.5f} (x{amp:.1f})")
        # at t = 0 the posterior mean IS x_start (coef1 = 1); with pred_x0 that is the raw network output
        assert e_img <= (3e-2 if (objective == "pred_x0" and t == 0) else 1e-2)
        # x_start = x / sqrt(ac) - sqrt(1/ac - 1) * eps amplifies the network error at large t for pred_noise (the
        # rederived noise divides it out again); the other objectives amplify on the way to pred_noise instead
        recipm1 = float(tab["sqrt_recipm1_alphas_cumprod"][t])
        x0_tol = 3e-2 * max(1.0, recipm1) if objective == "pred_noise" else 5e-2
        pn_tol = 5e-2 if objective == "pred_noise" else 5e-2 * max(1.0, amp)
        assert float((x0c - r_x0c).abs().max()) <= x0_tol and e_pn <= pn_tol
    with pytest.raises(NotImplementedError):
        m.p_sample(x, 5, cond, label, 2.0, 2.0)


def test_ddim_free_running_vs_oracle():
    """sample() -> ddim_sample: 32 of 1000 steps, pred_v / sigmoid betas, eta 0.4, class guidance 2.0, B = 2, seed 71."""
    p = O.GaussParams(1000, 32, "pred_v", "sigmoid", 0.4)
    m, gsd, spec = build(p)
    g = torch.Generator().manual_seed(4)
    cond01 = torch.rand(2, 3, 64, 64, generator=g).cuda()
    label = torch.tensor([1]).cuda()
    torch.manual_seed(71)
    img = m.sample(batch_size=2, condition_x=cond01, class_label=label, class_cond_scale=2.0)
    torch.manual_seed(71)
    with torch.inference_mode():
        ref = O.gauss_sample(gsd, spec, p, 2, cond01, label, class_cond_scale=2.0)
    psnr = G.psnr(img.cpu(), ref.cpu())
    print(f"DDIM 32 steps: PSNR {psnr:.2f} dB, max-abs {float((img - ref).abs().max()):.4f}")
    assert img.shape == (2, 3, 64, 64) and psnr >= 45.0


def test_ddpm_free_running_and_options_vs_oracle():
    """sample() -> p_sample_loop over all 48 timesteps of a 48-step model (pred_noise, linear betas) with LR-condition
    guidance from step 6 and generation_start_steps 4; with_images list lengths; num_sample_steps override for DDIM."""
    p = O.GaussParams(48, 48, "pred_noise", "linear")
    m, gsd, spec = build(p)
    assert not m.is_ddim_sampling
    g = torch.Generator().manual_seed(6)
    cond01 = torch.rand(2, 3, 64, 64, generator=g).cuda()
    label = torch.tensor([0]).cuda()
    kw = dict(cond_scale=1.5, guidance_start_steps=6, generation_start_steps=4)
    torch.manual_seed(71)
    img, frames, x0s = m.sample(batch_size=2, condition_x=cond01, class_label=label, with_images=True,
                                with_x0_images=True, **kw)
    torch.manual_seed(71)
    with torch.inference_mode():
        ref = O.gauss_sample(gsd, spec, p, 2, cond01, label, **kw)
    psnr = G.psnr(img.cpu(), ref.cpu())
    print(f"DDPM 48 steps: PSNR {psnr:.2f} dB, max-abs {float((img - ref).abs().max()):.4f}")
    assert len(frames) == 1 + 44 and len(x0s) == 1 + 44 and psnr >= 45.0
    p2 = O.GaussParams(1000, 250, "pred_x0", "cosine")
    m2, gsd2, _ = build(p2)
    torch.manual_seed(71)
    img = m2.sample(batch_size=2, condition_x=cond01, class_label=label, num_sample_steps=12)
    torch.manual_seed(71)
    with torch.inference_mode():
        ref = O.gauss_sample(gsd2, spec, p2, 2, cond01, label, num_sample_steps=12)
    assert G.psnr(img.cpu(), ref.cpu()) >= 45.0


def test_gauss_kernel_is_exact():
    """srgd_gauss_update against the reference's fp32 op sequence evaluated by torch on the CPU, for every objective
    and update mode, with and without guidance / noise: bit-exact."""
    import ctypes as C
    from srgd_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    n = 3 * 64 * 64 + 5
    hx, hc, hn, hz = (torch.randn(n, generator=g) for _ in range(4))
    x, oc, on, z = (t.cuda() for t in (hx, hc, hn, hz))
    f = lambda v: torch.tensor(v, dtype=torch.float32)
    vals = dict(guidance_scale=2.5, sqrt_recip_ac=1.31, sqrt_recipm1_ac=0.847, sqrt_ac=0.763, sqrt_1m_ac=0.646,
                coef1=0.21, coef2=0.78, noise_scale=0.37, sqrt_ac_next=0.81, c=0.55)
    v = {k: f(val) for k, val in vals.items()}
    st = _lib.current_stream()
    for obj in (_lib.OBJ_PRED_NOISE, _lib.OBJ_PRED_X0, _lib.OBJ_PRED_V):
        for mode in (_lib.GAUSS_DDPM, _lib.GAUSS_DDIM, _lib.GAUSS_DDIM_LAST):
            for guided, noisy, clip, red in ((True, True, 1, 1), (False, False, 1, 0), (True, False, 0, 0)):
                s = _lib.GaussScalars(obj, mode, clip, red, *[vals[k] for k in
                                      ("guidance_scale", "sqrt_recip_ac", "sqrt_recipm1_ac", "sqrt_ac", "sqrt_1m_ac",
                                       "coef1", "coef2", "noise_scale", "sqrt_ac_next", "c")])
                img, x0, pn = (torch.empty_like(x) for _ in range(3))
                _lib.check(lib.srgd_gauss_update(_lib.ptr(x), _lib.ptr(oc), _lib.ptr(on) if guided else None,
                                                 _lib.ptr(z) if noisy else None, _lib.ptr(img), _lib.ptr(x0),
                                                 _lib.ptr(pn), n, C.byref(s), st))
                out = hn + (hc - hn) * v["guidance_scale"] if guided else hc
                if obj == _lib.OBJ_PRED_NOISE:
                    r0 = v["sqrt_recip_ac"] * hx - v["sqrt_recipm1_ac"] * out
                elif obj == _lib.OBJ_PRED_X0:
                    r0 = out
                else:
                    r0 = v["sqrt_ac"] * hx - v["sqrt_1m_ac"] * out
                if clip:
                    r0 = r0.clamp(-1., 1.)
                rn = out
                if obj != _lib.OBJ_PRED_NOISE or (clip and red):
                    rn = (v["sqrt_recip_ac"] * hx - r0) / v["sqrt_recipm1_ac"]
                if mode == _lib.GAUSS_DDPM:
                    ri = v["coef1"] * r0 + v["coef2"] * hx
                    if noisy:
                        ri = ri + v["noise_scale"] * hz
                elif mode == _lib.GAUSS_DDIM:
                    ri = r0 * v["sqrt_ac_next"] + v["c"] * rn
                    if noisy:
                        ri = ri + v["noise_scale"] * hz
                else:
                    ri = r0
                tag = (obj, mode, guided, noisy, clip, red)
                assert torch.equal(x0.cpu(), r0), tag
                assert torch.equal(pn.cpu(), rn), tag
                assert torch.equal(img.cpu(), ri), tag
