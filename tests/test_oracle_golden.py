"""Pins oracle/srgd_oracle.py against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py; /root/reference/model.py).  CPU only.

Tolerances: the fixtures were computed on the build container's CPU in fp32; re-running the same
fp32 math on another CPU (different oneDNN/MKL kernels, thread counts) reorders sums, so exact
equality is not expected.  2e-4 abs on O(1) activations / eps; scalars to 1e-6 relative.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import srgd_oracle as O

SPECS = {"tiny": (O.UnetSpec(dim=16), 11), "mid": (O.UnetSpec(dim=64), 22), "full": (O.UnetSpec(dim=128), 1234)}
ATOL = 2e-4


def _load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name + ".npz")).items()}


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_param_layout_full():
    shapes = O.param_shapes(O.UnetSpec())
    assert len(shapes) == 280
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert n == 137_569_939
    assert shapes["model.init_conv.weight"] == (128, 6, 7, 7)
    assert shapes["model.ups.0.3.net.0.weight"] == (2048, 1024, 1, 1)
    assert shapes["model.final_conv.weight"] == (3, 128, 1, 1)


def test_schedule_scalars(golden_dir):
    g = _load(golden_dir, "scalars_250")
    steps = torch.linspace(1., 0., 251)
    assert torch.equal(steps, T(g["steps"]))
    tab = T(g["table"])
    for i in range(250):
        s = O.step_scalars(steps[i], steps[i + 1])
        row = torch.stack([s["log_snr"], s["log_snr_next"], s["c"], s["alpha"], s["sigma"],
                           s["alpha_next"], s["var"]])
        torch.testing.assert_close(row, tab[i], rtol=1e-6, atol=1e-7)
    assert abs(float(tab[0, 3]) - 6.74e-3) < 1e-4               # alpha_0, SURVEY.md §7


def test_tile_geometry(golden_dir):
    geo = json.load(open(os.path.join(golden_dir, "geometry.json")))
    for e in geo:
        h, w = e["hw"]
        coord, pad = O.get_coord_and_pad(h, w)
        assert list(coord) == e["coord"] and list(pad) == e["pad"]
        H, W = h + pad[2] + pad[3], w + pad[0] + pad[1]
        c0 = O.get_coords(H, W, 256, 256, 0)
        c1 = c0 if (H <= 256 and W <= 256) else O.get_coords(H - 256, W - 256, 256, 256, 128)
        assert [list(c) for c in c0] == e["c0"] and [list(c) for c in c1] == e["c1"]
        area, apad = O.get_area(c1, H, W)
        assert list(area) == e["area"] and list(apad) == e["apad"]


@pytest.mark.parametrize("tag", ["tiny", "mid", "full"])
def test_unet_forward(golden_dir, tag):
    spec, seed = SPECS[tag]
    g = _load(golden_dir, f"unet_{tag}")
    sd = O.make_state_dict(spec, seed)
    x, cond, lsnr, labels = T(g["x"]), T(g["cond"]), T(g["log_snr"]), T(g["labels"])
    taps = {}
    eps = O.unet_forward(sd, spec, x, lsnr, labels, cond, taps=taps)
    torch.testing.assert_close(eps, T(g["eps_label_cond"]), rtol=0, atol=ATOL)
    torch.testing.assert_close(O.unet_forward(sd, spec, x, lsnr, None, cond), T(g["eps_nolabel_cond"]), rtol=0, atol=ATOL)
    torch.testing.assert_close(O.unet_forward(sd, spec, x, lsnr, labels, None), T(g["eps_label_nocond"]), rtol=0, atol=ATOL)
    torch.testing.assert_close(O.unet_forward(sd, spec, x, lsnr, labels[:1], cond), T(g["eps_label1_cond"]), rtol=0, atol=ATOL)
    for name in ["init_conv", "downs.0.0", "downs.0.2", "downs.3.2", "mid_block1", "ups.0.3", "final_res_block"]:
        ref = T(g["act_" + name])
        a = taps[name]
        if name in ("downs.0.2", "downs.3.2"):
            continue          # reference hook captures attn(x) (pre-residual); oracle tap is attn(x)+x
        sub = a[:, ::max(1, a.shape[1] // 8), ::4, ::4]
        torch.testing.assert_close(sub, ref, rtol=0, atol=ATOL)


def test_p_sample_teacher_forced(golden_dir):
    spec, seed = SPECS["mid"]
    g = _load(golden_dir, "p_sample_mid")
    sd = O.make_state_dict(spec, seed)
    steps = torch.linspace(1., 0., 251)
    cond, label = T(g["cond"]), T(g["label"])
    for ci in range(int(g["ncases"])):
        i, cs, ccs = g[f"c{ci}_meta"]
        i = int(i)
        x, noise = T(g[f"c{ci}_x"]), T(g[f"c{ci}_noise"])
        img, x0 = O.p_sample(sd, spec, x, steps[i], cond, label, float(cs), float(ccs), steps[i + 1], noise=noise)
        # x0 = (x - sigma eps)/alpha amplifies eps error by sigma/alpha (148x at step 0) before the clamp
        torch.testing.assert_close(img, T(g[f"c{ci}_img"]), rtol=0, atol=2e-4)
        amp = float(O.step_scalars(steps[i], steps[i + 1])["sigma"] / O.step_scalars(steps[i], steps[i + 1])["alpha"])
        torch.testing.assert_close(x0, T(g[f"c{ci}_x0"]), rtol=0, atol=max(2e-4, 2e-5 * amp))
        mean, var, _ = O.p_mean_variance(sd, spec, x, steps[i], cond, label, float(cs), float(ccs), steps[i + 1])
        torch.testing.assert_close(mean, T(g[f"c{ci}_mean"]), rtol=0, atol=2e-4)
        torch.testing.assert_close(var, T(g[f"c{ci}_var"]), rtol=1e-6, atol=1e-9)


def test_both_scales_raise():
    spec, seed = SPECS["tiny"]
    sd = O.make_state_dict(spec, seed)
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(NotImplementedError):
        O.p_mean_variance(sd, spec, x, torch.tensor(1.0), x, torch.tensor([0]), 2.0, 2.0, torch.tensor(0.9))
    with pytest.raises(AssertionError):
        O.unet_forward(sd, spec, torch.zeros(1, 3, 36, 64), torch.zeros(1))


@pytest.mark.parametrize("tag", ["tiny", "mid"])
def test_sample_free_running(golden_dir, tag):
    spec, seed = SPECS[tag]
    g = _load(golden_dir, f"sample_{tag}")
    sd = O.make_state_dict(spec, seed)
    torch.manual_seed(int(g["seed"]))
    img = O.sample(sd, spec, 2, T(g["cond01"]), class_label=T(g["label"]), class_cond_scale=float(g["ccs"]),
                   num_sample_steps=int(g["nsteps"]), image_size=64)
    torch.testing.assert_close(img, T(g["img"]), rtol=0, atol=5e-4)


def test_tiled_sample(golden_dir):
    spec, seed = SPECS["tiny"]
    sd = O.make_state_dict(spec, seed)
    g = _load(golden_dir, "tiled_tiny_single")
    torch.manual_seed(int(g["seed"]))
    img = O.tiled_sample(sd, spec, int(g["batch_size"]), T(g["cond01"]), None, num_sample_steps=int(g["nsteps"]))
    torch.testing.assert_close(img, T(g["img"]), rtol=0, atol=5e-4)
    g = _load(golden_dir, "tiled_tiny")
    torch.manual_seed(int(g["seed"]))
    img = O.tiled_sample(sd, spec, int(g["batch_size"]), T(g["cond01"]), T(g["label"]),
                         class_cond_scale=float(g["ccs"]), num_sample_steps=int(g["nsteps"]))
    assert img.shape == (1, 3, 272, 264)
    torch.testing.assert_close(img, T(g["img"]), rtol=0, atol=5e-4)


def test_config1_full_size_three_steps(golden_dir):
    """The oracle at the shipped width and tile size (dim 128, 256x256, label 0, scale 1.0, seed 71) against three
    steps of the unmodified reference's tiled_sample (tests/golden/make_golden_config1.py with GOLDEN_STEPS=3)."""
    g = _load(golden_dir, "config1_3")
    spec = O.UnetSpec()
    sd = O.make_state_dict(spec, 1234, init="torch")
    cond01 = T(g["cond_u8"]).float().div(255.)
    torch.manual_seed(int(g["seed"]))
    img = O.tiled_sample(sd, spec, int(g["batch_size"]), cond01, torch.tensor([int(g["label"])]),
                         num_sample_steps=int(g["steps"]))
    torch.testing.assert_close(img[..., ::2, ::2], T(g["img_sub2"]), rtol=0, atol=5e-4)


def test_step_gating_options(golden_dir):
    """generation_start_steps / guidance_start_steps / class_guidance_start_steps of sample() and tiled_sample()
    (model.py:3196-3228, 3349-3356) against the unmodified reference (tests/golden/make_golden_options.py)."""
    g = _load(golden_dir, "options_tiny")
    spec = O.UnetSpec(dim=16)
    sd = O.make_state_dict(spec, 11)
    gen = torch.Generator().manual_seed(21)
    cond01 = torch.rand(2, 3, 64, 64, generator=gen)
    cond_t = torch.rand(1, 3, 272, 264, generator=gen)
    assert abs(float(cond01.double().sum()) - float(g["cond01_checksum"])) < 1e-6
    assert abs(float(cond_t.double().sum()) - float(g["cond_t_checksum"])) < 1e-6
    torch.manual_seed(71)
    a = O.sample(sd, spec, 2, cond01, class_label=torch.tensor([1]), class_cond_scale=2.5, class_guidance_start_steps=3,
                 generation_start_steps=2, num_sample_steps=8, image_size=64)
    torch.testing.assert_close(a, T(g["sample_a"]), rtol=0, atol=5e-4)
    torch.manual_seed(71)
    b = O.sample(sd, spec, 2, cond01, class_label=torch.tensor([0, 2]), cond_scale=1.7, guidance_start_steps=5,
                 num_sample_steps=8, image_size=64)
    torch.testing.assert_close(b, T(g["sample_b"]), rtol=0, atol=5e-4)
    torch.manual_seed(71)
    t = O.tiled_sample(sd, spec, 4, cond_t, torch.tensor([2]), class_cond_scale=2.0, class_guidance_start_steps=2,
                       generation_start_steps=1, num_sample_steps=4)
    torch.testing.assert_close(t[..., ::2, ::2], T(g["tiled_a_sub2"]), rtol=0, atol=5e-4)


def test_edm_family_oracle_vs_reference_golden(golden_dir):
    """the oracle's restatement of ConditionalElucidatedDiffusionSR (model.py:2059-2560) against
    outputs of the unmodified reference class (tests/golden/make_golden_edm.py): preconditioned forward with each
    guidance kind, stochastic Heun sample_org (incl. generation_start_steps + guidance_start_steps), DPM-Solver++ 2M,
    Heun tiled_sample -- bit-exact.  (The pip base class behind both is the shim's restatement: parity unpinned.)"""
    g = _load(golden_dir, "edm_tiny")
    T = lambda a: torch.from_numpy(np.asarray(a))
    spec, p = O.UnetSpec(dim=16), O.EdmParams()
    sd = O.make_state_dict(spec, 11, prefix="net.")
    cond, big, x, label, n = T(g["cond"]), T(g["big"]), T(g["x"]), T(g["label"]), int(g["steps"])
    gen = lambda: torch.Generator().manual_seed(71)
    with torch.inference_mode():
        for name in ("fwd_plain", "fwd_class", "fwd_cond"):
            sig, cs, ccs = (float(v) for v in g[name + "_meta"])
            assert torch.equal(O.edm_denoise(sd, spec, p, x, sig, cond * 2 - 1, label, cs, ccs, clamp=True), T(g[name])), name
        assert torch.equal(O.edm_sample_heun(sd, spec, p, 2, cond, label, class_cond_scale=2.0, num_sample_steps=n,
                                             generator=gen()), T(g["heun_class"]))
        assert torch.equal(O.edm_sample_heun(sd, spec, p, 1, cond[:1], label, cond_scale=1.5, guidance_start_steps=3,
                                             generation_start_steps=2, num_sample_steps=n, generator=gen()),
                           T(g["heun_cond_start2"]))
        assert torch.equal(O.edm_tiled_sample(sd, spec, p, 5, big, label, class_cond_scale=2.0, num_sample_steps=n,
                                              tile_size=32, tile_stride=32, generator=gen()), T(g["tiled_heun"]))
        assert torch.equal(O.edm_sample_dpmpp(sd, spec, p, 2, cond, label, class_cond_scale=2.0, num_sample_steps=n,
                                              generator=gen()), T(g["dpmpp_class"]))


def test_gaussian_family_oracle_vs_reference_golden(golden_dir):
    """the oracle's restatement of ConditionalGaussianDiffusionSR (model.py:1311-1660) against
    outputs of the unmodified reference class (tests/golden/make_golden_gauss.py): the U-Net with the fixed
    SinusoidalPosEmb, the registered buffers of every beta schedule, model_predictions for every objective, p_sample,
    a full DDPM loop and two DDIM runs (eta 0 / 0.7, both guidance kinds, generation_start_steps) -- bit-exact.
    (The five pip base-class helpers behind both are the shim's restatement: parity unpinned.)"""
    g = _load(golden_dir, "gauss_tiny")
    spec = O.UnetSpec(dim=16, learned_sinusoidal_cond=False)
    sd = O.make_state_dict(spec, 11, prefix="model.")
    assert "model.time_mlp.0.weights" not in sd and sd["model.time_mlp.1.weight"].shape == (64, 16)
    cond, x, label = T(g["cond"]), T(g["x"]), T(g["label"])
    gen = lambda: torch.Generator().manual_seed(71)
    lin = O.GaussParams(1000, 6, "pred_noise", "linear")
    with torch.inference_mode():
        assert torch.equal(O.unet_forward(sd, spec, x, torch.tensor([7, 500]), label, cond * 2 - 1), T(g["unet"]))
        tab = O.gauss_tables(lin)
        for name in ("betas", "alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_log_variance_clipped",
                     "posterior_mean_coef1", "posterior_mean_coef2"):
            assert torch.equal(tab[name], T(g["linear_" + name])), name
        t400 = torch.full((2,), 400, dtype=torch.long)
        for obj, sched in (("pred_noise", "linear"), ("pred_x0", "cosine"), ("pred_v", "sigmoid")):
            p = O.GaussParams(1000, 6, obj, sched)
            tb = O.gauss_tables(p)
            assert torch.equal(tb["alphas_cumprod"], T(g[sched + "_alphas_cumprod"])), sched
            pn, x0 = O.gauss_model_predictions(sd, spec, p, tb, x, t400, cond * 2 - 1, label, 1.0, 2.0,
                                               clip_x_start=True, rederive_pred_noise=True)
            assert torch.equal(pn, T(g[obj + "_noise"])) and torch.equal(x0, T(g[obj + "_x0"])), obj
        img, x0 = O.gauss_p_sample(sd, spec, lin, tab, x, 500, cond * 2 - 1, label, 1.5, 1.0, generator=gen())
        assert torch.equal(img, T(g["p_sample_500"])) and torch.equal(x0, T(g["p_sample_500_x0"]))
        assert torch.equal(O.gauss_p_sample(sd, spec, lin, tab, x, 0, cond * 2 - 1, label)[0], T(g["p_sample_0"]))
        assert torch.equal(O.gauss_sample(sd, spec, O.GaussParams(8, 8, "pred_x0", "cosine"), 2, cond, label,
                                          class_cond_scale=2.0, class_guidance_start_steps=3, generator=gen()),
                           T(g["ddpm_x0_cosine"]))
        assert torch.equal(O.gauss_sample(sd, spec, O.GaussParams(1000, 6, "pred_v", "sigmoid"), 2, cond, label,
                                          class_cond_scale=2.0, generator=gen()), T(g["ddim_v_sigmoid"]))
        assert torch.equal(O.gauss_sample(sd, spec, O.GaussParams(1000, 6, "pred_noise", "linear", 0.7), 2, cond, label,
                                          cond_scale=1.5, generation_start_steps=2, generator=gen()),
                           T(g["ddim_eps_eta"]))


def test_tiled_sample_with_overlapping_tiles(golden_dir):
    """tile_stride < tile_size and tile sizes / strides that do not divide the canvas (tests/golden/make_golden_stride.py,
    unmodified reference): the oracle's in-place, minibatch-by-minibatch canvas update reproduces the reference."""
    g = _load(golden_dir, "tiled_stride_tiny")
    spec = O.UnetSpec(dim=16)
    sd = O.make_state_dict(spec, 11)
    cond01 = torch.rand(1, 3, 104, 120, generator=torch.Generator().manual_seed(7))
    assert abs(float(cond01.double().sum()) - float(g["cond01_checksum"])) < 1e-6
    with torch.inference_mode():
        for name, (tile, stride) in {"t32_s16": (32, 16), "t64_s48": (64, 48), "t48_s48": (48, 48)}.items():
            got = O.tiled_sample(sd, spec, 5, cond01, torch.tensor([1]), class_cond_scale=2.0, num_sample_steps=3,
                                 tile_size=tile, tile_stride=stride, generator=torch.Generator().manual_seed(71))
            torch.testing.assert_close(got, T(g[name]), rtol=0, atol=5e-4, msg=name)
