"""GPU parity of the whole hot path through the reference-facing API (model.py names) against the
CPU oracle and the committed reference goldens -- run with `pytest -m gpu` on a B200.

Tolerances (BASELINE.json north_star): teacher-forced per-step max-abs error on img_next <= 1e-2
(bf16 kernels), final image PSNR >= 45 dB.  Raw eps is held to 6e-2 max-abs / 1.2e-2 rms (torch's
own bf16 autocast differs from fp32 by 1.6-1.9e-2 max-abs on one forward).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_util as G  # noqa: E402
from oracle import srgd_oracle as O  # noqa: E402  (checker only)
import model as M  # noqa: E402  repo-root drop-in module

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SPECS = {"mid": (O.UnetSpec(dim=64), 22), "full": (O.UnetSpec(dim=128), 1234)}
_cache = {}


def build(tag, image_size=64, steps=250):
    key = (tag, image_size, steps)
    if key not in _cache:
        spec, seed = SPECS[tag]
        unet = M.ConditionalSRUnet(dim=spec.dim, dim_mults=spec.dim_mults, full_attn=spec.full_attn,
                                   learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
        diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=image_size,
                                                              num_sample_steps=steps)
        sd = O.make_state_dict(spec, seed)
        diff.load_state_dict(sd, strict=True)
        diff = diff.eval().to("cuda")
        diff.progress = False
        _cache[key] = (diff, sd, spec)
    return _cache[key]


def make_diffusion(spec, sd, image_size, steps):
    unet = M.ConditionalSRUnet(dim=spec.dim, dim_mults=spec.dim_mults, full_attn=spec.full_attn,
                               learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=image_size, num_sample_steps=steps)
    diff.load_state_dict(sd, strict=True)
    diff = diff.eval().to("cuda")
    diff.progress = False
    return diff


def T(a):
    return torch.from_numpy(np.asarray(a))


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.mark.parametrize("tag", ["mid", "full"])
def test_unet_eps_vs_reference_golden(tag):
    diff, sd, spec = build(tag)
    g = load(f"unet_{tag}")
    x, cond, lsnr, labels = (T(g[k]).cuda() for k in ("x", "cond", "log_snr", "labels"))
    for key, lab, cnd in (("eps_label_cond", labels, cond), ("eps_nolabel_cond", None, cond),
                          ("eps_label_nocond", labels, None), ("eps_label1_cond", labels[:1], cond)):
        eps = diff.model(x, lsnr, lab, cnd).cpu()
        ref = T(g[key])
        err = (eps - ref).abs()
        print(f"{tag} {key}: max {float(err.max()):.4f} rms {float(err.pow(2).mean().sqrt()):.5f} "
              f"ref rms {float(ref.pow(2).mean().sqrt()):.3f}")
        assert float(err.max()) < 6e-2 and float(err.pow(2).mean().sqrt()) < 1.2e-2, key


def test_unet_debug_conv_path_agrees():
    """CUDA-core direct conv + stand-alone GN statistics (debug path) vs the tcgen05 path."""
    diff, sd, spec = build("full")
    g = load("unet_full")
    x, cond, lsnr, labels = (T(g[k]).cuda() for k in ("x", "cond", "log_snr", "labels"))
    a = diff.model(x, lsnr, labels, cond)
    diff.model.conv_impl = 3
    try:
        b = diff.model(x, lsnr, labels, cond)
    finally:
        diff.model.conv_impl = 0
    # two bf16 evaluations of the same network (tanh-form SiLU + fused LinearAttention vs exact SiLU + unfused):
    # held to the same eps tolerance as the comparison with the fp32 reference above
    assert float((a - b).abs().max()) < 6e-2


def test_unet_is_deterministic_and_batch_consistent():
    diff, sd, spec = build("full")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 3, 64, 64, generator=g).cuda()
    cond = (torch.rand(4, 3, 64, 64, generator=g) * 2 - 1).cuda()
    lsnr = torch.tensor([-2.0, 0.5, 3.0, 7.0]).cuda()
    lab = torch.tensor([0, 1, 2, 0]).cuda()
    a = diff.model(x, lsnr, lab, cond)
    b = diff.model(x, lsnr, lab, cond)
    assert torch.equal(a, b)                                  # no atomics on the data path
    c = diff.model(x[1:3], lsnr[1:3], lab[1:3], cond[1:3])
    # rows are independent; the LinearAttention context is merged from per-CTA partials whose split depends on
    # B (fp32 re-association, then one bf16 rounding of the 32x32 context), hence not bit-identical across B
    assert float((a[1:3] - c).abs().max()) < 2e-3


def test_input_validation():
    diff, sd, spec = build("full")
    with pytest.raises(AssertionError):
        diff.model(torch.zeros(1, 3, 36, 64, device="cuda"), torch.zeros(1, device="cuda"))
    with pytest.raises(RuntimeError):
        diff.model(torch.zeros(1, 3, 64, 64), torch.zeros(1))            # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        diff.p_sample(torch.zeros(1, 3, 64, 64, device="cuda"), torch.tensor(1.0), None, None, 2.0, 2.0,
                      torch.tensor(0.9))
    # through the sampler wrappers too: the divisibility assert (model.py:679) and mismatched condition sizes
    with pytest.raises(AssertionError):
        diff.p_sample(torch.zeros(1, 3, 60, 64, device="cuda"), torch.tensor(0.5), None, None, 1.0, 1.0, torch.tensor(0.4))
    with pytest.raises(RuntimeError, match="must match"):
        diff.p_sample(torch.zeros(2, 3, 64, 64, device="cuda"), torch.tensor(0.5), torch.zeros(2, 3, 32, 32, device="cuda"),
                      None, 1.0, 1.0, torch.tensor(0.4))
    with pytest.raises(RuntimeError, match="must match"):
        diff.sample(batch_size=2, condition_x=torch.rand(1, 3, 64, 64, device="cuda"), num_sample_steps=2)


def test_p_sample_teacher_forced_vs_reference_golden():
    diff, sd, spec = build("mid")
    g = load("p_sample_mid")
    steps = torch.linspace(1., 0., 251)
    cond, label = T(g["cond"]).cuda(), T(g["label"]).cuda()
    for ci in range(int(g["ncases"])):
        i, cs, ccs = g[f"c{ci}_meta"]
        i = int(i)
        x, noise = T(g[f"c{ci}_x"]).cuda(), T(g[f"c{ci}_noise"]).cuda()
        img, x0 = diff.p_sample(x, steps[i], cond, label, float(cs), float(ccs), steps[i + 1], noise=noise)
        err = float((img.cpu() - T(g[f"c{ci}_img"])).abs().max())
        err0 = float((x0.cpu() - T(g[f"c{ci}_x0"])).abs().max())
        print(f"case {ci} step {i} cs {cs} ccs {ccs}: img_next max-abs {err:.5f}  x0 max-abs {err0:.5f}")
        assert err <= 1e-2, (ci, err)
        mean, var, _ = diff.p_mean_variance(x, steps[i], cond, label, float(cs), float(ccs), steps[i + 1])
        assert float((mean.cpu() - T(g[f"c{ci}_mean"])).abs().max()) <= 1e-2
        assert abs(float(var) - float(g[f"c{ci}_var"])) <= 1e-7


def _oracle_on_gpu(sd):
    """The oracle is plain functional PyTorch: run it on the GPU in strict fp32 (TF32 off) so that a
    250-step free-running reference takes seconds instead of minutes of host time.  With tensors on
    the GPU its torch.randn calls draw from torch's CUDA generator with the reference's shapes and
    order -- the same stream the product path consumes after the same torch.manual_seed."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return {k: v.cuda() for k, v in sd.items()}


# Final-image PSNR bars.  north_star: >= 45 dB vs the reference at the full schedule, on random-init
# weights of the reference constructors = init "torch".  The golden fixtures use a
# harsher unit-gain init ("unit", activations 3x larger, strong class dependence); there the floor of
# ANY bf16-operand tensor-core implementation is 49.7 dB (scale 1.0) / 41.8 dB (scale 3.0) and torch's
# own bf16 autocast reaches 47.3 / 39.6 dB (profiles/r01_precision_study.log), so the CFG 3.0 bar for
# that init is set at 38 dB (must stay within ~2 dB of torch autocast), not 45.
FREE_RUNNING = [("torch", 1.0, 250, 45.0), ("torch", 3.0, 250, 45.0), ("unit", 1.0, 250, 45.0),
                ("unit", 3.0, 250, 38.0)]


@pytest.mark.parametrize("init,ccs,nsteps,bar", FREE_RUNNING)
def test_free_running_psnr_vs_oracle(init, ccs, nsteps, bar):
    """sample() end to end (full-width U-Net, 64x64, B=2, label 1, seed 71) vs the fp32 oracle."""
    spec = O.UnetSpec()
    sd = O.make_state_dict(spec, 1234, init=init)
    diff = make_diffusion(spec, sd, 64, nsteps)
    g = torch.Generator().manual_seed(4)
    cond01 = torch.rand(2, 3, 64, 64, generator=g).cuda()
    label = torch.tensor([1]).cuda()
    torch.manual_seed(71)
    img = diff.sample(batch_size=2, condition_x=cond01, class_label=label, class_cond_scale=ccs,
                      num_sample_steps=nsteps).cpu()
    gsd = _oracle_on_gpu(sd)
    torch.manual_seed(71)
    with torch.inference_mode():
        ref = O.sample(gsd, spec, 2, cond01, class_label=label, class_cond_scale=ccs, num_sample_steps=nsteps,
                       image_size=64).cpu()
    p = G.psnr(img, ref)
    print(f"free-running {nsteps} steps init={init} ccs={ccs}: PSNR {p:.2f} dB, max-abs {float((img - ref).abs().max()):.4f}")
    assert p >= bar


def test_tiled_sample_vs_oracle():
    """inference.py's path: 272x264 HR -> 768x768 canvas, 9 / 4 alternating tiles, odd-step re-noise,
    batch_size 4 (chunks 4,4,1), CFG 2.0, full 250-step schedule; dim-64 U-Net, reference-style init."""
    spec = O.UnetSpec(dim=64)
    sd = O.make_state_dict(spec, 22, init="torch")
    diff = make_diffusion(spec, sd, 256, 250)
    g = torch.Generator().manual_seed(9)
    cond01 = torch.rand(1, 3, 272, 264, generator=g).cuda()
    label = torch.tensor([0]).cuda()
    torch.manual_seed(71)
    img = diff.tiled_sample(batch_size=4, condition_x=cond01, class_label=label, class_cond_scale=2.0,
                            num_sample_steps=250).cpu()
    assert img.shape == (1, 3, 272, 264) and float(img.min()) >= 0 and float(img.max()) <= 1
    gsd = _oracle_on_gpu(sd)
    torch.manual_seed(71)
    with torch.inference_mode():
        ref = O.tiled_sample(gsd, spec, 4, cond01, label, class_cond_scale=2.0, num_sample_steps=250).cpu()
    p = G.psnr(img, ref)
    print(f"tiled_sample 250 steps: PSNR {p:.2f} dB, max-abs {float((img - ref).abs().max()):.4f}")
    assert p >= 45.0


def test_config1_vs_reference_golden():
    """BASELINE.json configs[0] end to end against the UNMODIFIED reference run on the CPU in fp32
    (tests/golden/make_golden_config1.py: shipped conf dim 128, 250 steps, one synthetic 64x64 LR image through PIL
    bicubic x4, label 0, class_cond_scale 1.0, seed 71, inference.py's tiled_sample call).  `rng_device = "cpu"`
    replays the reference's noise stream (global CPU generator, same shapes in the same order).
    Bar: final-image PSNR >= 45 dB (north_star)."""
    from PIL import Image
    g = load("config1_full")
    spec = O.UnetSpec()
    diff = make_diffusion(spec, O.make_state_dict(spec, 1234, init="torch"), 256, int(g["steps"]))
    diff.rng_device = "cpu"
    hr = Image.fromarray(g["lr"], mode="RGB").resize((256, 256), resample=Image.BICUBIC)      # inference.py:71-74
    cond01 = torch.from_numpy(np.array(hr, dtype=np.uint8)).permute(2, 0, 1).float().div(255.)[None]
    assert torch.equal(cond01, T(g["cond_u8"]).float().div(255.)), "PIL bicubic pre-upscale differs from the fixture"
    torch.manual_seed(int(g["seed"]))
    img = diff.tiled_sample(batch_size=int(g["batch_size"]), condition_x=cond01.cuda(),
                            class_label=torch.tensor([int(g["label"])]).cuda(), class_cond_scale=1.0,
                            num_sample_steps=int(g["steps"])).cpu()
    ref = T(g["img"])
    p = G.psnr(img, ref)
    print(f"config 1 (reference CPU fp32 vs B200 bf16, {int(g['steps'])} steps): PSNR {p:.2f} dB, "
          f"max-abs {float((img - ref).abs().max()):.4f}")
    assert img.shape == ref.shape and p >= 45.0


def test_config3_cfg_vs_reference_golden():
    """Classifier-free guidance end to end against the UNMODIFIED reference on the CPU (fp32): `sample()` at batch 1,
    test_label 2, class_cond_scale 3.0 -- two sequential U-Net calls per step there (model.py:3151-3154), ONE 2x-batch
    launch sequence + the guidance combine fused into the sampler update here -- full 250-step schedule, seed 71,
    same noise stream (tests/golden/make_golden_config3.py).  Bar: final-image PSNR >= 45 dB."""
    from PIL import Image
    if not os.path.exists(os.path.join(GOLDEN, "config3_full.npz")):
        pytest.skip("tests/golden/config3_full.npz not generated (25 CPU-minutes of the reference)")
    g = load("config3_full")
    spec = O.UnetSpec()
    diff = make_diffusion(spec, O.make_state_dict(spec, 1234, init="torch"), 256, int(g["steps"]))
    diff.rng_device = "cpu"
    hr = Image.fromarray(g["lr"], mode="RGB").resize((256, 256), resample=Image.BICUBIC)
    cond01 = torch.from_numpy(np.array(hr, dtype=np.uint8)).permute(2, 0, 1).float().div(255.)[None]
    torch.manual_seed(int(g["seed"]))
    img = diff.sample(batch_size=1, condition_x=cond01.cuda(), class_label=torch.tensor([int(g["label"])]).cuda(),
                      class_cond_scale=float(g["ccs"]), num_sample_steps=int(g["steps"])).cpu()
    ref = T(g["img"])
    p = G.psnr(img, ref)
    print(f"config 3 / CFG {float(g['ccs'])} (reference CPU fp32 vs B200 bf16, {int(g['steps'])} steps): PSNR {p:.2f} dB, "
          f"max-abs {float((img - ref).abs().max()):.4f}")
    assert img.shape == ref.shape and p >= 45.0


def test_launch_count_reported():
    diff, sd, spec = build("full")
    x = torch.randn(1, 3, 64, 64, device="cuda")
    diff.model(x, torch.zeros(1, device="cuda"), None, None)
    assert diff.model.last_launches > 100


def test_reload_through_the_diffusion_wrapper_repacks_the_device_weights():
    """a model that has already run on the GPU and then receives new weights through the
    documented `ema_model.module.load_state_dict(ckpt['ema_model'])` path must sample with the NEW weights."""
    spec = O.UnetSpec(dim=64)
    sd_a, sd_b = O.make_state_dict(spec, 22), O.make_state_dict(spec, 23)
    diff = make_diffusion(spec, sd_a, 64, 250)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 3, 64, 64, generator=g).cuda()
    lsnr = torch.tensor([0.3, -1.0]).cuda()
    a = diff.model(x, lsnr, None, None)
    diff.load_state_dict({k: v.cuda() for k, v in sd_b.items()}, strict=True)       # nested load, model already on GPU
    b = diff.model(x, lsnr, None, None)
    fresh = make_diffusion(spec, sd_b, 64, 250).model(x, lsnr, None, None)
    assert torch.equal(b, fresh) and not torch.equal(a, b)
    with torch.no_grad():                                                           # in-place edit of one parameter
        diff.model.final_conv.bias.add_(0.5)
    c = diff.model(x, lsnr, None, None)
    assert float((c - b - 0.5).abs().max()) < 1e-5


def test_class_guidance_sharing_is_bit_identical_to_the_full_batch(monkeypatch):
    """Class guidance runs rows b and b + B on the same x and condition (model.py:3151-3154): the input pack, init_conv
    and the first conv3x3 + GroupNorm statistics are computed once and broadcast (unet.cu `share`).  Every broadcast value
    is the value the full 2B-row launch computes for both rows, so the result must not change by a bit."""
    diff, sd, spec = build("full", image_size=256)
    g = torch.Generator().manual_seed(8)
    B = 3
    x = torch.randn(B, 3, 256, 256, generator=g).cuda()
    cond = (torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).cuda()
    noise = torch.randn(B, 3, 256, 256, generator=g).cuda()
    label = torch.tensor([1]).cuda()
    steps = torch.linspace(1., 0., 251)
    outs = {}
    for knob in ("1", "0"):
        monkeypatch.setenv("SRGD_CFG_SHARE", knob)
        outs[knob] = diff.p_sample(x, steps[40], cond, label, 1.0, 3.0, steps[41], noise=noise)
        n_launch = diff.last_step_launches
        outs[knob + "n"] = n_launch
    assert torch.equal(outs["1"][0], outs["0"][0]) and torch.equal(outs["1"][1], outs["0"][1])
    assert outs["1n"] == outs["0n"] + 2          # the final block's two convs over [x, r] run as two launches each
    # without a condition (n_cond_rows == 0) the halves share as well; LR-condition guidance (null rows drop the
    # condition) must NOT share
    for cs, ccs, c in ((1.0, 3.0, None), (2.0, 1.0, cond)):
        res = {}
        for knob in ("1", "0"):
            monkeypatch.setenv("SRGD_CFG_SHARE", knob)
            res[knob] = diff.p_sample(x, steps[40], c, label, cs, ccs, steps[41], noise=noise)[0]
        assert torch.equal(res["1"], res["0"])
