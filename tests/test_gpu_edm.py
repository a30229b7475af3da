"""GPU parity of the EDM sampler family (srgd_b200/edm.py; reference ConditionalElucidatedDiffusionSR,
model.py:2059-2560) against the fp32 oracle (oracle/srgd_oracle.py edm_*, pinned bit-exactly
to the unmodified reference class by tests/golden/edm_tiny.npz; its pip base class -- coefficients, sigma schedule -- is
a restatement: parity unpinned, like `Attend`).  Run with `pytest -m gpu` on a B200.

The oracle runs on the same GPU in strict fp32 and draws from torch's CUDA generator with the reference's shapes and
order, the stream the product path consumes after the same torch.manual_seed."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import gpu_util as G  # noqa: E402
from oracle import srgd_oracle as O  # noqa: E402  (checker only)
import model as M  # noqa: E402
from test_gpu_unet import _oracle_on_gpu  # noqa: E402

_cache = {}


def build(dim=64, image_size=64, steps=32, dpmpp=False):
    key = (dim, image_size, steps, dpmpp)
    if key not in _cache:
        spec = O.UnetSpec(dim=dim)
        sd = O.make_state_dict(spec, 22, prefix="net.", init="torch")
        unet = M.ConditionalSRUnet(dim=dim, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
        edm = M.ConditionalElucidatedDiffusionSR(unet, image_size=image_size, num_sample_steps=steps,
                                                 use_dpmpp_solver=dpmpp)
        edm.load_state_dict(sd, strict=True)
        edm = edm.eval().to("cuda")
        edm.progress = False
        _cache[key] = (edm, _oracle_on_gpu(sd), spec)
    return _cache[key]


def test_preconditioned_forward_vs_oracle():
    """c_skip x + c_out net with each guidance kind at small / medium / large sigma; the error of the denoised image is
    the bf16 network error scaled by c_out <= sigma_data."""
    edm, gsd, spec = build()
    p = O.EdmParams()
    g = torch.Generator().manual_seed(3)
    cond = (torch.rand(3, 3, 64, 64, generator=g) * 2 - 1).cuda()
    label = torch.tensor([2]).cuda()
    for sigma, cs, ccs in ((0.05, 1.0, 1.0), (1.3, 1.0, 2.5), (30.0, 1.8, 1.0)):
        x = (torch.randn(3, 3, 64, 64, generator=g) * math.hypot(sigma, 0.5)).cuda()
        got = edm.preconditioned_network_forward(x, sigma, cond, label, cs, ccs, clamp=True)
        with torch.inference_mode():
            ref = O.edm_denoise(gsd, spec, p, x, sigma, cond, label, cs, ccs, clamp=True)
        err = float((got - ref).abs().max())
        print(f"sigma {sigma} cs {cs} ccs {ccs}: denoised max-abs {err:.5f}")
        assert err <= 3e-2
    with pytest.raises(NotImplementedError):
        edm.preconditioned_network_forward(x, 1.0, cond, label, 2.0, 2.0)


@pytest.mark.parametrize("dpmpp", [False, True])
def test_sample_free_running_vs_oracle(dpmpp):
    """sample(): stochastic Heun (sample_org) and DPM-Solver++ 2M, 32 steps, class guidance 2.0, B = 2, seed 71."""
    edm, gsd, spec = build(dpmpp=dpmpp)
    p = O.EdmParams()
    g = torch.Generator().manual_seed(4)
    cond01 = torch.rand(2, 3, 64, 64, generator=g).cuda()
    label = torch.tensor([1]).cuda()
    torch.manual_seed(71)
    img = edm.sample(batch_size=2, condition_x=cond01, class_label=label, class_cond_scale=2.0, num_sample_steps=32)
    torch.manual_seed(71)
    fn = O.edm_sample_dpmpp if dpmpp else O.edm_sample_heun
    with torch.inference_mode():
        ref = fn(gsd, spec, p, 2, cond01, label, class_cond_scale=2.0, num_sample_steps=32)
    psnr = G.psnr(img.cpu(), ref.cpu())
    print(f"EDM {'dpm++ 2M' if dpmpp else 'Heun'} 32 steps: PSNR {psnr:.2f} dB, max-abs {float((img - ref).abs().max()):.4f}")
    assert img.shape == (2, 3, 64, 64) and float(img.min()) >= 0 and float(img.max()) <= 1
    assert psnr >= 45.0


def test_sample_options_vs_oracle():
    """generation_start_steps (start from the noised condition, one image: model.py:2192-2194), guidance_start_steps,
    LR-condition guidance; with_images / with_x0_images list lengths."""
    edm, gsd, spec = build()
    p = O.EdmParams()
    g = torch.Generator().manual_seed(6)
    cond01 = torch.rand(1, 3, 64, 64, generator=g).cuda()
    label = torch.tensor([0]).cuda()
    kw = dict(cond_scale=1.5, guidance_start_steps=8, generation_start_steps=4, num_sample_steps=32)
    torch.manual_seed(71)
    img, frames, x0s = edm.sample(batch_size=1, condition_x=cond01, class_label=label, with_images=True,
                                  with_x0_images=True, **kw)
    torch.manual_seed(71)
    with torch.inference_mode():
        ref = O.edm_sample_heun(gsd, spec, p, 1, cond01, label, **kw)
    assert len(frames) == 1 + 28 and len(x0s) == 1 + 28
    assert G.psnr(img.cpu(), ref.cpu()) >= 45.0


def test_tiled_sample_vs_oracle():
    """tiled_sample (Heun): 272x264 HR -> 768x768 canvas, 9 / 4 alternating tiles of 256, odd-step re-noise at sigma_i,
    batch_size 4, class guidance 2.0, 16 steps."""
    edm, gsd, spec = build(image_size=256, steps=16)
    p = O.EdmParams()
    g = torch.Generator().manual_seed(9)
    cond01 = torch.rand(1, 3, 272, 264, generator=g).cuda()
    label = torch.tensor([0]).cuda()
    torch.manual_seed(71)
    img = edm.tiled_sample(batch_size=4, condition_x=cond01, class_label=label, class_cond_scale=2.0, num_sample_steps=16)
    torch.manual_seed(71)
    with torch.inference_mode():
        ref = O.edm_tiled_sample(gsd, spec, p, 4, cond01, label, class_cond_scale=2.0, num_sample_steps=16)
    psnr = G.psnr(img.cpu(), ref.cpu())
    print(f"EDM tiled_sample 16 steps: PSNR {psnr:.2f} dB, max-abs {float((img - ref).abs().max()):.4f}")
    assert img.shape == (1, 3, 272, 264) and psnr >= 45.0


def test_edm_kernels_are_exact():
    """srgd_edm_perturb / srgd_edm_update / srgd_edm_dpmpp against the reference's fp32 op sequence evaluated by torch
    on the CPU (where `tensor / python_float` is a true division; torch's CUDA kernel multiplies by the reciprocal
    instead, so the reference itself differs between devices in the last bit): bit-exact."""
    import ctypes as C
    from srgd_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    n = 3 * 64 * 64 + 5
    host = [torch.randn(n, generator=g) for _ in range(5)]
    images, noise, net_c, net_n, dprev = (t.cuda() for t in host)
    h_images, h_noise, h_net_c, h_net_n, h_dprev = host
    st = _lib.current_stream()
    hat, xin = torch.empty_like(images), torch.empty_like(images)
    _lib.check(lib.srgd_edm_perturb(_lib.ptr(images), _lib.ptr(noise), 1.003, 0.7, 1.9, _lib.ptr(hat), _lib.ptr(xin), n, st))
    ref_hat = h_images + 0.7 * (1.003 * h_noise)
    assert torch.equal(hat.cpu(), ref_hat) and torch.equal(xin.cpu(), 1.9 * ref_hat)
    s = _lib.EdmScalars(0.31, 0.44, 2.5, 1.7, -0.6, 0.8, 1)
    out, d, den, xin2 = (torch.empty_like(images) for _ in range(4))
    _lib.check(lib.srgd_edm_update(_lib.ptr(hat), _lib.ptr(net_c), _lib.ptr(net_n), _lib.ptr(images), _lib.ptr(dprev),
                                   _lib.ptr(out), _lib.ptr(d), _lib.ptr(den), _lib.ptr(xin2), n, C.byref(s), st))
    f = lambda v: torch.tensor(v, dtype=torch.float32).item()
    o = f(0.31) * ref_hat + f(0.44) * h_net_c
    nul = f(0.31) * ref_hat + f(0.44) * h_net_n
    ref_den = (nul + (o - nul) * f(2.5)).clamp(-1., 1.)
    ref_d = (ref_hat - ref_den) / f(1.7)
    ref_out = h_images + f(-0.6) * (h_dprev + ref_d)
    assert torch.equal(den.cpu(), ref_den) and torch.equal(d.cpu(), ref_d) and torch.equal(out.cpu(), ref_out)
    assert torch.equal(xin2.cpu(), f(0.8) * ref_out)
    nxt = torch.empty_like(images)
    _lib.check(lib.srgd_edm_dpmpp(_lib.ptr(images), _lib.ptr(den), _lib.ptr(d), 0.9, -0.2, 1.5, -0.5, 0.0, _lib.ptr(nxt),
                                  None, n, st))
    assert torch.equal(nxt.cpu(), f(0.9) * h_images - f(-0.2) * (f(1.5) * ref_den + f(-0.5) * ref_d))


def test_cli_with_the_edm_family(tmp_path, capsys):
    """inference.py end to end with conf.model = conditional_elucidated: `<name>_out.png` at 4x, pixels equal to a hand
    call of the same chain (get_model -> tiled_sample) with the same seed."""
    import logging
    import numpy as np
    from PIL import Image
    import config
    import inference
    yaml_path = tmp_path / "c.yaml"
    yaml_path.write_text("model: conditional_elucidated\nunet_dim: 64\nimage_size: 256\nnum_sample_steps: 8\n"
                         "learned_sinusoidal_cond: true\nlearned_sinusoidal_dim: 32\nuse_dpmpp_solver: false\n")
    ckpt = tmp_path / "w.pth"
    torch.save({"ema_model": O.make_state_dict(O.UnetSpec(dim=64), 22, prefix="net.", init="torch")}, ckpt)
    in_dir, out_dir = tmp_path / "in", tmp_path / "out"
    in_dir.mkdir()
    lr = np.random.RandomState(5).randint(0, 256, (66, 70, 3), dtype=np.uint8)
    Image.fromarray(lr, mode="RGB").save(in_dir / "a.png")
    argv = ["-c", str(yaml_path), "-m", str(ckpt), "--input_dir", str(in_dir), "--output_dir", str(out_dir),
            "--num_sample_steps", "8", "--test_label", "1", "--class_cond_scale", "2.0", "--seed", "71", "--batch_size", "4"]
    inference.main(argv)
    capsys.readouterr()
    got = Image.open(out_dir / "a_out.png")
    assert got.size == (280, 264)
    conf = config.load_config(str(yaml_path))
    conf.num_sample_steps, conf.ckpt_path = 8, str(ckpt)
    sr = M.get_model(conf, logging.getLogger("t")).module.eval().to("cuda")
    sr.progress = False
    cond = inference._to_unit_tensor(Image.fromarray(lr, mode="RGB").resize((280, 264), resample=Image.BICUBIC)).cuda()
    inference.seed_everything(71)
    ref = sr.tiled_sample(batch_size=4, condition_x=cond, class_label=torch.tensor([1], device="cuda"),
                          class_cond_scale=2.0, num_sample_steps=8)
    assert np.array_equal(np.asarray(inference._to_image(ref[0])), np.asarray(got))
