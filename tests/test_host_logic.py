"""CPU-only tests: host logic, the C-ABI surface, the drop-in module names.  No compute calls."""
import ctypes as C
import os
import re

import pytest
import torch

from srgd_b200 import _lib, arch, tiling
from oracle import srgd_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "srgd_b200.h")).read()
    declared = set(re.findall(r"\b(srgd_[a-z0-9_]+)\s*\(", header))
    declared -= {"srgd_b200"}
    lib = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/srgd_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().srgd_version() == 100


def test_struct_layouts_match_header(tmp_path):
    """ctypes mirrors vs the real header: sizeof/offsetof printed by a gcc-compiled probe."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    probe = tmp_path / "probe.c"
    probe.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "srgd_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu ", sizeof(srgd_step_scalars), sizeof(srgd_conv_src), sizeof(srgd_conv_phase),
         sizeof(srgd_conv_desc), sizeof(srgd_unet_config));
  printf("%zu %zu %zu %zu %zu %zu ", offsetof(srgd_conv_desc, srcs), offsetof(srgd_conv_desc, phases),
         offsetof(srgd_conv_desc, weight), offsetof(srgd_conv_desc, act), offsetof(srgd_conv_desc, gn_partials),
         offsetof(srgd_unet_config, heads));
  printf("%zu %zu %zu %zu %zu\\n", sizeof(srgd_edm_scalars), sizeof(srgd_gauss_scalars),
         offsetof(srgd_gauss_scalars, guidance_scale), offsetof(srgd_gauss_scalars, c),
         offsetof(srgd_unet_config, fixed_sinusoidal));
  return 0;
}''')
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    D, U = _lib.ConvDesc, _lib.UnetConfig
    want = [C.sizeof(_lib.StepScalars), C.sizeof(_lib.ConvSrc), C.sizeof(_lib.ConvPhase), C.sizeof(D), C.sizeof(U),
            D.srcs.offset, D.phases.offset, D.weight.offset, D.act.offset, D.gn_partials.offset, U.heads.offset,
            C.sizeof(_lib.EdmScalars), C.sizeof(_lib.GaussScalars), _lib.GaussScalars.guidance_scale.offset,
            _lib.GaussScalars.c.offset, U.fixed_sinusoidal.offset]
    assert got == want


def test_no_cpu_fallback():
    """Without a GPU every compute entry point fails loudly."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    s = _lib.StepScalars(1, 1, 1, 0.5, 0, 1, 1)
    buf = (C.c_float * 8)()
    rc = lib.srgd_sampler_step(buf, buf, None, None, buf, None, 8, C.byref(s), None)
    assert rc == -2 and "no CPU fallback" in _lib.last_error()
    import model as M
    unet = M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        unet(torch.zeros(1, 3, 64, 64), torch.zeros(1))


def test_product_package_never_imports_oracle():
    for base in ("srgd_b200", "."):
        d = os.path.join(ROOT, base)
        for f in os.listdir(d):
            if f.endswith(".py") and f not in ("bench.py", "__graft_entry__.py"):
                src = open(os.path.join(d, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{base}/{f} mentions the oracle"


def test_state_dict_layout_matches_reference_checkpoint():
    import model as M
    spec = O.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=256)
    want = O.param_shapes(spec)                       # pinned to the reference by make_golden.py
    got = diff.state_dict()
    assert list(got.keys()) == list(want.keys())
    assert all(tuple(got[k].shape) == v for k, v in want.items())
    sd = O.make_state_dict(O.UnetSpec(dim=64), 3)
    small = M.ConditionalContinuousTimeGaussianDiffusionSR(
        model=M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3),
        image_size=64)
    small.load_state_dict(sd, strict=True)
    bad = dict(sd)
    bad.pop("model.final_conv.bias")
    with pytest.raises(RuntimeError):
        small.load_state_dict(bad, strict=True)


def test_packing_contract_names():
    from srgd_b200 import weights
    names = weights.param_names(arch.UnetSpec())
    assert names[0] == "init.w" and names[-1] == "final.b" and "class.table" in names
    assert len(names) == len(set(names))
    assert "ups.0.0.res.w" in names and "downs.0.0.res.w" not in names


def test_get_model_and_checkpoint_roundtrip(tmp_path):
    import config
    import logging
    import model as M
    yaml_path = tmp_path / "c.yaml"
    yaml_path.write_text("model: conditional_continuous\nunet_dim: 64\nimage_size: 64\nnum_sample_steps: 250\n"
                         "learned_sinusoidal_cond: true\nlearned_sinusoidal_dim: 32\nlr: 1e-4\n")
    conf = config.load_config(str(yaml_path))
    assert conf.lr == "1e-4" and conf.num_classes == 3        # PyYAML 1.1 float quirk kept
    sd = O.make_state_dict(O.UnetSpec(dim=64), 5)
    ck = tmp_path / "w.pth"
    torch.save({"ema_model": sd}, ck)
    conf.ckpt_path = str(ck)
    ema = M.get_model(conf, logging.getLogger("t"))
    got = ema.module.state_dict()
    assert all(torch.equal(got[k], v) for k, v in sd.items())
    assert ema.module.model.downsample_factor == 8 and ema.module.image_size == 64
    with pytest.raises(TypeError):
        config.Config(not_a_field=1)
    conf.model = "elucidated"
    with pytest.raises(NotImplementedError):
        M.get_model(conf, logging.getLogger("t"))
    with pytest.raises(ValueError):
        M.ConditionalContinuousTimeGaussianDiffusionSR(model=ema.module.model, image_size=64, noise_schedule="nope")


def test_shipped_yaml_loads():
    ref_yaml = "/root/reference/conf/conditional_continuous_linear_df8kost_dim128.yaml"
    if not os.path.exists(ref_yaml):
        pytest.skip("reference checkout not present on this box")
    import config
    conf = config.load_config(ref_yaml)
    assert (conf.model, conf.unet_dim, conf.image_size, conf.num_sample_steps) == ("conditional_continuous", 128, 256, 250)


def test_step_scalars_match_oracle():
    import model as M
    unet = M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=64)
    steps = torch.linspace(1., 0., 251)
    for i in (0, 1, 77, 248, 249):
        s = diff.step_scalars(steps[i], steps[i + 1])
        o = O.step_scalars(steps[i], steps[i + 1])
        assert s.alpha == float(o["alpha"]) and s.sigma == float(o["sigma"]) and s.c == float(o["c"])
        assert s.alpha_next == float(o["alpha_next"])
        assert s.noise_scale == (float(o["var"].sqrt()) if i < 249 else 0.0)


def test_tile_plan_counts():
    p = tiling.TilePlan(2048, 2048)
    assert (p.canvas_h, p.canvas_w, len(p.grids[0]), len(p.grids[1])) == (2304, 2304, 81, 64)
    assert p.tiles_per_image(250) == 18125                     # BASELINE.md work table
    assert tiling.TilePlan(512, 512).tiles_per_image(250) == 1625
    assert tiling.TilePlan(256, 256).tiles_per_image(250) == 250


def test_inference_cli_flags():
    import inference
    a = inference.parse_args(["-c", "x.yaml", "-m", "w.pth", "--input_dir", "i", "--output_dir", "o"])
    assert (a.batch_size, a.num_sample_steps, a.interpolation, a.cond_scale, a.class_cond_scale) == (8, 250, "bicubic", 1.0, 1.0)
    assert (a.guidance_start_steps, a.class_guidance_start_steps, a.generation_start_steps) == (0, 0, 0)
    assert (a.start_index, a.end_index, a.test_label, a.seed, a.backend, a.amp, a.use_dpmpp_solver) == \
        (0, None, None, 71, "ddp", True, True)
    a = inference.parse_args(["-c", "x", "-m", "w", "--input_dir", "i", "--output_dir", "o", "--class_cond_scale", "3.0",
                              "--test_label", "2", "--seed", "5", "--no_amp"])
    assert a.class_cond_scale == 3.0 and a.test_label == 2 and a.seed == 5 and a.amp is False


def test_inference_directory_loop_groups_equal_sizes(tmp_path, capsys):
    """The CLI's directory loop (reference inference.py:108-142) with --images_per_batch: consecutive equally sized
    inputs are sampled together, outputs are named <name>_out.png at 4x, existing outputs and unreadable files are
    skipped -- checked on the CPU with a stand-in for the sampler."""
    import numpy as np
    from PIL import Image
    import inference

    class Stub:
        device = torch.device("cpu")
        calls = []

        def tiled_sample(self, batch_size, condition_x, **kw):
            Stub.calls.append(tuple(condition_x.shape))
            return torch.zeros_like(condition_x)

    in_dir, out_dir = tmp_path / "in", tmp_path / "out"
    in_dir.mkdir()
    out_dir.mkdir()
    rs = np.random.RandomState(0)
    for name, (w, h) in {"a.png": (10, 8), "b.png": (10, 8), "c.png": (10, 8), "d.png": (12, 8), "e.png": (10, 8)}.items():
        Image.fromarray(rs.randint(0, 256, (h, w, 3), dtype=np.uint8), mode="RGB").save(in_dir / name)
    (in_dir / "bad.png").write_bytes(b"nope")
    Image.new("RGB", (40, 32)).save(out_dir / "e_out.png")                      # already done -> skipped
    inference.batch_sr_target_images(str(in_dir), str(out_dir), Stub(), num_sample_steps=2, images_per_batch=2)
    out = capsys.readouterr().out
    assert "skip" in out and "Invalid image" in out
    # a,b together; c alone (batch full); d alone (different size)
    assert Stub.calls == [(2, 3, 32, 40), (1, 3, 32, 40), (1, 3, 32, 48)]
    assert sorted(os.listdir(out_dir)) == ["a_out.png", "b_out.png", "c_out.png", "d_out.png", "e_out.png"]
    assert Image.open(out_dir / "d_out.png").size == (48, 32)


def test_nested_load_and_inplace_edit_invalidate_packed_weights():
    """`diffusion.load_state_dict(ckpt['ema_model'])` -- the documented path -- reaches the U-Net
    only through _load_from_state_dict, never through its load_state_dict override; the packed device weights must
    still be dropped, and an in-place edit of any parameter must change the pack key."""
    import model as M
    unet = M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=64)
    drops = []
    orig = unet._drop_handle
    unet._drop_handle = lambda: (drops.append(1), orig())[1]
    diff.load_state_dict(diff.state_dict(), strict=True)
    assert len(drops) == 1
    k0 = unet._weights_key("dev")
    with torch.no_grad():
        unet.final_conv.bias.add_(1.0)
    assert unet._weights_key("dev") != k0


def test_out_of_range_class_label_raises_like_nn_embedding():
    """the reference's nn.Embedding raises IndexError for label >= num_classes (model.py:612, 693)."""
    import model as M
    unet = M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    assert unet.labels_for(torch.tensor([2]), 4, "cpu").tolist() == [2, 2, 2, 2]
    for bad in (3, -1):
        with pytest.raises(IndexError):
            unet.labels_for(torch.tensor([bad]), 1, "cpu")


@pytest.mark.parametrize("init", ["unit", "torch"])
def test_seeded_state_dict_equals_the_oracles(init):
    """bench.py / smoke load `arch.seeded_state_dict` (product side, no oracle import); the parity tests load the
    oracle's `make_state_dict`: the two must be the same tensors so that parity is shown on the benchmarked weights."""
    spec_p, spec_o = arch.UnetSpec(dim=64), O.UnetSpec(dim=64)
    a, b = arch.seeded_state_dict(spec_p, 1234, init=init), O.make_state_dict(spec_o, 1234, init=init)
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_ingest_cache_validation_and_deferred_fp32(tmp_path):
    """<ckpt>.srgd_b200_pack is keyed by the checkpoint's sha256 (size / mtime as the fast check);
    a start from the cache leaves the fp32 parameters as placeholders until somebody asks for state_dict()."""
    import logging
    import config as Cfg
    import model as M
    from srgd_b200 import weights
    spec = arch.UnetSpec(dim=64)
    sd = arch.seeded_state_dict(spec, 5, init="torch")
    ckpt = str(tmp_path / "m.pth")
    torch.save({"ema_model": sd}, ckpt)
    fake_pack = {"init.w": torch.arange(12.).reshape(3, 4).bfloat16(), "init.b": torch.ones(3)}
    assert weights.load_pack_cache(ckpt, spec) is None
    assert weights.save_pack_cache(ckpt, spec, fake_pack) == weights.pack_cache_path(ckpt)
    got = weights.load_pack_cache(ckpt, spec)
    assert got is not None and all(torch.equal(got[k], v) for k, v in fake_pack.items())
    assert weights.load_pack_cache(ckpt, arch.UnetSpec(dim=128)) is None            # other architecture
    os.utime(ckpt, ns=(1, 1))                                                        # touched, same bytes: sha decides
    assert weights.load_pack_cache(ckpt, spec) is not None
    sd2 = dict(sd)
    sd2["model.final_conv.bias"] = sd["model.final_conv.bias"] + 1                   # same size, other contents
    torch.save({"ema_model": sd2}, ckpt)
    assert weights.load_pack_cache(ckpt, spec) is None
    # get_model from the cache: no torch.load of the fp32 checkpoint until state_dict() is requested
    weights.save_pack_cache(ckpt, spec, fake_pack)
    yaml_path = tmp_path / "c.yaml"
    yaml_path.write_text("model: conditional_continuous\nunet_dim: 64\nimage_size: 64\nnum_sample_steps: 250\n"
                         "learned_sinusoidal_cond: true\nlearned_sinusoidal_dim: 32\n")
    conf = Cfg.load_config(str(yaml_path))
    conf.ckpt_path = ckpt
    ema = M.get_model(conf, logging.getLogger("t"))
    unet = ema.module.model
    assert unet._deferred_ckpt == ckpt and unet._cached_pack is not None
    ema.module.to("cpu")                                                             # placeholder-aware _apply
    full = ema.module.state_dict()
    assert unet._deferred_ckpt is None and all(torch.equal(full[k], sd2[k]) for k in sd2)
    # loading other weights detaches the cache
    ema.module.load_state_dict(sd, strict=True)
    assert unet._cached_pack is None


def test_get_model_builds_the_edm_family(tmp_path):
    """conf.model == 'conditional_elucidated' (model.py:3593-3614): same U-Net under the key prefix `net.`, the
    reference's constructor arguments, checkpoint round trip."""
    import logging
    import config as Cfg
    import model as M
    yaml_path = tmp_path / "c.yaml"
    yaml_path.write_text("model: conditional_elucidated\nunet_dim: 64\nimage_size: 64\nnum_sample_steps: 32\n"
                         "learned_sinusoidal_cond: true\nlearned_sinusoidal_dim: 32\nsigma_max: 60\n")
    conf = Cfg.load_config(str(yaml_path))
    sd = O.make_state_dict(O.UnetSpec(dim=64), 5, prefix="net.")
    ck = tmp_path / "w.pth"
    torch.save({"ema_model": sd}, ck)
    conf.ckpt_path = str(ck)
    edm = M.get_model(conf, logging.getLogger("t")).module
    assert isinstance(edm, M.ConditionalElucidatedDiffusionSR) and edm.sigma_max == 60 and edm.use_dpmpp_solver is True
    got = edm.state_dict()
    assert list(got) == list(sd) and all(torch.equal(got[k], v) for k, v in sd.items())
    s = edm.sample_schedule(8)
    assert s.shape == (9,) and abs(float(s[0]) - 60.0) < 1e-4 and float(s[-1]) == 0.0 and abs(float(s[-2]) - 0.002) < 1e-6
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        edm.sample(batch_size=1, condition_x=torch.rand(1, 3, 64, 64), num_sample_steps=4)


def test_get_model_builds_the_discrete_time_family(tmp_path):
    """conf.model == 'conditional_gaussian' (model.py:3552-3570): the U-Net with the fixed SinusoidalPosEmb (no
    `time_mlp.0.weights`, a dim-wide first time-MLP layer), the reference's 13 registered buffers in front of the
    `model.` keys, strict checkpoint round trip, and the per-step scalars of DDPM / DDIM against the oracle's tables."""
    import logging
    import config as Cfg
    import model as M
    yaml_path = tmp_path / "c.yaml"
    yaml_path.write_text("model: conditional_gaussian\nunet_dim: 64\nimage_size: 64\nlearned_sinusoidal_cond: false\n"
                         "timesteps: 1000\nsampling_timesteps: 10\nobjective: pred_v\nbeta_schedule: sigmoid\n")
    conf = Cfg.load_config(str(yaml_path))
    spec = O.UnetSpec(dim=64, learned_sinusoidal_cond=False)
    p = O.GaussParams(1000, 10, "pred_v", "sigmoid")
    tab = O.gauss_tables(p)
    sd = dict(tab)
    sd["log_one_minus_alphas_cumprod"] = torch.log(1. - torch.cumprod(1. - O.gauss_betas("sigmoid", 1000), 0)).float()
    snr = torch.cumprod(1. - O.gauss_betas("sigmoid", 1000), 0)
    snr = snr / (1 - snr)
    sd["loss_weight"] = (snr / (snr + 1)).float()
    sd.update(O.make_state_dict(spec, 5, prefix="model."))
    ck = tmp_path / "w.pth"
    torch.save({"ema_model": sd}, ck)
    conf.ckpt_path = str(ck)
    m = M.get_model(conf, logging.getLogger("t")).module
    assert isinstance(m, M.ConditionalGaussianDiffusionSR) and m.is_ddim_sampling and m.num_timesteps == 1000
    assert not m.model.random_or_learned_sinusoidal_cond
    got = m.state_dict()
    assert set(got) == set(sd) and list(got)[:3] == ["betas", "alphas_cumprod", "alphas_cumprod_prev"]
    assert all(torch.equal(got[k], v) for k, v in sd.items())
    assert "model.time_mlp.0.weights" not in got and got["model.time_mlp.1.weight"].shape == (256, 64)
    # DDIM scalars (model.py:1608-1612) with the oracle's arithmetic
    m.ddim_sampling_eta = 0.7
    s = m._scalars(600, _lib.GAUSS_DDIM, True, True, 2.0, 300)
    a, an = tab["alphas_cumprod"][600], tab["alphas_cumprod"][300]
    sigma = 0.7 * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    assert s.noise_scale == float(sigma) and s.c == float((1 - an - sigma ** 2).sqrt()) and s.sqrt_ac_next == float(an.sqrt())
    assert s.objective == _lib.OBJ_PRED_V and s.sqrt_ac == float(tab["sqrt_alphas_cumprod"][600])
    s = m._scalars(17, _lib.GAUSS_DDPM, True, False, 1.0)
    assert s.noise_scale == float((0.5 * tab["posterior_log_variance_clipped"][17]).exp())
    assert s.coef1 == float(tab["posterior_mean_coef1"][17]) and s.coef2 == float(tab["posterior_mean_coef2"][17])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample(batch_size=1, condition_x=torch.rand(1, 3, 64, 64))
    with pytest.raises(ValueError, match="unknown beta schedule"):
        M.ConditionalGaussianDiffusionSR(m.model, image_size=64, beta_schedule="quadratic")


def test_unet_shape_errors_come_before_any_device_work():
    """Mismatched inputs fail in the reference inside torch.cat / the init conv / the divisibility assert
    (model.py:679-686); here they must be caught on the host, because the C side reads raw pointers."""
    import model as M
    unet = M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    x, t = torch.zeros(2, 3, 64, 64), torch.zeros(2)
    with pytest.raises(RuntimeError, match="must match"):
        unet(x, t, None, torch.zeros(2, 3, 32, 32))
    with pytest.raises(RuntimeError, match="must match"):
        unet(x, t, None, torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="expected x of shape"):
        unet(torch.zeros(2, 1, 64, 64), t)
    with pytest.raises(AssertionError, match="divisible by 8"):
        unet(torch.zeros(2, 3, 60, 64), t)
    with pytest.raises(RuntimeError, match="for a batch of"):
        unet(x, torch.zeros(3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        unet(x, t)
