"""End-to-end run of the drop-in command line (reference: inference.py:21-44 flags, 59-98 per-image pipeline,
108-142 directory loop) on a B200: yaml config + `{'ema_model': state_dict}` checkpoint + a directory of PNGs in,
`<name>_out.png` at exactly 4x the size out, existing outputs skipped, unreadable files skipped -- and the pixels equal
what a direct `tiled_sample` call with the same seed produces."""
import logging
import os

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu

from oracle import srgd_oracle as O  # noqa: E402  (deterministic random-init weights only)


def test_inference_cli_end_to_end(tmp_path, capsys):
    import config
    import inference
    import model as M
    yaml_path = tmp_path / "c.yaml"
    yaml_path.write_text("model: conditional_continuous\nnoise_schedule: linear\nunet_dim: 64\nimage_size: 256\n"
                         "num_sample_steps: 250\nlearned_sinusoidal_cond: true\nlearned_sinusoidal_dim: 32\nlr: 1e-4\n")
    sd = O.make_state_dict(O.UnetSpec(dim=64), 22, init="torch")
    ckpt = tmp_path / "w.pth"
    torch.save({"ema_model": sd}, ckpt)
    in_dir, out_dir = tmp_path / "in", tmp_path / "out"
    in_dir.mkdir()
    rs = np.random.RandomState(3)
    # (w, h): one 256x256 tile / a 768x768 canvas with 9 + 4 tiles / same size as b (batched with it below)
    sizes = {"a.png": (40, 36), "b.png": (70, 66), "c.png": (70, 66)}
    for name, (w, h) in sizes.items():
        Image.fromarray(rs.randint(0, 256, (h, w, 3), dtype=np.uint8), mode="RGB").save(in_dir / name)
    (in_dir / "broken.png").write_bytes(b"not a png")
    argv = ["-c", str(yaml_path), "-m", str(ckpt), "--input_dir", str(in_dir), "--output_dir", str(out_dir),
            "--num_sample_steps", "6", "--test_label", "1", "--class_cond_scale", "2.0", "--seed", "71",
            "--batch_size", "4"]
    inference.main(argv)
    out = capsys.readouterr().out
    assert "Invalid image or unable to open image" in out
    produced = sorted(os.listdir(out_dir))
    assert produced == ["a_out.png", "b_out.png", "c_out.png"]
    for name, (w, h) in sizes.items():
        img = Image.open(out_dir / name.replace(".png", "_out.png"))
        assert img.mode == "RGB" and img.size == (4 * w, 4 * h)
    # the same call chain by hand: identical pixels
    conf = config.load_config(str(yaml_path))
    conf.num_sample_steps, conf.ckpt_path = 6, str(ckpt)
    sr = M.get_model(conf, logging.getLogger("t")).module.eval().to("cuda")
    lr = Image.open(in_dir / "b.png").convert("RGB")
    cond = inference._to_unit_tensor(lr.resize((280, 264), resample=Image.BICUBIC)).cuda()
    inference.seed_everything(71)
    with torch.inference_mode():
        ref = sr.tiled_sample(batch_size=4, condition_x=cond, class_label=torch.tensor([1], device="cuda"),
                              class_cond_scale=2.0, num_sample_steps=6)
    assert np.array_equal(np.asarray(inference._to_image(ref[0])), np.asarray(Image.open(out_dir / "b_out.png")))
    # --images_per_batch: b and c advance together (stacked denoiser batches, shared noise stream); the images are the
    # ones of the one-at-a-time run up to the bf16 kernels' batch-composition jitter, which six coarse steps on random
    # weights amplify (exact equality of the logic is pinned on the CPU: tests/test_tiled_gloo.py)
    out2 = tmp_path / "out2"
    inference.main(argv[:7] + [str(out2)] + argv[8:] + ["--images_per_batch", "2"])
    capsys.readouterr()
    for name in ("a_out.png", "b_out.png", "c_out.png"):
        one = np.asarray(Image.open(out_dir / name)).astype(np.int32)
        two = np.asarray(Image.open(out2 / name)).astype(np.int32)
        assert one.shape == two.shape
        psnr = 10 * np.log10(255.0 ** 2 / max(np.mean((one - two) ** 2.0), 1e-9))
        assert psnr >= 30.0, (name, psnr)                 # a wrong noise stream or tile mapping gives ~8 dB
    # second run: everything already there -> skipped, files untouched
    stamp = os.path.getmtime(out_dir / "a_out.png")
    inference.main(argv)
    assert capsys.readouterr().out.count("skip") >= 3 and os.path.getmtime(out_dir / "a_out.png") == stamp
