"""`inference.py --shard_tiles` under torchrun (not a pytest test; run on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_cli_shard_check.py

Every rank walks the same directory, the tiles of every image are split over the ranks (tiled_sample's exact mode), rank 0
writes the PNGs.  Checked: the files equal, pixel for pixel, what ONE GPU produces in exact mode with the same seed."""
import logging
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import srgd_oracle as O  # noqa: E402  (deterministic random-init weights only)
import config  # noqa: E402
import inference  # noqa: E402
import model as M  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    tmp = os.path.join(tempfile.gettempdir(), "srgd_cli_shard")
    if rank == 0:
        os.makedirs(os.path.join(tmp, "in"), exist_ok=True)
        open(os.path.join(tmp, "c.yaml"), "w").write(
            "model: conditional_continuous\nnoise_schedule: linear\nunet_dim: 64\nimage_size: 256\nnum_sample_steps: 250\n"
            "learned_sinusoidal_cond: true\nlearned_sinusoidal_dim: 32\n")
        torch.save({"ema_model": O.make_state_dict(O.UnetSpec(dim=64), 22, init="torch")}, os.path.join(tmp, "w.pth"))
        rs = np.random.RandomState(3)
        for name, (w, h) in {"a.png": (70, 66), "b.png": (40, 36), "c.png": (130, 70)}.items():
            Image.fromarray(rs.randint(0, 256, (h, w, 3), dtype=np.uint8), mode="RGB").save(os.path.join(tmp, "in", name))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
    out_dir = os.path.join(tmp, f"out_w{world}")
    argv = ["-c", os.path.join(tmp, "c.yaml"), "-m", os.path.join(tmp, "w.pth"), "--input_dir", os.path.join(tmp, "in"),
            "--output_dir", out_dir, "--num_sample_steps", "8", "--test_label", "1", "--class_cond_scale", "2.0",
            "--seed", "71", "--batch_size", "4", "--shard_tiles"]
    inference.main(argv)
    dist.barrier()
    # one GPU, exact mode: a group of this rank alone
    solo = [dist.new_group([r]) for r in range(world)][rank]
    conf = config.load_config(os.path.join(tmp, "c.yaml"))
    conf.num_sample_steps, conf.ckpt_path = 8, os.path.join(tmp, "w.pth")
    sr = M.get_model(conf, logging.getLogger("t")).module.eval().to(torch.device("cuda", local))
    sr.progress = False
    ok = True
    if rank == 0:
        for name in ("a.png", "b.png", "c.png"):
            lr = Image.open(os.path.join(tmp, "in", name)).convert("RGB")
            cond = inference._to_unit_tensor(lr.resize((lr.size[0] * 4, lr.size[1] * 4), resample=Image.BICUBIC)).cuda()
            inference.seed_everything(71)
            with torch.inference_mode():
                ref = sr.tiled_sample(batch_size=4, condition_x=cond, class_label=torch.tensor([1], device="cuda"),
                                      class_cond_scale=2.0, num_sample_steps=8, shard_tiles=True, shard_group=solo)
            got = np.asarray(Image.open(os.path.join(out_dir, name.replace(".png", "_out.png"))))
            same = np.array_equal(np.asarray(inference._to_image(ref[0])), got)
            print(f"{name}: {got.shape[1]}x{got.shape[0]} written by rank 0 of {world}; equals the one-GPU exact-mode image: {same}",
                  flush=True)
            ok = ok and same
    dist.barrier()
    dist.destroy_process_group()
    assert ok


if __name__ == "__main__":
    main()
