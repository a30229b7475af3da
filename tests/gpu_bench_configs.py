"""Throughput of the other BASELINE.json configurations on one B200 (not a pytest test, not the bench line):

    python tests/gpu_bench_configs.py            # prints one line per configuration

  configs[2]  sample(): batch 32, class_cond_scale 3.0 (cond + null-label rows = 64 U-Net rows per step)
  configs[3]  tiled_sample(): one 512x512 LR image (2048^2 HR, 2304^2 canvas, 81 / 64 tiles per step)
  configs[4]  tiled_sample(): one 128x128 LR image (512^2 HR, 768^2 canvas, 9 / 4 tiles per step)
Each is timed over a few steps of the real 250-step schedule (even + odd tile grids) with CUDA events and
extrapolated to the full schedule; inputs are synthetic (RandomState(71), PIL bicubic x4).
"""
import argparse
import os
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import srgd_oracle as O  # noqa: E402  (deterministic random-init weights only)
import model as M  # noqa: E402


def synth(lr_size, idx=0):
    lr = np.random.RandomState(71 + idx).randint(0, 256, (lr_size, lr_size, 3), dtype=np.uint8)
    hr = Image.fromarray(lr, mode="RGB").resize((4 * lr_size, 4 * lr_size), resample=Image.BICUBIC)
    return torch.from_numpy(np.array(hr, dtype=np.uint8)).permute(2, 0, 1).float().div(255.)[None]


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:                       # torchrun: the tiled configurations with the tiles of a step split over the ranks
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = O.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=256, num_sample_steps=250)
    diff.load_state_dict(O.make_state_dict(spec, 1234), strict=True)
    diff = diff.eval().to(torch.device("cuda", local))
    diff.progress = False
    n = a.steps
    with torch.inference_mode():
        # configs[2]: CFG 3.0, batch 32, each test_label
        cond = torch.cat([synth(64, i) for i in range(32)]).cuda()
        for lab in ((0, 1, 2) if world == 1 else ()):
            label = torch.tensor([lab], device="cuda")
            diff.sample(batch_size=32, condition_x=cond, class_label=label, class_cond_scale=3.0, num_sample_steps=2)
            # generation_start_steps skips the first 250-n steps: the timed steps are the last n of the real schedule
            ms = timed(lambda: diff.sample(batch_size=32, condition_x=cond, class_label=label, class_cond_scale=3.0,
                                           num_sample_steps=250, generation_start_steps=250 - n))
            per = ms / n
            print(f"configs[2] label {lab}: sample() B=32 CFG 3.0: {per:.2f} ms/step -> {32 / (per * 250e-3):.3f} images/s, "
                  f"{64 / (per * 1e-3):.0f} U-Net steps/s", flush=True)
        # configs[3] / configs[4]: tiled_sample on one large image
        for name, lr, bs in (("configs[4]", 128, 9), ("configs[3]", 512, 27)):
            c = synth(lr).cuda()
            label = torch.tensor([0], device="cuda")
            if world > 1:
                bs = max(1, -(-81 // world) if lr == 512 else -(-9 // world))      # one minibatch per rank per step
            shard = world > 1
            diff.tiled_sample(batch_size=bs, condition_x=c, class_label=label, num_sample_steps=250,
                              generation_start_steps=248, shard_tiles=shard)
            if world > 1:
                dist.barrier()
            ms = timed(lambda: diff.tiled_sample(batch_size=bs, condition_x=c, class_label=label,
                                                 num_sample_steps=250, generation_start_steps=250 - n,
                                                 shard_tiles=shard))
            if world > 1:
                t = torch.tensor([ms], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t)
            per = ms / n
            if rank == 0:
                print(f"{name}: tiled_sample() {lr}x{lr} LR on {world} GPU(s), minibatch {bs}: {per:.2f} ms/step (mean of "
                      f"even+odd grids) -> {1 / (per * 250e-3):.4f} images/s", flush=True)
        # configs[4] as a many-image workload: N 128x128-LR images advance together (tiles of different images share
        # the denoiser batches, run_tiled's N > 1 extension)
        if world == 1:
            for nimg in (4, 16):
                c = torch.cat([synth(128, k) for k in range(nimg)]).cuda()
                diff.tiled_sample(batch_size=9, condition_x=c, class_label=label, num_sample_steps=250,
                                  generation_start_steps=248)
                ms = timed(lambda: diff.tiled_sample(batch_size=9, condition_x=c, class_label=label,
                                                     num_sample_steps=250, generation_start_steps=250 - n))
                per = ms / n
                print(f"configs[4] x{nimg} images per batch: {per:.2f} ms/step -> {nimg / (per * 250e-3):.4f} images/s",
                      flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
