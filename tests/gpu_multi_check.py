"""Multi-GPU consistency check (not a pytest test; run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_multi_check.py

  1. tiled_sample(shard_tiles=True) (exact mode): contiguous tile ranges per rank + one all-gather per step must give
     every rank the image the same mode produces on ONE GPU, bit for bit (dim-128 U-Net, 768^2 and 2304^2 canvases);
     the difference to the reference's minibatch partition (fp32 re-association only) is printed;
  2. sample_sharded: batch rows split over the ranks + one final gather == the single-GPU sample() with replicated RNG.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import srgd_oracle as O  # noqa: E402  (deterministic random-init weights only)
import model as M  # noqa: E402
from srgd_b200 import sharding  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    solo = [dist.new_group([r]) for r in range(world)][rank]                 # a group of this rank alone = "1 GPU"
    spec = O.UnetSpec()
    unet = M.ConditionalSRUnet(dim=128, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=256, num_sample_steps=6)
    diff.load_state_dict(O.make_state_dict(spec, 1234, init="torch"), strict=True)
    diff = diff.eval().to(dev)
    diff.progress = False
    g = torch.Generator().manual_seed(5)
    label = torch.tensor([2], device=dev)
    flags_all = []
    # 768^2 canvas (9 / 4 tiles per step; full 250-step schedule, so that the distance to the reference's minibatch
    # partition is a meaningful PSNR) and config 4's 2304^2 canvas (81 / 64 tiles per step; 6 steps), full-width U-Net
    for hw, bs, nsteps in (((288, 320), 2, 250), ((2048, 2048), 8, 6)):
        cond01 = torch.rand(1, 3, hw[0] // 4, hw[1] // 4, generator=g)
        cond01 = torch.nn.functional.interpolate(cond01, size=hw, mode="bicubic", align_corners=False).clamp(0, 1).to(dev)
        outs = {}
        for mode, kw in (("reference partition", dict(shard_tiles=False)),
                         ("exact mode, this rank alone", dict(shard_tiles=True, shard_group=solo)),
                         (f"exact mode, {world} ranks", dict(shard_tiles=True))):
            torch.manual_seed(71)
            torch.cuda.manual_seed(71)
            outs[mode] = diff.tiled_sample(batch_size=bs, condition_x=cond01, class_label=label, class_cond_scale=3.0,
                                           num_sample_steps=nsteps, **kw)
        a, b, c = outs.values()
        same = bool(torch.equal(b, c))
        flags = [None] * world
        dist.all_gather_object(flags, same)
        flags_all += flags
        if rank == 0:
            print(f"{hw[0]}x{hw[1]} HR, world {world}: exact mode on {world} ranks == exact mode on 1 GPU, per rank: {flags}; "
                  f"exact mode vs the reference's minibatch partition ({nsteps} steps): max-abs "
                  f"{float((a - b).abs().max()):.2e}, PSNR {float(-10 * torch.log10(((a - b) ** 2).mean())):.1f} dB", flush=True)
    # batch sharding: rows split over the ranks + one final gather.  In batch-invariant mode the rows do not depend on
    # the batch they are computed in, so the sharded result equals the single-GPU result bit for bit.
    from srgd_b200 import _lib
    lib = _lib.load()
    prev = lib.srgd_set_batch_invariant(1)
    cond_b = torch.rand(5, 3, 64, 64, generator=g).to(dev)
    diff64 = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=64, num_sample_steps=6)
    diff64.progress = False
    full = sharding.sample_sharded(diff64, cond_b, class_label=torch.tensor([1], device=dev), class_cond_scale=3.0,
                                   num_sample_steps=6, seed=9)
    ok = True
    if rank == 0:
        gen = torch.Generator(device=dev)
        gen.manual_seed(9)
        img = torch.randn(5, 3, 64, 64, generator=gen, device=dev)
        steps = torch.linspace(1., 0., 7)
        for i in range(6):
            noise = None if i == 5 else torch.randn(5, 3, 64, 64, generator=gen, device=dev)
            img, _ = diff64.p_sample(img, steps[i], cond_b * 2 - 1, torch.tensor([1], device=dev), 1.0, 3.0, steps[i + 1],
                                     noise=noise)
        ref = diff64._finalize(img)
        ok = bool(torch.equal(full, ref))
        print(f"multi-GPU check, world {world}: tiled exact mode == 1 GPU on every rank: {flags_all}; "
              f"sample_sharded == single-GPU sample (batch-invariant mode): {ok}", flush=True)
    lib.srgd_set_batch_invariant(prev)
    dist.barrier()
    dist.destroy_process_group()
    assert all(flags_all), flags_all
    assert ok


if __name__ == "__main__":
    main()
