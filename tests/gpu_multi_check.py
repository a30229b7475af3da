"""Multi-GPU consistency check (not a pytest test; run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_multi_check.py

  1. tiled_sample(shard_tiles=True): the tiles of every step split over the ranks with one all-gather per step must
     give every rank the image the single-GPU run produces, bit for bit;
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
    spec = O.UnetSpec(dim=64)
    unet = M.ConditionalSRUnet(dim=64, learned_sinusoidal_cond=True, learned_sinusoidal_dim=32, num_classes=3)
    diff = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=256, num_sample_steps=6)
    diff.load_state_dict(O.make_state_dict(spec, 22), strict=True)
    diff = diff.eval().to(dev)
    diff.progress = False
    g = torch.Generator().manual_seed(5)
    cond01 = torch.rand(1, 3, 288, 320, generator=g).to(dev)                 # canvas 768x768: 9 / 4 tiles per step
    label = torch.tensor([2], device=dev)
    outs = {}
    for shard in (False, True):
        torch.manual_seed(71)
        torch.cuda.manual_seed(71)
        outs[shard] = diff.tiled_sample(batch_size=2, condition_x=cond01, class_label=label, class_cond_scale=3.0,
                                        num_sample_steps=6, shard_tiles=shard)
    same = bool(torch.equal(outs[False], outs[True]))
    flags = [None] * world
    dist.all_gather_object(flags, same)
    # batch sharding
    cond_b = torch.rand(5, 3, 64, 64, generator=g).to(dev)
    diff64 = M.ConditionalContinuousTimeGaussianDiffusionSR(model=unet, image_size=64, num_sample_steps=6)
    diff64.progress = False
    ref = None
    if rank == 0:                                                             # single-GPU reference with the same RNG
        class One:                                                            # world-size-1 view of sample_sharded
            pass
    full = sharding.sample_sharded(diff64, cond_b, class_label=torch.tensor([1], device=dev), class_cond_scale=3.0,
                                   num_sample_steps=6, seed=9)
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
        err = float((full - ref).abs().max())
        print(f"multi-GPU check, world {world}: tiled shard==single on every rank: {flags}; "
              f"sample_sharded vs single max-abs diff {err:.2e}")
        assert all(flags), flags
        assert err < 2e-3, err          # rows are batch-independent up to the LinearAttention split (see test_gpu_unet)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
