"""World-size-2 (and 3) CPU tests of the multi-GPU host logic over the gloo backend: balanced sharding, the
fixed-size gather, and RNG replication (gathered result == single-process result, bit-exact).  The CUDA
denoiser is replaced by a deterministic CPU stand-in with the reference's p_sample surface -- the sharding
code never looks inside it."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from srgd_b200 import sharding


class ToyDiffusion:
    """p_sample/_finalize stand-in: a cheap per-row update that uses every argument the sharding code forwards."""
    num_sample_steps = 5

    def p_sample(self, x, t, cond, label, cs, ccs, t_next, noise=None):
        lab = 0.0 if label is None else label.float().reshape(-1, 1, 1, 1)
        out = 0.9 * x + 0.1 * cond * float(t) + 0.01 * lab * ccs
        if noise is not None:
            out = out + float(t_next) * noise
        return out, x

    def _finalize(self, img):
        return (img.clamp(-1, 1) + 1) * 0.5


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        cond = torch.rand(B, 3, 8, 8, generator=g)
        label = torch.arange(B) % 3
        out = sharding.sample_sharded(ToyDiffusion(), cond, class_label=label, class_cond_scale=3.0, seed=71)
        if rank == 0:
            ret["out"] = out.clone()
        else:
            assert out is None
        # uneven gather of an arbitrary tensor
        lo, hi = sharding.shard_range(B, world, rank)
        rows = torch.arange(B * 2, dtype=torch.float32).reshape(B, 2)[lo:hi]
        full = sharding.gather_rows(rows, sharding.shard_counts(B, world))
        if rank == 0:
            ret["rows"] = full.clone()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_range_is_balanced_and_covers():
    for n in (0, 1, 5, 16, 17, 81):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
            assert sizes == sharding.shard_counts(n, world)
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


@pytest.mark.parametrize("world,B", [(2, 6), (2, 5), (3, 4)])
def test_sharded_sampling_matches_single_process(world, B):
    g = torch.Generator().manual_seed(3)
    cond = torch.rand(B, 3, 8, 8, generator=g)
    label = torch.arange(B) % 3
    ref = sharding.sample_sharded(ToyDiffusion(), cond, class_label=label, class_cond_scale=3.0, seed=71)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), B, ret), nprocs=world, join=True)
    assert torch.equal(ret["out"], ref)                       # bit-identical for any world size
    assert torch.equal(ret["rows"], torch.arange(B * 2, dtype=torch.float32).reshape(B, 2))
