"""CPU tests of the large-image orchestration (srgd_b200/tiled.py): the product's `run_tiled` loop driven by a
plain-torch ops stand-in must reproduce the oracle's tiled_sample (model.py:3288-3413) bit for bit in one process
(the reference's minibatch partition), and -- in tile-granular exact mode (`shard=True`: contiguous tile ranges per
rank, regrouped denoiser calls, one all_gather_into_tensor per step, replicated RNG) -- yield the same image on
1, 2 and 3 gloo ranks bit for bit.  The exact mode needs a denoiser whose rows do not depend on their batch
(the library's batch-invariant mode on the GPU); the CPU stand-in gets that by denoising row by row."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import srgd_oracle as O  # checker + CPU stand-in for the CUDA denoiser
from srgd_b200.tiled import run_tiled
from srgd_b200.tiling import TilePlan

TILE, STEPS, BATCH = 32, 5, 5
SPEC = O.UnetSpec(dim=16)


class TorchOps:
    """ops interface of run_tiled on CPU tensors; the denoiser is the oracle's p_sample."""

    def __init__(self, sd, gen, row_by_row=False):
        self.sd, self.gen, self.row_by_row = sd, gen, row_by_row
        self.invariant_calls = []

    def set_batch_invariant(self, on):
        self.invariant_calls.append(bool(on))
        return False

    def randn(self, shape, device):
        return torch.randn(shape, generator=self.gen)

    def p_sample(self, xt, t, ct, label, cs, ccs, t_next, noise):
        if self.row_by_row:                      # batch-invariant stand-in: a row never sees its batch neighbours
            rows = [O.p_sample(self.sd, SPEC, xt[k:k + 1], t, ct[k:k + 1], label, cs, ccs, t_next,
                               noise=None if noise is None else noise[k:k + 1]) for k in range(xt.shape[0])]
            return torch.cat([r[0] for r in rows], 0), torch.cat([r[1] for r in rows], 0)
        return O.p_sample(self.sd, SPEC, xt, t, ct, label, cs, ccs, t_next, noise=noise)

    def gather(self, canvas, coords, tile):
        return torch.cat([canvas[:, :, y:y + tile, x:x + tile] for y, x in coords], 0)

    def scatter(self, canvas, coords, tiles, tile):
        for k, (y, x) in enumerate(coords):
            canvas[:, :, y:y + tile, x:x + tile] = tiles[k]

    def renoise_outside(self, canvas, noise, sigma, inner):
        it, ib, il, ir = inner
        fresh = noise * sigma
        fresh[:, :, it:ib, il:ir] = canvas[:, :, it:ib, il:ir]
        canvas.copy_(fresh)

    def sigma(self, t):
        return float((-O.log_snr_linear(torch.as_tensor(t, dtype=torch.float32))).sigmoid().sqrt())


def _inputs():
    g = torch.Generator().manual_seed(7)
    cond01 = torch.rand(1, 3, 104, 120, generator=g)
    return cond01, torch.tensor([1])


def _product_path(sd, shard, max_rows=64, stride=TILE, tile=TILE, nsteps=STEPS):
    """What ConditionalContinuousTimeGaussianDiffusionSR.tiled_sample does around run_tiled, on CPU tensors."""
    cond01, label = _inputs()
    gen = torch.Generator().manual_seed(71)
    cond = cond01 * 2 - 1
    plan = TilePlan(cond.shape[2], cond.shape[3], tile, stride)
    cond = F.pad(cond, plan.canvas_pad, mode="reflect")
    img = torch.randn(cond.shape, generator=gen)
    it, ib, il, ir = plan.inner
    cond_canvas = torch.zeros_like(cond)
    cond_canvas[:, :, it:ib, il:ir] = cond[:, :, it:ib, il:ir]
    steps = torch.linspace(1., 0., nsteps + 1)
    ops = TorchOps(sd, gen, row_by_row=shard)
    img, _ = run_tiled(ops, img, cond_canvas, plan, steps, nsteps, BATCH, label, 1.0, 0, 2.0, 0, 0,
                       shard=shard, max_rows=max_rows)
    assert ops.invariant_calls == ([True, False] if shard else [])      # switched on for the loop, restored after
    top, bottom, left, right = plan.crop
    return (img[:, :, top:bottom, left:right].clamp(-1, 1) + 1) * 0.5


_memo = {}


def _reference(sd):
    """The oracle's image for the fixed inputs and weights of this file (computed once per process)."""
    if "ref" not in _memo:
        cond01, label = _inputs()
        gen = torch.Generator().manual_seed(71)
        _memo["ref"] = O.tiled_sample(sd, SPEC, BATCH, cond01, label, class_cond_scale=2.0, num_sample_steps=STEPS,
                                      tile_size=TILE, tile_stride=TILE, generator=gen)
    return _memo["ref"]


def test_overlapping_tiles_follow_the_reference_partition_and_refuse_exact_mode():
    """tile_stride < tile_size: the shifted grid's tiles overlap and the reference advances the canvas in place,
    minibatch by minibatch (model.py:3374-3385).  The default mode keeps that partition and equals the oracle bit for
    bit; exact mode would regroup the calls and is refused."""
    torch.set_num_threads(2)
    sd = O.make_state_dict(SPEC, 11)
    cond01, label = _inputs()
    ref = O.tiled_sample(sd, SPEC, BATCH, cond01, label, class_cond_scale=2.0, num_sample_steps=3, tile_size=64,
                         tile_stride=48, generator=torch.Generator().manual_seed(71))
    assert torch.equal(_product_path(sd, shard=False, stride=48, tile=64, nsteps=3), ref)
    with pytest.raises(ValueError, match="disjoint"):
        _product_path(sd, shard=True, stride=48, tile=64, nsteps=3)
    # a tile size that does not divide the canvas: the last tile of each axis is pulled back and overlaps its neighbour
    ref = O.tiled_sample(sd, SPEC, BATCH, cond01, label, class_cond_scale=2.0, num_sample_steps=3, tile_size=48,
                         tile_stride=48, generator=torch.Generator().manual_seed(71))
    assert torch.equal(_product_path(sd, shard=False, stride=48, tile=48, nsteps=3), ref)
    with pytest.raises(ValueError, match="disjoint"):
        _product_path(sd, shard=True, stride=48, tile=48, nsteps=3)


def _single_exact(sd):
    """The single-process exact-mode image (computed once per process)."""
    if "exact" not in _memo:
        _memo["exact"] = _product_path(sd, shard=True, max_rows=64)
    return _memo["exact"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = _product_path(O.make_state_dict(SPEC, 11), shard=True)
        ret[rank] = out.clone()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_run_tiled_matches_oracle_single_process():
    torch.set_num_threads(2)
    sd = O.make_state_dict(SPEC, 11)
    assert torch.equal(_product_path(sd, shard=False), _reference(sd))


def test_run_tiled_exact_mode_single_process():
    """shard=True without a process group: the regrouped calls (any max_rows) give one and the same image, which
    differs from the reference's minibatch partition only by the CPU conv's batch-dependent blocking (<= 5e-4)."""
    torch.set_num_threads(2)
    sd = O.make_state_dict(SPEC, 11)
    a = _single_exact(sd)
    b = _product_path(sd, shard=True, max_rows=3)
    assert torch.equal(a, b)
    torch.testing.assert_close(a, _reference(sd), rtol=0, atol=5e-4)


@pytest.mark.parametrize("world", [2, 3])
def test_run_tiled_sharded_over_gloo_is_bit_identical(world):
    """20 / 12 tiles per step over 2 / 3 ranks (contiguous ranges of 10+10 / 7+7+6 and 6+6 / 4+4+4 tiles): every
    replica ends with the single-process exact-mode image, bit for bit."""
    torch.set_num_threads(2)
    sd = O.make_state_dict(SPEC, 11)
    single = _single_exact(sd)
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        assert torch.equal(ret[r], single), f"rank {r} differs from the single-process exact-mode run"
    torch.testing.assert_close(single, _reference(sd), rtol=0, atol=5e-4)


def _multi_image_path(sd, conds01, label, max_rows, nsteps):
    """N same-sized images advanced together (run_tiled's N > 1 extension): shared noise stream, stacked minibatches."""
    gen = torch.Generator().manual_seed(71)
    cond = conds01 * 2 - 1
    plan = TilePlan(cond.shape[2], cond.shape[3], TILE, TILE)
    cond = F.pad(cond, plan.canvas_pad, mode="reflect")
    img = torch.randn((1,) + tuple(cond.shape[1:]), generator=gen).repeat(cond.shape[0], 1, 1, 1)
    it, ib, il, ir = plan.inner
    cond_canvas = torch.zeros_like(cond)
    cond_canvas[:, :, it:ib, il:ir] = cond[:, :, it:ib, il:ir]
    steps = torch.linspace(1., 0., nsteps + 1)
    img, _ = run_tiled(TorchOps(sd, gen), img, cond_canvas, plan, steps, nsteps, BATCH, label, 1.0, 0, 2.0, 0, 0,
                       max_rows=max_rows)
    top, bottom, left, right = plan.crop
    return (img[:, :, top:bottom, left:right].clamp(-1, 1) + 1) * 0.5


@pytest.mark.parametrize("max_rows", [4, 64])
def test_run_tiled_many_images_equal_consecutive_single_runs(max_rows):
    """Two images of one size in one run == two runs of one image each, every run reseeded with the same seed (what the
    reference CLI does per image, inference.py:81).  max_rows = 4 forces one image per denoiser call."""
    torch.set_num_threads(4)
    nsteps = 3
    sd = O.make_state_dict(SPEC, 11)
    g = torch.Generator().manual_seed(13)
    conds01 = torch.rand(2, 3, 104, 120, generator=g)
    label = torch.tensor([2])
    together = _multi_image_path(sd, conds01, label, max_rows, nsteps)
    for k in range(2):
        alone = O.tiled_sample(sd, SPEC, BATCH, conds01[k:k + 1], label, class_cond_scale=2.0, num_sample_steps=nsteps,
                               tile_size=TILE, tile_stride=TILE, generator=torch.Generator().manual_seed(71))
        # rows of a stacked denoiser batch are independent, but the CPU conv blocks differently per batch size and the
        # first steps amplify by 1/alpha: compare to 5e-4
        torch.testing.assert_close(together[k:k + 1], alone, rtol=0, atol=5e-4)   # (a logic error would be O(0.1))


def _multi_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd = O.make_state_dict(SPEC, 11)
        g = torch.Generator().manual_seed(13)
        conds01 = torch.rand(2, 3, 104, 120, generator=g)
        gen = torch.Generator().manual_seed(71)
        cond = conds01 * 2 - 1
        plan = TilePlan(cond.shape[2], cond.shape[3], TILE, TILE)
        cond = F.pad(cond, plan.canvas_pad, mode="reflect")
        img = torch.randn((1,) + tuple(cond.shape[1:]), generator=gen).repeat(2, 1, 1, 1)
        it, ib, il, ir = plan.inner
        cond_canvas = torch.zeros_like(cond)
        cond_canvas[:, :, it:ib, il:ir] = cond[:, :, it:ib, il:ir]
        steps = torch.linspace(1., 0., 3)
        img, _ = run_tiled(TorchOps(sd, gen, row_by_row=True), img, cond_canvas, plan, steps, 2, BATCH,
                           torch.tensor([2]), 1.0, 0, 2.0, 0, 0, shard=True, max_rows=64)
        ret[rank] = img.clone()
        if world > 1:
            dist.barrier()
    finally:
        dist.destroy_process_group()


def test_run_tiled_many_images_sharded_over_gloo():
    """Two images advancing together with the tiles of every step split over two gloo ranks: both replicas end
    with the canvases of the single-process exact-mode run, bit for bit."""
    ret = mp.Manager().dict()
    mp.spawn(_multi_worker, args=(1, _free_port(), ret), nprocs=1, join=True)
    single = ret[0]
    ret2 = mp.Manager().dict()
    mp.spawn(_multi_worker, args=(2, _free_port(), ret2), nprocs=2, join=True)
    for r in range(2):
        assert torch.equal(ret2[r], single), f"rank {r} differs from the single-process run"


def test_run_tiled_step_gating_options():
    """generation_start_steps (start from q_sample(condition), skip the first steps) and class_guidance_start_steps
    (scale 1 before that step) in the product loop == the oracle's tiled_sample (model.py:3305-3309, 3349-3360)."""
    torch.set_num_threads(4)
    sd = O.make_state_dict(SPEC, 11)
    cond01, label = _inputs()
    gen_start, cls_start = 1, 3
    gen = torch.Generator().manual_seed(71)
    cond = cond01 * 2 - 1
    plan = TilePlan(cond.shape[2], cond.shape[3], TILE, TILE)
    cond = F.pad(cond, plan.canvas_pad, mode="reflect")
    # what ConditionalContinuousTimeGaussianDiffusionSR.tiled_sample does for generation_start_steps > 0
    start = torch.tensor(1. - gen_start / STEPS)
    img, _ = O.q_sample(cond, start, generator=gen)
    it, ib, il, ir = plan.inner
    cond_canvas = torch.zeros_like(cond)
    cond_canvas[:, :, it:ib, il:ir] = cond[:, :, it:ib, il:ir]
    steps = torch.linspace(1., 0., STEPS + 1)
    img, _ = run_tiled(TorchOps(sd, gen), img, cond_canvas, plan, steps, STEPS, BATCH, label, 1.0, 0, 2.0, cls_start,
                       gen_start)
    top, bottom, left, right = plan.crop
    got = (img[:, :, top:bottom, left:right].clamp(-1, 1) + 1) * 0.5
    ref = O.tiled_sample(sd, SPEC, BATCH, cond01, label, class_cond_scale=2.0, class_guidance_start_steps=cls_start,
                         generation_start_steps=gen_start, num_sample_steps=STEPS, tile_size=TILE, tile_stride=TILE,
                         generator=torch.Generator().manual_seed(71))
    assert torch.equal(got, ref)
