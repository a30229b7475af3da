"""Multi-GPU sharding of the sampling path: one process per GPU, independent units, one gather.

The reference has no distributed code at all; its only scale-out mechanism is the manual
`--start_index/--end_index` split of the sorted file list (inference.py:36-37, 120).  Images and
`sample()` batch rows share nothing, so ranks take contiguous balanced slices, run the *unchanged*
hot path on their slice with replicated weights, and the finished images are gathered once at the
end (`torch.distributed`: NCCL over NVLink on GPUs, gloo in the CPU tests).  There is no collective
inside the sampling loop.

RNG parity: `sample()` draws `randn([B,3,H,W])` once and `randn_like` once per step (model.py:3203,
3187), so noise depends on the batch size.  `sample_sharded(..., replicate_rng=True)` makes every
rank draw the FULL-batch tensors from identically seeded generators and keep its own rows: the
gathered result is then bit-identical to the single-process result for any world size.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced slice [start, end) of `n_items` for `rank` (first `n % world` ranks get one more)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_counts(n_items: int, world_size: int) -> List[int]:
    return [shard_range(n_items, world_size, r)[1] - shard_range(n_items, world_size, r)[0] for r in range(world_size)]


def gather_rows(local: torch.Tensor, counts: Sequence[int], dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather per-rank row blocks `local[count_r, ...]` to `dst`; returns the concatenation there, None elsewhere.
    Blocks are padded to the largest count so that one fixed-size `dist.gather` suffices."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    assert len(counts) == world and local.shape[0] == counts[rank]
    width = max(counts)
    padded = local
    if local.shape[0] < width:
        padded = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    padded = padded.contiguous()
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], 0)


@torch.inference_mode()
def sample_sharded(diffusion, condition_x: torch.Tensor, class_label=None, cond_scale: float = 1.0,
                   class_cond_scale: float = 1.0, num_sample_steps: Optional[int] = None, seed: Optional[int] = None,
                   replicate_rng: bool = True, dst: int = 0, group=None):
    """`diffusion.sample()` over the rows of `condition_x` ([B,3,S,S] in [0,1], the same on every rank), sharded
    across the ranks of `group`.  Returns the [B,3,S,S] result on `dst` (None on other ranks).

    `diffusion` only needs the reference surface: `p_sample(x, t, cond, label, cs, ccs, t_next, noise=)`,
    `num_sample_steps`, and `_finalize(img)` (clamp + [0,1], model.py:3237-3238)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    B = condition_x.shape[0]
    lo, hi = shard_range(B, world, rank)
    counts = shard_counts(B, world)
    steps_n = diffusion.num_sample_steps if num_sample_steps is None else num_sample_steps
    dev = condition_x.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) if seed is not None else 0)
    cond = condition_x * 2 - 1                                                  # model.py:40
    label = class_label
    if label is not None and label.numel() == B and B > 1:
        label_local = label[lo:hi]
    else:
        label_local = label
    shape = tuple(cond.shape)

    def draw():
        if replicate_rng:
            return torch.randn(shape, generator=gen, device=dev)[lo:hi]
        return torch.randn((hi - lo,) + shape[1:], generator=gen, device=dev)

    img = draw()                                                                # model.py:3203
    steps = torch.linspace(1., 0., steps_n + 1)
    cond_local = cond[lo:hi].contiguous()
    for i in range(steps_n):
        last = (i == steps_n - 1)
        noise = None if last else draw()                                        # model.py:3184-3187
        if hi > lo:
            img, _ = diffusion.p_sample(img, steps[i], cond_local, label_local, cond_scale, class_cond_scale,
                                        steps[i + 1], noise=noise)
    out = diffusion._finalize(img) if hi > lo else img
    return gather_rows(out, counts, dst=dst, group=group)
