"""Large-image sampling orchestration (reference: tiled_sample, model.py:3288-3413) written against a small
`ops` interface so that the same host logic runs

  * on the B200 through the CUDA library (`CudaTiledOps`: batched tile gather / scatter kernels, fused re-noise,
    the fused `p_sample`), and
  * in the CPU unit tests over the gloo backend with a plain-torch stand-in (tests/test_tiled_gloo.py).

Algorithm (unchanged from the reference): reflect-pad the condition to the tile canvas, then per step denoise
every 256x256 tile of the current grid -- the aligned grid on even steps, the grid shifted by half a tile on odd
steps -- in minibatches of `batch_size` tiles; after odd steps everything outside the hull of the shifted grid is
replaced by fresh noise at the next noise level.

Multi-GPU ("exact mode"): tiles of one step are independent, but step i+1's grid straddles
step i's tiles, so every rank keeps a full replica of the canvas, denoises the minibatches assigned to it
(round-robin over the reference's minibatch order) and the freshly written tiles are all-gathered once per step.
RNG parity: every rank draws the noise of ALL minibatches in the reference's order (same generator state on every
rank) and uses only its own, so the result is bit-identical to the single-process run for any world size.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .tiling import TilePlan


class CudaTiledOps:
    """Product ops: everything that touches pixels runs in libsrgd_b200.so."""

    def __init__(self, diffusion):
        from . import _lib
        self._lib = _lib
        self.d = diffusion

    def randn(self, shape, device):
        return self.d._randn(shape, device)

    def p_sample(self, xt, t, ct, label, cs, ccs, t_next, noise):
        return self.d.p_sample(xt, t, ct, label, cs, ccs, t_next, noise=noise)

    def _coords(self, coords: Sequence[Tuple[int, int]]):
        tc = self._lib.TileCoords()
        tc.n = len(coords)
        for k, (y, x) in enumerate(coords):
            tc.yx[k][0], tc.yx[k][1] = int(y), int(x)
        return tc

    def gather(self, canvas: torch.Tensor, coords, tile: int) -> torch.Tensor:
        _, ch, H, W = canvas.shape
        lib, L = self._lib.load(), self._lib
        out = torch.empty(len(coords), ch, tile, tile, device=canvas.device, dtype=torch.float32)
        for first in range(0, len(coords), L.SRGD_MAX_TILES_PER_CALL):
            part = coords[first:first + L.SRGD_MAX_TILES_PER_CALL]
            tc = self._coords(part)
            with torch.cuda.device(canvas.device):
                L.check(lib.srgd_gather_tiles(L.ptr(canvas), L.ptr(out[first:]), C.byref(tc), ch, H, W, tile,
                                              L.current_stream()), "srgd_gather_tiles")
        return out

    def scatter(self, canvas: torch.Tensor, coords, tiles: torch.Tensor, tile: int) -> None:
        _, ch, H, W = canvas.shape
        lib, L = self._lib.load(), self._lib
        tiles = tiles.contiguous()
        for first in range(0, len(coords), L.SRGD_MAX_TILES_PER_CALL):
            part = coords[first:first + L.SRGD_MAX_TILES_PER_CALL]
            tc = self._coords(part)
            with torch.cuda.device(canvas.device):
                L.check(lib.srgd_scatter_tiles(L.ptr(canvas), L.ptr(tiles[first:]), C.byref(tc), ch, H, W, tile,
                                               L.current_stream()), "srgd_scatter_tiles")

    def renoise_outside(self, canvas: torch.Tensor, noise: torch.Tensor, sigma: float, inner) -> None:
        _, ch, H, W = canvas.shape
        it, ib, il, ir = inner
        lib, L = self._lib.load(), self._lib
        with torch.cuda.device(canvas.device):
            L.check(lib.srgd_renoise_outside(L.ptr(canvas), L.ptr(noise.contiguous()), ch, H, W, it, ib, il, ir,
                                             float(sigma), L.current_stream()), "srgd_renoise_outside")

    def sigma(self, t) -> float:
        from .diffusion import _host_scalar
        return float((-self.d.log_snr(_host_scalar(t))).sigmoid().sqrt())

    def finalize(self, img):
        return self.d._finalize(img)


def _chunks(tiles: List[Tuple[int, int]], batch_size: int) -> List[List[Tuple[int, int]]]:
    return [tiles[i:i + batch_size] for i in range(0, len(tiles), batch_size)]


def _all_gather_tiles(local: Optional[torch.Tensor], counts: List[int], shape_tail, device, dtype, group):
    """All-gather per-rank tile stacks of different lengths (padded to the largest)."""
    width = max(counts)
    buf = torch.zeros((width,) + tuple(shape_tail), device=device, dtype=dtype)
    if local is not None and local.shape[0] > 0:
        buf[:local.shape[0]] = local
    outs = [torch.empty_like(buf) for _ in counts]
    dist.all_gather(outs, buf, group=group)
    return [o[:c] for o, c in zip(outs, counts)]


def run_tiled(ops, img: torch.Tensor, cond_canvas: torch.Tensor, plan: TilePlan, steps: torch.Tensor,
              num_sample_steps: int, batch_size: int, class_label, cond_scale: float, guidance_start_steps: int,
              class_cond_scale: float, class_guidance_start_steps: int, generation_start_steps: int,
              x_start: Optional[torch.Tensor] = None, on_step=None, group=None, shard: bool = False,
              max_rows: int = 64):
    """The sampling loop of tiled_sample (model.py:3345-3401) on an initial noise canvas `img` [N,3,H,W] and the
    hull-masked condition canvas.  Updates and returns `img` (and `x_start` if given).  With `shard=True` and an
    initialised process group the minibatches of every step are split over the ranks (see module docstring).

    N > 1 (extension; the reference's loop only works for N = 1): N images of the same size advance together and
    SHARE the noise stream, which is exactly what N consecutive runs of the reference CLI produce -- it reseeds every
    generator before each image (inference.py:81).  The tiles of a minibatch are stacked image-major into one
    denoiser batch of at most `max_rows` rows, so the 4-tile odd steps of a small image still fill the GPU."""
    tile = plan.tile_size
    world = dist.get_world_size(group) if (shard and dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = img.device
    n_img = img.shape[0]
    # the condition tiles of the two grids never change: gather them once per minibatch
    cond_cache = {}

    def gather_all(canvas, chunk, lo, hi):
        if hi - lo == 1:
            return ops.gather(canvas[lo:lo + 1], chunk, tile)
        return torch.cat([ops.gather(canvas[k:k + 1], chunk, tile) for k in range(lo, hi)], 0)

    def scatter_all(canvas, coords, stack):                # stack: [n_img * len(coords), ...] image-major
        n = len(coords)
        for k in range(n_img):
            ops.scatter(canvas[k:k + 1], coords, stack[k * n:(k + 1) * n], tile)

    for i in range(num_sample_steps):
        if i < generation_start_steps:
            continue
        cs = 1.0 if i < guidance_start_steps else cond_scale
        ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
        last = float(steps[i + 1]) == 0.0
        chunks = _chunks(plan.grids[i % 2], batch_size)
        mine_out, mine_x0, mine_idx = [], [], []
        for ci, chunk in enumerate(chunks):
            # RNG: one draw per minibatch in the reference's order on EVERY rank (model.py:3187), none on the last step
            noise = None if last else ops.randn((len(chunk), img.shape[1], tile, tile), dev)
            if ci % world != rank:
                continue
            per_call = max(1, max_rows // len(chunk))          # images per denoiser call
            outs, x0s = [], []
            for lo in range(0, n_img, per_call):
                hi = min(n_img, lo + per_call)
                key = (i % 2, ci, lo)
                if key not in cond_cache:
                    cond_cache[key] = gather_all(cond_canvas, chunk, lo, hi)
                xt = gather_all(img, chunk, lo, hi)
                nz = noise if (noise is None or hi - lo == 1) else noise.repeat(hi - lo, 1, 1, 1)
                out, x0 = ops.p_sample(xt, steps[i], cond_cache[key], class_label, cs, ccs, steps[i + 1], nz)
                outs.append(out)
                x0s.append(x0)
            # [n_img * len(chunk), ...] image-major
            mine_out.append(outs[0] if len(outs) == 1 else torch.cat(outs, 0))
            mine_x0.append(x0s[0] if len(x0s) == 1 else torch.cat(x0s, 0))
            mine_idx.append(ci)

        if world == 1:
            for ci, out, x0 in zip(mine_idx, mine_out, mine_x0):
                scatter_all(img, chunks[ci], out)
                if x_start is not None:
                    scatter_all(x_start, chunks[ci], x0)
        else:
            # exchange step: every rank contributes the tiles it wrote; all replicas apply all of them
            counts = [n_img * sum(len(chunks[ci]) for ci in range(r, len(chunks), world)) for r in range(world)]
            tail = (img.shape[1], tile, tile)
            local = torch.cat(mine_out, 0) if mine_out else None
            gathered = _all_gather_tiles(local, counts, tail, dev, img.dtype, group)
            gathered_x0 = None
            if x_start is not None:
                local0 = torch.cat(mine_x0, 0) if mine_x0 else None
                gathered_x0 = _all_gather_tiles(local0, counts, tail, dev, img.dtype, group)
            for r in range(world):
                pos = 0
                for ci in range(r, len(chunks), world):
                    n = n_img * len(chunks[ci])
                    scatter_all(img, chunks[ci], gathered[r][pos:pos + n])
                    if gathered_x0 is not None:
                        scatter_all(x_start, chunks[ci], gathered_x0[r][pos:pos + n])
                    pos += n
        if i % 2 == 1:
            # outside the hull of the shifted grid the state is replaced by fresh noise at the next noise level
            # (q_sample of zeros, model.py:3392-3396); the draw covers the whole canvas like the reference's
            fresh = ops.randn((1,) + tuple(img.shape[1:]), dev)
            sig = ops.sigma(steps[i + 1])
            for k in range(n_img):
                ops.renoise_outside(img[k:k + 1], fresh, sig, plan.inner)
        if on_step is not None:
            on_step(i, img, x_start)
    return img, x_start
