"""Large-image sampling orchestration (reference: tiled_sample, model.py:3288-3413) written against a small
`ops` interface so that the same host logic runs

  * on the B200 through the CUDA library (`CudaTiledOps`: batched tile gather / scatter kernels, fused re-noise,
    the fused `p_sample`), and
  * in the CPU unit tests over the gloo backend with a plain-torch stand-in (tests/test_tiled_gloo.py).

Algorithm (unchanged from the reference): reflect-pad the condition to the tile canvas, then per step denoise
every 256x256 tile of the current grid -- the aligned grid on even steps, the grid shifted by half a tile on odd
steps -- in minibatches of `batch_size` tiles; after odd steps everything outside the hull of the shifted grid is
replaced by fresh noise at the next noise level.

Multi-GPU ("exact mode", SURVEY.md section 8e, `shard=True`): tiles of one step are independent, but step i+1's
grid straddles step i's tiles, so every rank keeps a full replica of the canvas, denoises a CONTIGUOUS RANGE OF
TILES of the step's grid (row bands; balanced to within one tile for any world size and independent of the
reference's `batch_size`), and the freshly written tiles travel in ONE pre-sized `all_gather_into_tensor` per step,
issued asynchronously so that the odd steps' full-canvas noise draw overlaps it.
RNG parity: every rank draws the noise of ALL minibatches in the reference's order and shapes (same generator state
on every rank) and uses the rows of its own tiles.  Denoiser calls are regrouped (up to `max_rows` rows per call
instead of `batch_size`), which is only legitimate because the library runs in batch-invariant mode there
(`srgd_set_batch_invariant`: a tile's result does not depend on which other tiles share its launch) -- the image is
then bit-identical for every world size, 1 included.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .sharding import shard_range
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

    def set_batch_invariant(self, on: bool) -> bool:
        """Returns the previous setting (include/srgd_b200.h: srgd_set_batch_invariant)."""
        return bool(self._lib.load().srgd_set_batch_invariant(1 if on else 0))


def _chunks(tiles: List[Tuple[int, int]], batch_size: int) -> List[List[Tuple[int, int]]]:
    return [tiles[i:i + batch_size] for i in range(0, len(tiles), batch_size)]


def _rows_of(noises: List[torch.Tensor], batch_size: int, lo: int, hi: int) -> torch.Tensor:
    """Rows [lo, hi) of the per-minibatch noise draws `noises` (minibatch k holds tiles [k * batch_size, ...))."""
    parts = []
    t = lo
    while t < hi:
        k, off = divmod(t, batch_size)
        take = min(hi - t, noises[k].shape[0] - off)
        parts.append(noises[k][off:off + take])
        t += take
    return parts[0] if len(parts) == 1 else torch.cat(parts, 0)


class _Exchange:
    """Pre-sized send / receive buffers of one tile grid: send [n_img, width, C, T, T], recv [world, n_img, width, ...]
    where width = the largest per-rank tile count; one all_gather_into_tensor per step fills `recv`."""

    def __init__(self, n_tiles, n_img, world, tail, device, dtype, want_x0):
        self.ranges = [shard_range(n_tiles, world, r) for r in range(world)]
        width = max(1, max(hi - lo for lo, hi in self.ranges))
        shape = (n_img, width) + tuple(tail)
        self.send = torch.zeros(shape, device=device, dtype=dtype)
        self.recv = torch.empty((world,) + shape, device=device, dtype=dtype) if world > 1 else self.send[None]
        self.send_x0 = torch.zeros(shape, device=device, dtype=dtype) if want_x0 else None
        self.recv_x0 = None
        if want_x0:
            self.recv_x0 = torch.empty((world,) + shape, device=device, dtype=dtype) if world > 1 else self.send_x0[None]


def run_tiled(ops, img: torch.Tensor, cond_canvas: torch.Tensor, plan: TilePlan, steps: torch.Tensor,
              num_sample_steps: int, batch_size: int, class_label, cond_scale: float, guidance_start_steps: int,
              class_cond_scale: float, class_guidance_start_steps: int, generation_start_steps: int,
              x_start: Optional[torch.Tensor] = None, on_step=None, group=None, shard: bool = False,
              max_rows: int = 64):
    """The sampling loop of tiled_sample (model.py:3345-3401) on an initial noise canvas `img` [N,3,H,W] and the
    hull-masked condition canvas.  Updates and returns `img` (and `x_start` if given).

    shard=False: the reference's own partition -- one denoiser call per minibatch of `batch_size` tiles.
    shard=True : tile-granular exact mode (module docstring): contiguous tile ranges per rank of the initialised
    process group (or the whole grid without one), calls of up to `max_rows` rows, batch-invariant kernels, one
    all-gather per step.

    N > 1 (extension; the reference's loop only works for N = 1): N images of the same size advance together and
    SHARE the noise stream, which is exactly what N consecutive runs of the reference CLI produce -- it reseeds every
    generator before each image (inference.py:81).  The tiles of a call are stacked image-major into one denoiser
    batch of at most `max_rows` rows, so the 4-tile odd steps of a small image still fill the GPU."""
    tile = plan.tile_size
    if shard and not plan.disjoint:
        raise ValueError("shard_tiles (exact mode) regroups the tiles of a step into different denoiser calls, which is "
                         "only equivalent for disjoint tiles: use tile_stride == tile_size with a tile size that divides "
                         "the 256-aligned canvas, or the default mode")
    world = dist.get_world_size(group) if (shard and dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = img.device
    n_img = img.shape[0]
    tail = (img.shape[1], tile, tile)
    # the condition tiles of the two grids never change: gather them once per denoiser call slot
    cond_cache = {}

    def gather_all(canvas, chunk, lo, hi):
        if hi - lo == 1:
            return ops.gather(canvas[lo:lo + 1], chunk, tile)
        return torch.cat([ops.gather(canvas[k:k + 1], chunk, tile) for k in range(lo, hi)], 0)

    disjoint = plan.disjoint

    def scatter_all(canvas, coords, stack):                # stack: [n_img * len(coords), ...] image-major
        n = len(coords)
        for k in range(n_img):
            if disjoint:
                ops.scatter(canvas[k:k + 1], coords, stack[k * n:(k + 1) * n], tile)
            else:                                          # overlapping tiles: the later tile wins, as in the
                for j, c in enumerate(coords):             # reference's sequential slice assignments (model.py:3383)
                    ops.scatter(canvas[k:k + 1], [c], stack[k * n + j:k * n + j + 1], tile)

    def denoise(i, parity, slot, chunk, noise, cs, ccs, emit):
        """One group of tiles `chunk` for all images, in calls of at most max_rows rows; emit(lo, hi, out, x0) receives
        the image-major results of images [lo, hi)."""
        per_call = max(1, max_rows // len(chunk))          # images per denoiser call
        for lo in range(0, n_img, per_call):
            hi = min(n_img, lo + per_call)
            key = (parity, slot, lo)
            if key not in cond_cache:
                cond_cache[key] = gather_all(cond_canvas, chunk, lo, hi)
            xt = gather_all(img, chunk, lo, hi)
            nz = noise if (noise is None or hi - lo == 1) else noise.repeat(hi - lo, 1, 1, 1)
            out, x0 = ops.p_sample(xt, steps[i], cond_cache[key], class_label, cs, ccs, steps[i + 1], nz)
            emit(lo, hi, out, x0)

    exchanges = {}
    prev_invariant = None
    if shard and hasattr(ops, "set_batch_invariant"):
        prev_invariant = ops.set_batch_invariant(True)
    try:
        for i in range(num_sample_steps):
            if i < generation_start_steps:
                continue
            cs = 1.0 if i < guidance_start_steps else cond_scale
            ccs = 1.0 if i < class_guidance_start_steps else class_cond_scale
            last = float(steps[i + 1]) == 0.0
            parity = i % 2
            tiles = plan.grids[parity]
            chunks = _chunks(tiles, batch_size)
            fresh = None
            if not shard:
                outs = []
                for ci, chunk in enumerate(chunks):
                    # RNG: one draw per minibatch in the reference's order (model.py:3187), none on the last step
                    noise = None if last else ops.randn((len(chunk),) + tail, dev)
                    got = []
                    denoise(i, parity, ci, chunk, noise, cs, ccs, lambda lo, hi, out, x0: got.append((out, x0)))
                    # [n_img * len(chunk), ...] image-major
                    outs.append((chunk, got[0][0] if len(got) == 1 else torch.cat([g[0] for g in got], 0),
                                 got[0][1] if len(got) == 1 else torch.cat([g[1] for g in got], 0)))
                    if not disjoint:
                        # overlapping tiles (tile_stride < tile_size): the reference advances the canvas in place,
                        # so the next minibatch must see this one's pixels (model.py:3374-3385)
                        scatter_all(img, *outs[-1][:2])
                        if x_start is not None:
                            scatter_all(x_start, outs[-1][0], outs[-1][2])
                        outs.pop()
                for chunk, out, x0 in outs:
                    scatter_all(img, chunk, out)
                    if x_start is not None:
                        scatter_all(x_start, chunk, x0)
            else:
                ex = exchanges.get(parity)
                if ex is None:
                    ex = exchanges[parity] = _Exchange(len(tiles), n_img, world, tail, dev, img.dtype,
                                                       x_start is not None)
                # RNG: every rank draws the noise of ALL minibatches, same shapes and order as the reference
                noises = None if last else [ops.randn((len(chunk),) + tail, dev) for chunk in chunks]
                lo_t, hi_t = ex.ranges[rank]
                for a in range(lo_t, hi_t, max_rows):
                    sub = tiles[a:min(hi_t, a + max_rows)]
                    nz = None if last else _rows_of(noises, batch_size, a, a + len(sub))

                    def emit(lo, hi, out, x0, a=a, n=len(sub)):
                        ex.send[lo:hi, a - lo_t:a - lo_t + n] = out.reshape((hi - lo, n) + tail)
                        if ex.send_x0 is not None:
                            ex.send_x0[lo:hi, a - lo_t:a - lo_t + n] = x0.reshape((hi - lo, n) + tail)

                    denoise(i, parity, a, sub, nz, cs, ccs, emit)
                works = []
                if world > 1:
                    # exchange step: ONE pre-sized all-gather of the tiles every rank wrote (a gather, never a reduction)
                    works.append(dist.all_gather_into_tensor(ex.recv.flatten(0, 1), ex.send, group=group, async_op=True))
                    if ex.send_x0 is not None:
                        works.append(dist.all_gather_into_tensor(ex.recv_x0.flatten(0, 1), ex.send_x0, group=group,
                                                                 async_op=True))
                if parity == 1:
                    fresh = ops.randn((1,) + tuple(img.shape[1:]), dev)      # overlaps the exchange
                for w in works:
                    w.wait()
                for r, (lo_r, hi_r) in enumerate(ex.ranges):
                    if hi_r > lo_r:
                        for k in range(n_img):
                            ops.scatter(img[k:k + 1], tiles[lo_r:hi_r], ex.recv[r, k, :hi_r - lo_r], tile)
                            if x_start is not None:
                                ops.scatter(x_start[k:k + 1], tiles[lo_r:hi_r], ex.recv_x0[r, k, :hi_r - lo_r], tile)
            if parity == 1:
                # outside the hull of the shifted grid the state is replaced by fresh noise at the next noise level
                # (q_sample of zeros, model.py:3392-3396); the draw covers the whole canvas like the reference's
                if fresh is None:
                    fresh = ops.randn((1,) + tuple(img.shape[1:]), dev)
                sig = ops.sigma(steps[i + 1])
                for k in range(n_img):
                    ops.renoise_outside(img[k:k + 1], fresh, sig, plan.inner)
            if on_step is not None:
                on_step(i, img, x_start)
    finally:
        if prev_invariant is not None:
            ops.set_batch_invariant(prev_invariant)
    return img, x_start
