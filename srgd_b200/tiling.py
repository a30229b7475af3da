"""Integer tile geometry for large-image sampling (reference: model.py:116-179 and the grid setup of
tiled_sample, model.py:3301-3342).  Pure host logic."""
from __future__ import annotations

from typing import List, Tuple

_CANVAS_UNIT = 256   # the reference pads to multiples of 256 regardless of tile_size (model.py:3301)


def _axis_starts(extent: int, tile: int, stride: int) -> List[int]:
    starts = list(range(0, extent - tile + 1, stride))
    if (extent - tile) % stride != 0:
        starts.append(extent - tile)          # last tile is pulled back to end at the border
    return starts


def get_coord_and_pad(height: int, width: int, tile_size: int = 256):
    """Canvas size, position of the image inside it and the F.pad amounts (model.py:116-135)."""
    if height <= tile_size and width <= tile_size:
        canvas_h = canvas_w = tile_size
    else:
        canvas_h = -(-height // tile_size) * tile_size + tile_size
        canvas_w = -(-width // tile_size) * tile_size + tile_size
    left, top = (canvas_w - width) // 2, (canvas_h - height) // 2
    coord = (left, top, left + width, top + height)
    pad = (left, canvas_w - left - width, top, canvas_h - top - height)
    return coord, pad


def get_coords(h: int, w: int, tile_size: int, tile_stride: int, diff: int = 0):
    """Tile rectangles (y0, y1, x0, x1), row-major (model.py:137-150)."""
    return [(y + diff, y + tile_size + diff, x + diff, x + tile_size + diff)
            for y in _axis_starts(h, tile_size, tile_stride) for x in _axis_starts(w, tile_size, tile_stride)]


def get_area(coords, height: int, width: int):
    """Bounding box of a tile list and the padding that restores (height, width) (model.py:152-179)."""
    top = min([height] + [c[0] for c in coords])
    bottom = max([0] + [c[1] for c in coords])
    left = min([width] + [c[2] for c in coords])
    right = max([0] + [c[3] for c in coords])
    pad = (left, width - right, top, height - bottom)
    return (left, top, right, bottom), pad


class TilePlan:
    """Everything tiled_sample needs to know about one (h, w) image: canvas padding, crop window,
    the two alternating tile grids (aligned / shifted by tile/2) and the inner hull outside of which
    condition and state are reset."""

    def __init__(self, h: int, w: int, tile_size: int = 256, tile_stride: int = 256):
        (left, top, right, bottom), pad = get_coord_and_pad(h, w, _CANVAS_UNIT)
        self.canvas_pad: Tuple[int, int, int, int] = pad
        self.canvas_h, self.canvas_w = h + pad[2] + pad[3], w + pad[0] + pad[1]
        self.crop = (top, bottom, left, right)
        H, W = self.canvas_h, self.canvas_w
        aligned = get_coords(H, W, tile_size, tile_size, 0)
        if H <= tile_size and W <= tile_size:
            shifted = get_coords(H, W, tile_size, tile_stride, 0)
        else:
            shifted = get_coords(H - tile_size, W - tile_size, tile_size, tile_stride, tile_size // 2)
        (il, it, ir, ib), _ = get_area(shifted, H, W)
        self.inner = (it, ib, il, ir)
        self.grids: List[List[Tuple[int, int]]] = [[(c[0], c[2]) for c in aligned], [(c[0], c[2]) for c in shifted]]
        self.tile_size = tile_size

    @property
    def disjoint(self) -> bool:
        """True if no two tiles of a grid overlap (tile_stride == tile_size on the 256-aligned canvas).  With a smaller
        stride, or a tile size that does not divide the canvas (the last tile is pulled back to the border), later
        tiles of a step read pixels earlier minibatches of the same step have already advanced (the reference updates
        the canvas in place, model.py:3374-3385): the result then depends on the minibatch partition."""
        # a grid is the product of its row and column starts: two tiles overlap iff they do on both axes, and two tiles
        # of the same column (row) always exist for any pair of row (column) starts
        def spaced(starts):
            u = sorted(set(starts))
            return all(b - a >= self.tile_size for a, b in zip(u, u[1:]))
        return all(spaced([y for y, _ in g]) and spaced([x for _, x in g]) for g in self.grids)

    def tiles_per_image(self, num_steps: int) -> int:
        even = (num_steps + 1) // 2
        return even * len(self.grids[0]) + (num_steps - even) * len(self.grids[1])
