"""Line rasteriser behind the reference's class API (`import bresenham as drawing`).

bresenham(start, end).path is the list of (x, y) cells of W12 course_agv_slam/scripts/
bresenham.py:2-58, traced by the CUDA library (float64 error accumulator, canonical
direction).  `paths(segs)` rasterises many segments in one launch.
"""
import numpy as np

from b2slam import _lib


def paths(segs):
    """segs (S,4) int = x0,y0,x1,y1 -> (cells (total,2) int32, offsets (S+1,) int64)."""
    segs = np.ascontiguousarray(segs, dtype=np.int32).reshape(-1, 4)
    dx = np.abs(segs[:, 2].astype(np.int64) - segs[:, 0])
    dy = np.abs(segs[:, 3].astype(np.int64) - segs[:, 1])
    length = np.maximum(dx, dy) + 1
    length[(dx == 0) & (dy == 0)] = 0  # same cell -> empty path ([BRES]:10-11)
    offsets = np.zeros(segs.shape[0] + 1, dtype=np.int64)
    np.cumsum(length, out=offsets[1:])
    cells = np.empty((int(offsets[-1]), 2), dtype=np.int32)
    _lib.require_device()
    _lib.check(_lib.lib().b2s_bresenham_host(_lib.ptr(segs), segs.shape[0], _lib.ptr(offsets),
                                             _lib.ptr(cells) if cells.size else None))
    return cells, offsets


class bresenham(object):
    def __init__(self, start, end):
        self.start = start
        self.end = end
        cells, _ = paths([[int(start[0]), int(start[1]), int(end[0]), int(end[1])]])
        self.path = [tuple(int(v) for v in c) for c in cells]
