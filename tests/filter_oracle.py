"""CPU restatement of the reference's grid filter (test infrastructure).

Reference: Pyramid::extrema_filter_grid (src/popsift/s_filtergrid.cu:112-325), called from Pyramid::orientation
(s_orientation.cu:380-383) when filter_max_extrema > 0 and int(1.1 * filter_max_extrema) < number of extrema.
Every extremum carries a grid cell (s_extrema.cu:499: floor(y / (H_o / g)) * g + floor(x / (W_o / g)), octave-local
coordinates) and a scale sigma * 2^octave.  The filter
  1. counts the extrema per cell,
  2. finds the per-cell limit: with the counts sorted ascending c[0..n), sumup[i] = prefix[i] + c[i] * (n-1-i) is the
     total if every larger cell were cut down to c[i]; ct = #{i : sumup[i] > max}; the limit is
     ceil( mean(largest ct counts) - (total - max) / ct )   with (total - max) / ct an INTEGER division,
  3. keeps, in every cell, the first `limit` extrema in (scale ascending | scale descending | arrival) order.
Quirk kept: the per-cell counts come from a reduce_by_key over the sorted cell ids, so an EMPTY cell shifts the
counts of all later cells by one position (s_filtergrid.cu:190-196); `counts_as_reference` reproduces that."""
from __future__ import annotations

import math

import numpy as np


def cell_of(x_oct, y_oct, w_oct, h_oct, grid):
    f32 = np.float32
    wd, hd = f32(w_oct) / f32(grid), f32(h_oct) / f32(grid)
    return (np.floor(f32(y_oct) / hd) * f32(grid) + np.floor(f32(x_oct) / wd)).astype(np.int64)


def limit_for(counts_sorted_asc, total, max_extrema):
    n = len(counts_sorted_asc)
    c = np.asarray(counts_sorted_asc, dtype=np.int64)
    prefix = np.cumsum(c)
    sumup = prefix + c * (n - 1 - np.arange(n))
    ct = int((sumup > max_extrema).sum())
    if ct == 0:
        return None
    tail = np.float32(c[n - ct:].sum()) / np.float32(ct)
    return int(math.ceil(float(np.float32(tail - np.float32((total - max_extrema) // ct)))))


def keep_mask(cells, scales, grid, max_extrema, mode):
    """boolean mask of the extrema the reference keeps; mode in {"up", "down", "random"} (random: arrival order is
    not reproducible -- only the per-cell counts are)"""
    n = len(cells)
    keep = np.ones(n, bool)
    if max_extrema <= 0 or int(max_extrema * 1.1) >= n:
        return keep, None
    ncell = grid * grid
    present = np.unique(cells)
    counts = np.array([(cells == c).sum() for c in present], dtype=np.int64)           # reduce_by_key: non-empty cells only
    padded = np.concatenate([counts, np.zeros(ncell - len(counts), np.int64)])        # device vector of g*g entries
    limit = limit_for(np.sort(padded), n, max_extrema)
    if limit is None:
        return keep, None
    for c in present:
        idx = np.nonzero(cells == c)[0]
        if len(idx) <= limit:
            continue
        if mode == "up":
            order = idx[np.argsort(scales[idx], kind="stable")]
        elif mode == "down":
            order = idx[np.argsort(-scales[idx], kind="stable")]
        else:
            order = idx
        keep[order[limit:]] = False
    return keep, limit
