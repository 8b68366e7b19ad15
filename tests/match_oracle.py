"""CPU restatement of the reference's brute-force descriptor matcher (test infrastructure for SURVEY.md 8f
rank 2; the product does not contain a matcher yet).

Reference: `FeaturesDev::match` -> `compute_distance` (src/popsift/features.cu:165-222,267-304).  For every
left descriptor l the kernel walks all right descriptors r in index order, computes the squared L2 distance
in float32 (a warp: each lane squares the four differences of its float4, `x*x + y*y + z*z + w*w`, then a
shuffle-down tree 16, 8, 4, 2, 1), keeps the best and second-best with strict `<` comparisons (ties keep the
earlier index), and accepts the match when `best / second < 0.8f` (ratio of SQUARED distances).
Result per left descriptor: (best index, second index, accept)."""
from __future__ import annotations

import numpy as np


def _warp_sq_dist(l: np.ndarray, r: np.ndarray) -> np.float32:
    """squared distance of two 128-float descriptors with the reference's float32 summation order"""
    f32 = np.float32
    d = (l.astype(f32) - r.astype(f32)).reshape(32, 4)
    lane = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(f32) + d[:, 2] * d[:, 2]).astype(f32)
    lane = (lane + d[:, 3] * d[:, 3]).astype(f32)
    for s in (16, 8, 4, 2, 1):                 # shuffle_down tree; lane 0 ends with the total
        up = np.concatenate([lane[s:], lane[:s]])
        lane = (lane + up).astype(f32)
    return lane[0]


def sq_dist_matrix(left: np.ndarray, right: np.ndarray, fused: bool = True) -> np.ndarray:
    """all-pairs squared distances with the reference's float32 summation order, vectorised.
    fused=True evaluates the per-lane sum as the reference's SASS does (nvcc contracts
    x*x + y*y + z*z + w*w into FMUL, FFMA, FFMA, FFMA: products of the fused steps are not rounded)."""
    f32 = np.float32
    out = np.empty((len(left), len(right)), f32)
    R = right.astype(f32).reshape(len(right), 32, 4)
    for i in range(len(left)):
        d = (left[i].astype(f32).reshape(1, 32, 4) - R).astype(f32)
        if fused:
            d64 = d.astype(np.float64)
            lane = (d[..., 0] * d[..., 0]).astype(f32)
            for k in (1, 2, 3):                     # fma(d_k, d_k, lane): exact product, one rounding
                lane = (d64[..., k] * d64[..., k] + lane.astype(np.float64)).astype(f32)
        else:
            lane = ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(f32) + d[..., 2] * d[..., 2]).astype(f32)
            lane = (lane + d[..., 3] * d[..., 3]).astype(f32)
        for s in (16, 8, 4, 2, 1):
            lane = (lane + np.concatenate([lane[:, s:], lane[:, :s]], axis=1)).astype(f32)
        out[i] = lane[:, 0]
    return out


def best_two(dist: np.ndarray) -> np.ndarray:
    """(best, second, accept) per row with the reference's scan: strict `<`, ties keep the earlier index"""
    nl, nr = dist.shape
    out = np.zeros((nl, 3), np.int32)
    for i in range(nl):
        row = dist[i]
        i1 = int(np.argmin(row))                    # first minimum
        b1 = row[i1]
        rest = row.copy(); rest[i1] = np.inf
        i2 = int(np.argmin(rest)) if nr > 1 else 0
        b2 = rest[i2] if nr > 1 else np.float32(np.inf)
        with np.errstate(divide="ignore", invalid="ignore"):
            accept = bool(np.float32(b1) / np.float32(b2) < np.float32(0.8))
        out[i] = (i1, i2, int(accept))
    return out


def match(left: np.ndarray, right: np.ndarray, exact_order: bool = True) -> np.ndarray:
    """(n_left, 3) int32 rows (best, second, accept).  exact_order=False uses a float64 matrix product --
    the form a tensor-core kernel will compute -- and is only equal up to ties / rounding."""
    nl, nr = len(left), len(right)
    out = np.zeros((nl, 3), np.int32)
    if nr == 0:
        return out
    if exact_order:
        dist = np.empty((nl, nr), np.float32)
        for i in range(nl):
            for j in range(nr):
                dist[i, j] = _warp_sq_dist(left[i], right[j])
    else:
        a = left.astype(np.float64); b = right.astype(np.float64)
        dist = ((a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * a @ b.T).astype(np.float32)
    for i in range(nl):
        b1 = b2 = np.float32(np.inf)
        i1 = i2 = 0
        for j in range(nr):
            v = dist[i, j]
            if v < b1:
                b2, i2, b1, i1 = b1, i1, v, j
            elif v < b2:
                b2, i2 = v, j
        with np.errstate(divide="ignore", invalid="ignore"):
            accept = bool(np.float32(b1) / np.float32(b2) < np.float32(0.8))
        out[i] = (i1, i2, int(accept))
    return out
