"""Deterministic synthetic grayscale frames.

Broadband texture (4 octaves of bilinearly up-sampled uniform noise) plus filled
rectangles / discs so that keypoint density is realistic.  Same seed -> same bytes.
"""
from __future__ import annotations

import numpy as np

SEEDS = {"640x480": 1, "3840x2160": 7}  # 1080p stream: 100 + i


def _upsample_bilinear(a: np.ndarray, h: int, w: int) -> np.ndarray:
    """Bilinear resize of a small 2-D array to (h, w), align-corners style."""
    sh, sw = a.shape
    ys = np.linspace(0.0, sh - 1.0, h)
    xs = np.linspace(0.0, sw - 1.0, w)
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, sh - 1)
    x1 = np.minimum(x0 + 1, sw - 1)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    top = a[y0][:, x0] * (1.0 - fx) + a[y0][:, x1] * fx
    bot = a[y1][:, x0] * (1.0 - fx) + a[y1][:, x1] * fx
    return top * (1.0 - fy) + bot * fy


def make_frame(width: int, height: int, seed: int) -> np.ndarray:
    """Return a (height, width) uint8 frame."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((height, width), dtype=np.float64)
    for div in (64, 32, 16, 8):
        sh, sw = max(2, height // div), max(2, width // div)
        acc += _upsample_bilinear(rng.uniform(-1.0, 1.0, size=(sh, sw)), height, width)
    img = 128.0 + 48.0 * acc / 2.0
    n_shapes = max(4, int(200 * width * height / 1e6))
    for _ in range(n_shapes):
        amp = 24.0 * (1.0 if rng.random() < 0.5 else -1.0) * rng.uniform(0.5, 1.5)
        cx = int(rng.integers(0, width))
        cy = int(rng.integers(0, height))
        if rng.random() < 0.5:
            hw = int(rng.integers(2, max(3, width // 40)))
            hh = int(rng.integers(2, max(3, height // 40)))
            img[max(0, cy - hh):cy + hh + 1, max(0, cx - hw):cx + hw + 1] += amp
        else:
            r = int(rng.integers(2, max(3, min(width, height) // 40)))
            y0, y1 = max(0, cy - r), min(height, cy + r + 1)
            x0, x1 = max(0, cx - r), min(width, cx + r + 1)
            ly, lx = np.mgrid[y0:y1, x0:x1]
            img[y0:y1, x0:x1] += amp * (((ly - cy) ** 2 + (lx - cx) ** 2) <= r * r)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def affine_warp(img: np.ndarray, A: np.ndarray) -> np.ndarray:
    """Warp a uint8 image with the 2x3 affine map A (source -> destination pixel coordinates): the
    destination has the source's size, every destination pixel is the bilinear sample of the source
    at A^-1 (x, y), clamp-to-edge.  Pure numpy float64, so the same bytes everywhere."""
    h, w = img.shape
    M = np.vstack([np.asarray(A, dtype=np.float64), [0.0, 0.0, 1.0]])
    Mi = np.linalg.inv(M)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    sx = Mi[0, 0] * xs + Mi[0, 1] * ys + Mi[0, 2]
    sy = Mi[1, 0] * xs + Mi[1, 1] * ys + Mi[1, 2]
    sx = np.clip(sx, 0.0, w - 1.0)
    sy = np.clip(sy, 0.0, h - 1.0)
    x0 = np.floor(sx).astype(np.int64)
    y0 = np.floor(sy).astype(np.int64)
    x1 = np.minimum(x0 + 1, w - 1)
    y1 = np.minimum(y0 + 1, h - 1)
    fx, fy = sx - x0, sy - y0
    f = img.astype(np.float64)
    top = f[y0, x0] * (1.0 - fx) + f[y0, x1] * fx
    bot = f[y1, x0] * (1.0 - fx) + f[y1, x1] * fx
    return np.clip(np.rint(top * (1.0 - fy) + bot * fy), 0, 255).astype(np.uint8)


def affine_set(width: int = 800, height: int = 640, seed: int = 31):
    """Stand-in for the Oxford affine-covariant sets (BASELINE.json configs[4]; the PGMs are not
    available offline): image 1 = make_frame(seed), images 2..6 = image 1 under known
    affine maps about the image centre -- "boat"-like zoom + rotation (k = 1..3) and "graffiti"-like
    viewpoint shear (k = 4, 5).  Returns [(image, A_k)] with A_1 = identity; A maps image-1 pixel
    coordinates to image-k coordinates."""
    base = make_frame(width, height, seed)
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    out = [(base, np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]))]
    params = [(1.12, 10.0, 0.0), (1.30, 25.0, 0.0), (0.80, -40.0, 0.0), (1.0, 0.0, 0.25), (0.9, 15.0, 0.45)]
    for scale, deg, shear in params:
        th = np.deg2rad(deg)
        R = scale * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        S = np.array([[1.0, shear], [0.0, 1.0 / (1.0 + shear)]])      # foreshortening along y with a shear in x
        L = R @ S
        t = np.array([cx, cy]) - L @ np.array([cx, cy])
        A = np.hstack([L, t[:, None]])
        out.append((affine_warp(base, A), A))
    return out


def write_pgm(path: str, img: np.ndarray) -> None:
    """Binary PGM (P5, maxval 255) as read by the reference's pgmread.cpp:180-197."""
    assert img.dtype == np.uint8 and img.ndim == 2
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


def read_pgm(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        data = f.read()
    toks, pos = [], 0
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        toks.append(data[pos:end])
        pos = end
    assert toks[0] == b"P5" and int(toks[3]) == 255
    w, h = int(toks[1]), int(toks[2])
    pos += 1
    return np.frombuffer(data, dtype=np.uint8, count=w * h, offset=pos).reshape(h, w).copy()


if __name__ == "__main__":
    import sys
    if sys.argv[1] == "affine":          # python -m popsift_b200.synth affine OUT_PREFIX  -> OUT_PREFIX1..6.pgm
        for k, (im, _A) in enumerate(affine_set(), 1):
            write_pgm("%s%d.pgm" % (sys.argv[2], k), im)
    else:
        w, h, seed, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
        write_pgm(out, make_frame(w, h, seed))
