"""Seeded synthetic front / bird-view images (numpy only, no cv2) for tests and bench.

Recipe follows SURVEY.md §8d: low-pass noise texture + 200 filled rectangles + 100 filled discs +
N(0,2) pixel noise, clipped to u8.  It yields roughly 5-10x nFeatures FAST candidates per level and a
healthy share of cells that need the minThFAST fallback, so the octree distribution is always exercised.
The second frame of a pair-in-time is the same scene shifted by (+3,+2) px so frame-to-frame matches exist.
"""
from __future__ import annotations

import zlib

import numpy as np


def _upsample4(base: np.ndarray, h: int, w: int) -> np.ndarray:
    """Separable linear up-sampling of a (h/4, w/4) grid to (h, w) in float64."""
    bh, bw = base.shape
    ys = np.minimum(np.arange(h) / 4.0, bh - 1.0)
    xs = np.minimum(np.arange(w) / 4.0, bw - 1.0)
    y0 = np.floor(ys).astype(np.int64)
    x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, bh - 1)
    x1 = np.minimum(x0 + 1, bw - 1)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    b = base.astype(np.float64)
    top = b[y0][:, x0] * (1 - fx) + b[y0][:, x1] * fx
    bot = b[y1][:, x0] * (1 - fx) + b[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


def scene(h: int, w: int, seed: int, margin: int = 8) -> np.ndarray:
    """Noise-free float scene of size (h+2*margin, w+2*margin); crop windows out of it."""
    rng = np.random.default_rng(seed)
    H, W = h + 2 * margin, w + 2 * margin
    base = rng.integers(0, 256, size=(H // 4 + 2, W // 4 + 2), dtype=np.int64)
    im = _upsample4(base, H, W)
    nrect, ncirc = 200, 100
    rx = rng.integers(0, W, nrect)
    ry = rng.integers(0, H, nrect)
    rw = rng.integers(4, max(5, W // 8), nrect)
    rh = rng.integers(4, max(5, H // 8), nrect)
    rg = rng.integers(0, 256, nrect)
    for i in range(nrect):
        im[ry[i]:ry[i] + rh[i], rx[i]:rx[i] + rw[i]] = rg[i]
    cx = rng.integers(0, W, ncirc)
    cy = rng.integers(0, H, ncirc)
    cr = rng.integers(3, max(4, min(H, W) // 12), ncirc)
    cg = rng.integers(0, 256, ncirc)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(ncirc):
        y0, y1 = max(0, cy[i] - cr[i]), min(H, cy[i] + cr[i] + 1)
        x0, x1 = max(0, cx[i] - cr[i]), min(W, cx[i] + cr[i] + 1)
        m = (yy[y0:y1, x0:x1] - cy[i]) ** 2 + (xx[y0:y1, x0:x1] - cx[i]) ** 2 <= cr[i] ** 2
        im[y0:y1, x0:x1][m] = cg[i]
    return im


def frame(h: int, w: int, seed: int, shift=(0, 0), noise_seed: int | None = None, margin: int = 8) -> np.ndarray:
    """One u8 frame: the seeded scene seen through a window displaced by `shift` = (dx, dy)."""
    sc = scene(h, w, seed, margin)
    dx, dy = shift
    assert abs(dx) <= margin and abs(dy) <= margin
    win = sc[margin - dy:margin - dy + h, margin - dx:margin - dx + w]
    nrng = np.random.default_rng((seed if noise_seed is None else noise_seed) * 7919 + 13)
    out = win + nrng.normal(0.0, 2.0, size=(h, w))
    return np.ascontiguousarray(np.clip(np.rint(out), 0, 255).astype(np.uint8))


def frame_pair_in_time(h: int, w: int, seed: int):
    """(previous, current) frames of one camera: same scene, current shifted by (+3,+2) px."""
    return frame(h, w, seed, (0, 0), noise_seed=2 * seed), frame(h, w, seed, (3, 2), noise_seed=2 * seed + 1)


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def cheap_batch(n: int, h: int, w: int, seed: int) -> np.ndarray:
    """n consecutive frames for throughput runs: a handful of fully synthesised scenes, each viewed by a run of
    consecutive frames through a window that drifts a few pixels per frame with fresh pixel noise (so consecutive frames
    match, as in a real sequence; building thousands of scenes from scratch would dominate the bench set-up)."""
    nscene = min(n, 8)
    run = (n + nscene - 1) // nscene
    out = np.empty((n, h, w), np.uint8)
    rng = np.random.default_rng(seed ^ 0x5EED)
    m = 8
    sc = None
    for i in range(n):
        j = i % run
        if j == 0:
            sc = scene(h, w, seed * 1000 + i // run)
        k = j % 9                       # drift (+2,+1) px per frame, back and forth inside the +-8 px margin
        dx, dy = (2 * k - 8 if k <= 8 else 0), (k - 4)
        win = sc[m - dy:m - dy + h, m - dx:m - dx + w]
        out[i] = np.clip(np.rint(win + rng.normal(0.0, 2.0, size=(h, w))), 0, 255).astype(np.uint8)
    return out
