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


def road_frame(h: int, w: int, seed: int, shift=(0, 0), margin: int = 8) -> np.ndarray:
    """A frame with road-scene statistics (a SECOND workload beside the SURVEY recipe above, whose 4-px noise texture makes
    ~20 % of all pixels FAST corners): smooth low-frequency shading, a horizon, a few dozen low-contrast facades / vehicles,
    bright lane markings, sensor noise N(0, 1.2).  A few percent of the pixels are corners and many cells need the
    minThFAST fallback, like camera footage."""
    rng = np.random.default_rng(seed * 31 + 7)
    H, W = h + 2 * margin, w + 2 * margin
    coarse = rng.uniform(60, 170, size=(H // 64 + 3, W // 64 + 3))
    ys, xs = np.arange(H) / 64.0, np.arange(W) / 64.0
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
    im = (coarse[y0][:, x0] * (1 - fx) + coarse[y0][:, x0 + 1] * fx) * (1 - fy) + (coarse[y0 + 1][:, x0] * (1 - fx) + coarse[y0 + 1][:, x0 + 1] * fx) * fy
    hz = int(H * rng.uniform(0.35, 0.5))
    im[hz:] = im[hz:] * 0.6 + 25                                   # darker road below the horizon
    for _ in range(60):                                            # facades, vehicles, signs above / around the horizon
        rw, rh = int(rng.integers(10, W // 6)), int(rng.integers(8, H // 5))
        x, y = int(rng.integers(0, W - rw)), int(rng.integers(0, max(1, hz + H // 8 - rh)))
        im[y:y + rh, x:x + rw] = im[y:y + rh, x:x + rw] * 0.3 + rng.uniform(20, 220) * 0.7
        if rng.random() < 0.5:                                     # windows
            for wy in range(y + 3, y + rh - 6, 9):
                for wx in range(x + 3, x + rw - 6, 11):
                    im[wy:wy + 4, wx:wx + 6] -= rng.uniform(15, 60)
    for k in range(int(rng.integers(3, 7))):                       # lane markings: bright dashes converging to the horizon
        xb = rng.uniform(0.1, 0.9) * W
        for t in np.arange(0.05, 1.0, 0.12):
            yy = int(hz + t * (H - hz))
            xx = int(W / 2 + (xb - W / 2) * t)
            ww, hh = max(2, int(10 * t)), max(2, int(22 * t))
            im[yy:yy + hh, xx:xx + ww] = rng.uniform(190, 240)
    dx, dy = shift
    win = im[margin - dy:margin - dy + h, margin - dx:margin - dx + w]
    nrng = np.random.default_rng(seed * 7919 + 17 + 13 * (dx + 50) + (dy + 50))
    return np.ascontiguousarray(np.clip(np.rint(win + nrng.normal(0.0, 1.2, size=(h, w))), 0, 255).astype(np.uint8))


def road_batch(n: int, h: int, w: int, seed: int) -> np.ndarray:
    """n consecutive road-like frames (scene changes every 8 frames, window drifts (+2, +1) px per frame)."""
    out = np.empty((n, h, w), np.uint8)
    for i in range(n):
        k = i % 8
        out[i] = road_frame(h, w, seed * 100 + i // 8, (2 * k - 7, k - 4))
    return out

