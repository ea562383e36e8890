"""Seeded inputs of the bird-view guidance / refinement row, numpy only so they travel to the GPU box."""
import numpy as np

from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200._lib import KP_DTYPE

EDGE = 31          # cv::ORB's default edgeThreshold: detected keypoints keep this distance from the border


def bird_image(seed: int, rows: int = 384, cols: int = 384) -> np.ndarray:
    return np.ascontiguousarray(synth.frame(rows, cols, 7000 + seed))


def contour_image(seed: int, rows: int = 384, cols: int = 384) -> np.ndarray:
    """mBirdviewContourICP stand-in: 0 = free, 100 = edge (outlines), 255 = free space (filled areas), mostly empty."""
    rng = np.random.default_rng(8000 + seed)
    c = np.zeros((rows, cols), np.uint8)
    for _ in range(6):
        y0, x0 = int(rng.integers(0, rows - 40)), int(rng.integers(0, cols - 40))
        h, w = int(rng.integers(10, 60)), int(rng.integers(10, 60))
        c[y0:y0 + h, x0:x0 + w] = 255
        c[y0, x0:x0 + w] = 100; c[min(y0 + h, rows) - 1, x0:x0 + w] = 100
        c[y0:y0 + h, x0] = 100; c[y0:y0 + h, min(x0 + w, cols) - 1] = 100
    for _ in range(12):                                   # isolated low-valued pixels: 9 is "free", 10 already counts
        c[int(rng.integers(0, rows)), int(rng.integers(0, cols))] = int(rng.choice([5, 9, 10, 149, 150]))
    return c


def corner_points(img: np.ndarray, seed: int, n: int = 1500, edge: int = EDGE) -> np.ndarray:
    """Corner-like integer locations (largest min-eigenvalue-free proxy: local gradient energy) at least `edge` px inside, plus
    random sub-pixel points; float32 [n, 2] as (x, y)."""
    rng = np.random.default_rng(9000 + seed)
    f = img.astype(np.float32)
    gx = np.abs(f[1:-1, 2:] - f[1:-1, :-2]); gy = np.abs(f[2:, 1:-1] - f[:-2, 1:-1])
    e = np.zeros_like(f); e[1:-1, 1:-1] = np.minimum(gx, gy)
    e[:edge] = 0; e[-edge:] = 0; e[:, :edge] = 0; e[:, -edge:] = 0
    order = np.argsort(-e, axis=None, kind="stable")[: n * 2 // 3]
    ys, xs = np.unravel_index(order, e.shape)
    strong = np.stack([xs, ys], 1).astype(np.float32)
    rows, cols = img.shape
    rnd = np.stack([rng.uniform(edge, cols - 1 - edge, n - len(strong)), rng.uniform(edge, rows - 1 - edge, n - len(strong))], 1)
    return np.ascontiguousarray(np.concatenate([strong, rnd.astype(np.float32)]))


def border_points(rows: int, cols: int, seed: int, n: int = 300) -> np.ndarray:
    """Points whose sampled windows cross the image border (never produced by cv::ORB; exercised GPU vs oracle only)."""
    rng = np.random.default_rng(9500 + seed)
    p = np.stack([rng.uniform(0, cols - 1, n), rng.uniform(0, rows - 1, n)], 1).astype(np.float32)
    side = rng.integers(0, 4, n)
    p[side == 0, 0] = rng.uniform(0, 8, (side == 0).sum()); p[side == 1, 0] = rng.uniform(cols - 9, cols - 1, (side == 1).sum())
    p[side == 2, 1] = rng.uniform(0, 8, (side == 2).sum()); p[side == 3, 1] = rng.uniform(rows - 9, rows - 1, (side == 3).sum())
    fixed = np.float32([[0, 0], [cols - 1, rows - 1], [0, rows - 1], [cols - 1, 0], [2.5, rows - 3.8], [cols - 2.3, 1.2]])
    return np.ascontiguousarray(np.concatenate([p, fixed]))


def as_kps(xy: np.ndarray) -> np.ndarray:
    k = np.zeros(len(xy), KP_DTYPE)
    k["x"], k["y"] = xy[:, 0], xy[:, 1]
    k["size"], k["angle"], k["octave"], k["class_id"] = 31, -1, 0, -1
    k["response"] = np.arange(len(xy), dtype=np.float32)          # a tag: the compaction must keep records whole and in order
    return k


def bird_mask(seed: int, rows: int = 384, cols: int = 384):
    """mBirdviewMask stand-ins: seed 0 -> no mask, 1 -> a vehicle-shaped hole (0) in 255 with a few non-binary values, 2 -> left half."""
    if seed % 3 == 0:
        return None
    m = np.full((rows, cols), 255, np.uint8)
    if seed % 3 == 1:
        m[rows // 2 - 70:rows // 2 + 70, cols // 2 - 35:cols // 2 + 35] = 0
        m[:40, :] = 7                       # non-zero, non-255: cv2 4.13 keeps these on every level
        m[rows - 30:, cols - 120:] = 0
    else:
        m[:, cols // 2 + 8:] = 0
    return m
