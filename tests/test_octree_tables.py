"""The identity k_octree's histogram pass relies on (csrc/octree.cu: descend(), descend_axis(), xtab / ytab): the walk of a key down
the data-independent quad-tree geometry of DistributeOctTree / DivideNode (src/ORBextractor.cc:481-537, 543-570 of the reference)
takes its x and y decisions independently, so the depth-D path of (x, y) under root r is

    (r << 2D) + xdigits(x, root r's x-interval) + ydigits(y, [0, H))

with one base-4 digit per depth whose bit 0 comes from x and bit 1 from y -- and a key's path at depth d < D is that value >> 2(D - d).
Plain-Python restatement of both device functions, checked against each other on random level geometries (no GPU needed; the CUDA
kernel itself is pinned by the -m gpu extraction and octree tests)."""
import numpy as np
import pytest


def f32(v):
    return np.float32(v)


def descend_paths(x, y, r, hx, H, D):
    """descend(): the path visited at every depth 0..D (root bounds from hX * r like the reference, :555-558)."""
    x0, x1, y0, y1 = int(f32(hx) * f32(r)), int(f32(hx) * f32(r + 1)), 0, H
    path, out = r, [r]
    for _ in range(D):
        mx, my = x0 + ((x1 - x0 + 1) >> 1), y0 + ((y1 - y0 + 1) >> 1)       # UL + ceil(extent / 2) (:483-484)
        q = 0
        if x < mx:
            x1 = mx
        else:
            x0 = mx; q = 1
        if y < my:
            y1 = my
        else:
            y0 = my; q += 2
        path = 4 * path + q
        out.append(path)
    return out


def descend_axis(v, lo, hi, D, bit):
    path = 0
    for _ in range(D):
        m = lo + ((hi - lo + 1) >> 1)
        q = 0
        if v < m:
            hi = m
        else:
            lo = m; q = bit
        path = 4 * path + q
    return path


def root_of(x, hx, nini):
    return min(max(int(f32(x) / f32(hx)), 0), nini - 1)


@pytest.mark.parametrize("seed", range(6))
def test_axis_tables_equal_the_two_axis_walk(seed):
    rng = np.random.default_rng(seed)
    for _ in range(40):
        W, H = int(rng.integers(40, 1400)), int(rng.integers(40, 1400))
        nini = max(1, int(round(W / H)))
        hx = f32(W) / f32(nini)
        D = int(rng.integers(1, 7))
        xtab = [(root_of(x, hx, nini) << (2 * D)) + descend_axis(x, int(f32(hx) * f32(root_of(x, hx, nini))),
                                                                 int(f32(hx) * f32(root_of(x, hx, nini) + 1)), D, 1) for x in range(W)]
        ytab = [descend_axis(y, 0, H, D, 2) for y in range(H)]
        xs, ys = rng.integers(0, W, 300), rng.integers(0, H, 300)
        # the edges of every interval matter most: add the columns / rows next to each split of the first root
        for x, y in zip(list(xs) + [0, W - 1, W // 2, W // 2 - 1], list(ys) + [0, H - 1, H // 2, H // 2 - 1]):
            x, y = int(x), int(y)
            paths = descend_paths(x, y, root_of(x, hx, nini), hx, H, D)
            leaf = xtab[x] + ytab[y]
            assert leaf == paths[D]
            for d in range(D + 1):
                assert leaf >> (2 * (D - d)) == paths[d]


def test_all_pixels_of_a_small_level():
    W, H, D = 97, 61, 4
    nini = max(1, int(round(W / H)))
    hx = f32(W) / f32(nini)
    for x in range(W):
        r = root_of(x, hx, nini)
        xt = (r << (2 * D)) + descend_axis(x, int(f32(hx) * f32(r)), int(f32(hx) * f32(r + 1)), D, 1)
        for y in range(H):
            assert xt + descend_axis(y, 0, H, D, 2) == descend_paths(x, y, r, hx, H, D)[D]
