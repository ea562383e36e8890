"""Small synthetic frames for matcher tests (no extraction needed): clustered keypoints with correlated descriptors."""
import numpy as np

from conftest import make_kps
from fishbirdeyevisualslam_b200.matcher import Frame


def flip_bits(rng, d, nmax):
    d = d.copy()
    for i in range(len(d)):
        k = int(rng.integers(0, nmax + 1))
        if k:
            bits = rng.choice(256, k, replace=False)
            for b in bits:
                d[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return d


def frame_pair(rng, n=300, w=640, h=480, bird=False, max_octave=3, shift=(3.0, 2.0), flips=30, dup=0.15):
    """Two frames of the same scene: F2 = F1 shifted, descriptors perturbed, some points duplicated nearby (ambiguity)."""
    x = rng.uniform(5, w - 5, n).astype(np.float32)
    y = rng.uniform(5, h - 5, n).astype(np.float32)
    # quantise a part of the coordinates so that exact ties / cell-boundary hits occur
    q = rng.random(n) < 0.3
    x[q] = np.round(x[q] / 10) * 10
    y[q] = np.round(y[q] / 7.5) * 7.5
    octv = rng.integers(0, max_octave + 1, n).astype(np.int32)
    octv[rng.random(n) < 0.5] = 0
    ang = (rng.uniform(0, 360, n)).astype(np.float32)
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    k1 = make_kps(x, y, octv, ang)
    # second frame: permuted, shifted, rotated a little, with duplicates
    perm = rng.permutation(n)
    x2 = x[perm] + np.float32(shift[0]) + rng.normal(0, 0.7, n).astype(np.float32)
    y2 = y[perm] + np.float32(shift[1]) + rng.normal(0, 0.7, n).astype(np.float32)
    ang2 = np.mod(ang[perm] + rng.choice([0, 0, 0, 29, 95, 200], n).astype(np.float32) + rng.normal(0, 3, n).astype(np.float32), 360).astype(np.float32)
    d2 = flip_bits(rng, d1[perm], flips)
    nd = int(dup * n)
    src = rng.integers(0, n, nd)
    x2 = np.concatenate([x2, x2[src] + rng.normal(0, 4, nd).astype(np.float32)])
    y2 = np.concatenate([y2, y2[src] + rng.normal(0, 4, nd).astype(np.float32)])
    ang2 = np.concatenate([ang2, ang2[src]])
    o2 = np.concatenate([octv[perm], octv[perm][src]])
    d2 = np.concatenate([d2, flip_bits(rng, d2[src], 6)])
    k2 = make_kps(x2, y2, o2, ang2)
    if bird:
        return Frame.bird(k1, d1, w, h), Frame.bird(k2, d2, w, h)
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    return Frame.front(k1, d1, w, h, sf), Frame.front(k2, d2, w, h, sf)


def featvec(node_of):
    ids = np.unique(node_of)
    start, items = [0], []
    for t in ids:
        it = np.nonzero(node_of == t)[0]
        items.extend(it.tolist())
        start.append(len(items))
    return ids.astype(np.int32), np.array(start, np.int32), np.array(items, np.int32)
