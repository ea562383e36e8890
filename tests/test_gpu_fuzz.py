"""Randomised geometry sweep of the extractor (GPU vs oracle): odd widths and heights, levels that shrink below one FAST
cell, wide / tall aspect ratios, every scale factor, feature budgets far below and above what the image offers.  The
tile / strip / TMA-box geometry of the kernels is derived from these numbers, so this is where an off-by-one would hide."""
import numpy as np
import pytest

from fishbirdeyevisualslam_b200 import synth
from test_gpu_extract import compare

pytestmark = pytest.mark.gpu


def configs(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        h, w = int(rng.integers(64, 430)), int(rng.integers(64, 700))
        scale = float(rng.choice([1.1, 1.2, 1.25, 1.33, 1.5, 2.0]))
        nl = int(rng.integers(1, 9))
        # keep what the reference itself can run: every level needs at least one 30-px FAST cell in its detection band
        # [16, size-16) (nCols = width/30 divides, ORBextractor.cc:785-788) and at least one octree root
        # (nIni = round(W/H) != 0, :543-545); the library reports both as FBE_E_UNSUPPORTED
        ok = True
        for l in range(nl):
            hh, ww = h / scale ** l, w / scale ** l
            if hh - 32 < 31.5 or ww - 32 < 31.5 or (ww - 32) / (hh - 32) < 0.55:
                ok = False
        if not ok:
            continue
        nf = int(rng.choice([30, 100, 400, 1000, 2500]))
        ini, mn = int(rng.choice([7, 15, 20, 40])), int(rng.choice([2, 5, 7]))
        out.append((h, w, nf, scale, nl, ini, min(mn, ini), int(rng.integers(0, 10 ** 6)), bool(rng.random() < 0.3)))
    return out


@pytest.mark.parametrize("cfg", configs(48, 2024))
def test_random_geometry(oracle, cfg):
    from fishbirdeyevisualslam_b200.extractor import ORBextractor
    h, w, nf, scale, nl, ini, mn, seed, road = cfg
    img = synth.road_frame(h, w, seed) if road else synth.frame(h, w, seed % 1000)
    o = oracle.OracleExtractor(nf, scale, nl, ini, mn)
    ko, do = o(img)
    g = ORBextractor(nf, scale, nl, ini, mn)
    kg, dg = g(img)
    for l in range(nl):
        assert np.array_equal(g.pyramid_level(l), o.level_padded(l)), f"pyramid level {l}"
        assert np.array_equal(g.debug_candidates(l), o.candidates(l)), f"candidates level {l}"
    compare(kg, dg, ko, do, o.boundary)


@pytest.mark.parametrize("shape", [(406, 98, 1), (67, 561, 2), (300, 40, 1)])
def test_configurations_the_reference_cannot_run(shape):
    """Tall images give round(W/H) = 0 octree roots (division by zero at ORBextractor.cc:545), bands narrower than one FAST
    cell give nCols = 0 (:785-788): the library refuses them with an error instead of inventing a behaviour."""
    from fishbirdeyevisualslam_b200._lib import FbeError
    from fishbirdeyevisualslam_b200.extractor import ORBextractor
    h, w, nl = shape
    with pytest.raises(FbeError):
        ORBextractor(100, 1.1, nl, 15, 5)(synth.frame(h, w, 1))

