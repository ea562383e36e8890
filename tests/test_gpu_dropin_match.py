"""The C++ drop-in matcher as a maintainer would build it -- the reference's OWN ORBmatcher class (its headers, Frame /
KeyFrame / MapPoint objects, its remaining method bodies) with fishbirdeyevisualslam_b200/host/ORBmatcher_fbe.cc providing
the hot-path bodies on top of libfbe_b200.so (oracle/Makefile target `dropinmatch`, prebuilt in the build container and
shipped in oracle/_ref) -- against the outputs of the reference's verbatim CPU build on the same scenes
(tests/golden/match.npz).  Pointer-level semantics (NULL after an orientation prune, map-point identity, DMatch lists)
are therefore compared exactly as a caller of the reference sees them."""
import os

import numpy as np
import pytest

import test_oracle_vs_refmatch as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dropin(oracle):
    lib = oracle.dropinmatch()
    if lib is None:
        pytest.skip("oracle/_ref/libfbe_dropinmatch.so not built (needs the reference sources at build time)")
    return oracle.RefMatch(lib)


@pytest.mark.parametrize("seed", T.SEEDS)
def test_dropin_class_equals_reference_outputs(dropin, seed):
    g = np.load(T.GOLD)
    got = T.scene_outputs(dropin, seed, True)
    for k, v in got.items():
        assert np.array_equal(v, g[f"s{seed}_{k}"], equal_nan=True), k


def test_dropin_descriptor_distance(dropin):
    rng = np.random.default_rng(5)
    a, b = rng.integers(0, 256, (64, 32), dtype=np.uint8), rng.integers(0, 256, (64, 32), dtype=np.uint8)
    for x, y in zip(a, b):
        assert dropin.hamming256(x, y) == int(np.unpackbits(x ^ y).sum())


@pytest.mark.parametrize("seed", range(3, 13))
def test_dropin_class_equals_verbatim_reference_live(oracle, dropin, seed):
    """Seeds without committed outputs: the verbatim CPU build (shipped prebuilt in oracle/_ref) runs beside the drop-in class."""
    if oracle.refmatch() is None:
        pytest.skip("oracle/_ref/libfbe_refmatch.so not built")
    want = T.scene_outputs(oracle.RefMatch(), seed, True)
    got = T.scene_outputs(dropin, seed, True)
    for k, v in want.items():
        assert np.array_equal(got[k], v, equal_nan=True), k

