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


def test_dropin_frame_undistort_keypoints(oracle, dropin):
    """Frame::UndistortKeyPoints through host/Frame_fbe.cc on a real Frame object: mvKeysUn equals the oracle (which is pinned
    to cv2's fisheye.undistortPoints); k1 == 0 copies the keypoints."""
    import ctypes as C
    from conftest import make_kps
    rng = np.random.default_rng(12)
    n = 3000
    kps = make_kps(rng.uniform(0, 960, n).astype(np.float32), rng.uniform(0, 600, n).astype(np.float32),
                   rng.integers(0, 8, n).astype(np.int32), rng.uniform(0, 360, n).astype(np.float32))
    K = np.float32([348.5, 347.0, 480.0, 302.0])
    for D in (np.float32([-0.0488316, 0.000298406, -0.00591118, 0.00193258]), np.float32([0, 0.1, 0, 0])):
        out = np.empty_like(kps)
        dropin.L.refm_undistort_keypoints(kps.ctypes.data_as(C.c_void_p), n, K.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p),
                                          out.ctypes.data_as(C.c_void_p))
        if D[0] == 0:
            assert out.tobytes() == kps.tobytes()
            continue
        want = oracle.fisheye_undistort(np.stack([kps["x"], kps["y"]], 1), K, D)
        assert np.array_equal(out["x"], want[:, 0]) and np.array_equal(out["y"], want[:, 1])
        for f in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(out[f], kps[f])



def test_dropin_frame_bird_feature_block(oracle, dropin):
    """The bird feature block of the Frame constructor (src/Frame.cc:336-355) through host/Frame_fbe.cc's FbeBirdFeatures on
    cv::Mat inputs: mvKeysBird (order included) and mDescriptorsBird equal the oracle chain cv::ORB detect -> nearEdges ->
    cornerSubPix -> cv::ORB compute, which is pinned to cv2 4.13."""
    import ctypes as C
    import bird_scenes as S
    from fishbirdeyevisualslam_b200._lib import KP_DTYPE
    for i, with_mask, with_contour in ((0, False, True), (1, True, True), (2, True, False)):
        img = S.bird_image(i)
        mask = S.bird_mask(i) if with_mask else None
        contour = S.contour_image(i) if with_contour else None
        det = oracle.cvorb_detect(img, mask)
        kept = det[oracle.bird_near_edges(contour, np.stack([det["x"], det["y"]], 1)).astype(bool)] if with_contour else det
        xy, _ = oracle.corner_subpix(img, np.stack([kept["x"], kept["y"]], 1))
        moved = kept.copy()
        moved["x"], moved["y"] = xy[:, 0], xy[:, 1]
        rk, rd = oracle.cvorb_compute(img, moved)
        cap = 8192
        out = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        p = lambda a: None if a is None else np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)
        n = dropin.L.refm_bird_features(p(img), p(mask), p(contour), 384, 384, p(out), p(desc), cap)
        assert n == len(rk) > 100
        assert out[:n].tobytes() == rk.tobytes() and np.array_equal(desc[:n], rd), i
