"""Frame::isInFrustum: the oracle restatement against values produced with REAL OpenCV 4.13 calls for every
cv::Mat expression of the reference (tests/golden/frustum.npz, tools/gen_golden_frustum.py) -- and, where cv2 is importable,
against the same generator run live.  CPU only."""
import os

import numpy as np
import pytest

from frustum_scenes import scene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frustum.npz")


def run_oracle(oracle, sc):
    v = oracle.frustum_view(sc["Rcw"], sc["tcw"], sc["Ow"], sc["K"], sc["bounds"], sc["mbf"], sc["log_scale"], sc["n_levels"])
    return oracle.is_in_frustum(v, sc["pos"], sc["normal"], sc["min_dist"], sc["max_dist"], sc["cos_limit"])


@pytest.mark.parametrize("seed", range(3))
def test_oracle_equals_opencv_generated_golden(oracle, seed):
    g = np.load(GOLD)
    sc = scene(seed)
    sc["Ow"] = g[f"s{seed}_Ow"]
    o = run_oracle(oracle, sc)
    for k in ("in_view", "proj", "proj_xr", "level", "view_cos"):
        assert np.array_equal(o[k], g[f"s{seed}_{k}"], equal_nan=True), k
    n_in = int(o["in_view"].sum())
    assert 600 < n_in < 2000                                     # every rejection branch is exercised
    assert len(np.unique(o["level"][o["in_view"] == 1])) == 8 and o["level_boundary"].sum() > 20


def test_oracle_equals_live_opencv(oracle):
    cv2 = pytest.importorskip("cv2")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import gen_golden_frustum as G
    sc = scene(11, n=1500, cv2_pose=True)
    want = G.cv_is_in_frustum(sc)
    o = run_oracle(oracle, sc)
    for k in ("in_view", "proj", "proj_xr", "level", "view_cos"):
        assert np.array_equal(o[k], want[k], equal_nan=True), k


def test_rejection_reasons(oracle):
    """One hand-built point per branch of Frame.cc:449-480 (identity pose, fx = fy = 100, principal point (50, 40))."""
    v = oracle.frustum_view(np.eye(3), np.zeros(3), np.zeros(3), (100, 100, 50, 40), (0, 100, 0, 80), 0, np.log(np.float32(1.2)), 8)
    pos = np.array([[0, 0, -1],        # behind the camera
                    [1, 0, 1],         # u = 150 > mnMaxX
                    [0, -1, 1],        # v = -60 < mnMinY
                    [0, 0, 1],         # dist 1 < 0.8 * 2
                    [0, 0, 10],        # dist 10 > 1.2 * 8
                    [0, 0, 2],         # normal sideways: viewCos 0 < 0.5
                    [0, 0, 2],         # accepted, level = ceil(log(8/2)/log(1.2)) = 8 -> clamped to 7
                    [0.1, 0.1, 2],     # accepted, level 0 (ratio < 1)
                    [0.5, 0.4, 1]], np.float32)   # accepted exactly on the bounds u = 100, v = 80 (inclusive)
    normal = np.array([[0, 0, 1]] * 5 + [[1, 0, 0]] + [[0, 0, 1]] * 3, np.float32)
    min_d = np.array([0, 0, 0, 2, 0, 0, 0, 0, 0], np.float32)
    max_d = np.array([100, 100, 100, 100, 8, 100, 8, 1.9, 100], np.float32)
    o = oracle.is_in_frustum(v, pos, normal, min_d, max_d, 0.5)
    assert o["in_view"].tolist() == [0, 0, 0, 0, 0, 0, 1, 1, 1]
    assert o["level"][6] == 7 and o["level"][7] == 0
    assert o["proj"][8].tolist() == [100.0, 80.0] and o["proj"][6].tolist() == [50.0, 40.0]
