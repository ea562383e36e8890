"""Row f-1: Frame::UndistortKeyPoints + ComputeImageBounds (src/Frame.cc:638-669, 741-795).
CPU: the oracle's restatement of cv::fisheye::undistortPoints equals cv2 4.13.0 bit for bit on the committed known answers
(tests/golden/undistort.npz, tools/gen_golden_undistort.py).  GPU: the C-ABI equals the oracle within 1e-3 px (the stated
bar for this floating-point row; in practice the float results are identical) and feeds the same grid assignment."""
import os

import numpy as np
import pytest

from fishbirdeyevisualslam_b200._lib import KP_DTYPE

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "undistort.npz")
TOL_PX = 1e-3


def test_oracle_equals_cv2_golden(oracle):
    g = np.load(GOLD)
    assert str(g["cv2_version"]) == "4.13.0"
    for i in range(int(g["ncases"])):
        o = oracle.fisheye_undistort(g[f"pts{i}"], g[f"K{i}"], g[f"D{i}"])
        assert o.tobytes() == g[f"out{i}"].tobytes()


def _kps(pts):
    k = np.zeros(len(pts), KP_DTYPE)
    k["x"], k["y"] = pts[:, 0], pts[:, 1]
    k["size"], k["angle"], k["response"], k["octave"], k["class_id"] = 31, 12.5, 20, 0, -1
    return k


@pytest.mark.gpu
def test_undistort_keypoints_and_bounds(oracle):
    from fishbirdeyevisualslam_b200.matcher import ComputeImageBounds, UndistortKeyPoints, grid_assign
    g = np.load(GOLD)
    for i in range(int(g["ncases"])):
        pts, K, D = g[f"pts{i}"], g[f"K{i}"], g[f"D{i}"]
        w, h = (int(v) for v in g[f"size{i}"])
        kin = _kps(pts)
        ku = UndistortKeyPoints(kin, K, D)
        o = oracle.fisheye_undistort(pts, K, D)
        got = np.stack([ku["x"], ku["y"]], 1)
        assert np.abs(got - o).max() <= TOL_PX
        assert (got.view(np.uint32) != o.view(np.uint32)).mean() < 0.01          # float results identical but for rare ulp cases
        for f in ("size", "angle", "response", "octave", "class_id"):           # kp = mvKeys[i]; only pt changes
            assert np.array_equal(ku[f], kin[f])
        # bounds from the four corners, reference min/max quirk included
        b = ComputeImageBounds(w, h, K, D)
        c = oracle.fisheye_undistort(np.float32([[0, 0], [w, 0], [0, h], [w, h]]), K, D)
        fmin = np.finfo(np.float32).tiny
        want = (c[:, 0].min(), max(c[:, 0].max(), fmin), c[:, 1].min(), max(c[:, 1].max(), fmin))
        assert np.allclose(b, want, atol=TOL_PX)
        # the undistorted keypoints bucket exactly like the oracle's on the same bounds (AssignFeaturesToGrid on mvKeysUn)
        ko = kin.copy(); ko["x"], ko["y"] = o[:, 0], o[:, 1]
        inv_w, inv_h = np.float32(64.0) / np.float32(b[1] - b[0]), np.float32(48.0) / np.float32(b[3] - b[2])
        same = got.view(np.uint32) == o.view(np.uint32)
        if same.all():
            s1, i1 = grid_assign(ku, b[0], b[2], inv_w, inv_h, 64, 48)
            s2, i2 = oracle.grid_assign(ko, b[0], b[2], inv_w, inv_h, 64, 48)
            assert np.array_equal(s1, s2) and np.array_equal(i1, i2)
    # k1 == 0: plain copy, rectangle bounds
    kin = _kps(g["pts0"][:50])
    assert UndistortKeyPoints(kin, g["K0"], np.zeros(4, np.float32)).tobytes() == kin.tobytes()
    assert ComputeImageBounds(950, 400, g["K0"], np.zeros(4, np.float32)) == (0.0, 950.0, 0.0, 400.0)
