"""Frame::isInFrustum on the GPU (fbe_is_in_frustum through the C-ABI) against the oracle: acceptance, projections and
viewing cosine bit-equal; predicted level equal except on the flagged logf rounding boundaries."""
import numpy as np
import pytest

from frustum_scenes import scene

pytestmark = pytest.mark.gpu
GOLD_OW = None


def views(oracle, sc):
    from fishbirdeyevisualslam_b200 import _lib
    vo = oracle.frustum_view(sc["Rcw"], sc["tcw"], sc["Ow"], sc["K"], sc["bounds"], sc["mbf"], sc["log_scale"], sc["n_levels"])
    vg = _lib.FrustumView.from_buffer_copy(bytes(vo))
    return vo, vg


@pytest.mark.parametrize("seed", range(4))
def test_gpu_equals_oracle(oracle, seed):
    from fishbirdeyevisualslam_b200.matcher import isInFrustum
    sc = scene(seed, n=20000 if seed == 3 else 3000)
    R, t = sc["Rcw"].astype(np.float64), sc["tcw"].astype(np.float64)
    sc["Ow"] = (-(R.T @ t)).astype(np.float32)                  # an input of the function; any float triple will do
    vo, vg = views(oracle, sc)
    o = oracle.is_in_frustum(vo, sc["pos"], sc["normal"], sc["min_dist"], sc["max_dist"], sc["cos_limit"])
    g = isInFrustum(vg, sc["pos"], sc["normal"], sc["min_dist"], sc["max_dist"], sc["cos_limit"])
    for k in ("in_view", "proj", "proj_xr", "view_cos"):
        assert np.array_equal(g[k], o[k], equal_nan=True), k
    diff = g["level"] != o["level"]
    assert not (diff & (o["level_boundary"] == 0)).any()
    assert diff.sum() <= max(2, o["level_boundary"].sum())
    assert o["in_view"].sum() > 500


def test_empty_and_null_outputs():
    from fishbirdeyevisualslam_b200 import _lib
    from fishbirdeyevisualslam_b200.matcher import isInFrustum
    v = _lib.FrustumView()
    v.n_levels = 8
    o = isInFrustum(v, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32), 0.5)
    assert len(o["in_view"]) == 0
