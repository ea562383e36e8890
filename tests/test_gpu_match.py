"""Parity of the CUDA matchers (through the C-ABI) with the oracle: index lists and counts bit-exact."""
import ctypes as C

import numpy as np
import pytest

import pyref
from conftest import make_kps
from fishbirdeyevisualslam_b200 import synth
from scenes import featvec, flip_bits, frame_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def front_pair():
    from fishbirdeyevisualslam_b200.extractor import ORBextractor
    from fishbirdeyevisualslam_b200.matcher import Frame
    h, w = 480, 640
    a, b = synth.frame_pair_in_time(h, w, 11)
    ex = ORBextractor(1000, 1.2, 8, 15, 5)
    ka, da = ex(a)
    kb, db = ex(b)
    sf = ex.GetScaleFactors()
    return Frame.front(ka, da, w, h, sf), Frame.front(kb, db, w, h, sf)


@pytest.fixture(scope="module")
def bird_pair():
    from fishbirdeyevisualslam_b200.extractor import ORBextractor
    from fishbirdeyevisualslam_b200.matcher import Frame
    a, b = synth.frame_pair_in_time(384, 384, 21)
    ex = ORBextractor(1000, 1.2, 8, 15, 5)
    ka, da = ex(a)
    kb, db = ex(b)
    return Frame.bird(ka, da, 384, 384), Frame.bird(kb, db, 384, 384)


def test_grid_assign(oracle, front_pair, bird_pair):
    for F in (*front_pair, *bird_pair):
        s_g, i_g = F.AssignFeaturesToGrid()
        s_o, i_o = oracle.grid_assign(F.kps, F.min_x, F.min_y, F.inv_w, F.inv_h, F.gcols, F.grows)
        assert np.array_equal(s_g, s_o) and np.array_equal(i_g, i_o)
    # quirk Q1 + empty + out-of-grid keypoints
    from fishbirdeyevisualslam_b200.matcher import grid_assign
    k = make_kps(np.float32([14.9, 15.0, 635.0, 634.9, -20.0, 3000.0]), np.float32([10, 10, 10, 10, 10, 10]))
    s_g, i_g = grid_assign(k, 0.0, 0.0, 0.1, 0.1, 64, 48)
    s_o, i_o = oracle.grid_assign(k, 0.0, 0.0, 0.1, 0.1, 64, 48)
    assert np.array_equal(s_g, s_o) and np.array_equal(i_g, i_o) and len(i_g) == 3
    s_g, i_g = grid_assign(k[:0], 0.0, 0.0, 0.1, 0.1, 64, 48)
    assert s_g.sum() == 0 and len(i_g) == 0


@pytest.mark.parametrize("ratio,ori,win", [(0.9, True, 100), (0.9, False, 100), (0.7, True, 30), (0.9, True, 400), (0.6, True, 10)])
def test_search_for_initialization(oracle, front_pair, ratio, ori, win):
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    F1, F2 = front_pair
    m = ORBmatcher(ratio, ori)
    pm_g = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
    pm_o = pm_g.copy()
    for _ in range(2):     # second round starts from the updated vbPrevMatched, as Tracking does
        n_g, m_g = m.SearchForInitialization(F1, F2, pm_g, win)
        n_o, m_o = oracle.search_for_initialization(F1, F2, pm_o, win, ratio, ori)
        assert n_g == n_o and np.array_equal(m_g, m_o) and np.array_equal(pm_g, pm_o)
    assert n_g > 20 or win <= 10


def test_initialization_row_capacity_growth(oracle):
    """A window that holds more candidates than the initial row capacity (128) forces the grow-and-redo path."""
    from fishbirdeyevisualslam_b200.matcher import Frame, ORBmatcher
    rng = np.random.default_rng(3)
    n = 700
    k2 = make_kps(rng.uniform(280, 360, n).astype(np.float32), rng.uniform(200, 280, n).astype(np.float32))
    d2 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    k1 = make_kps(np.float32([320, 322, 318]), np.float32([240, 241, 239]))
    d1 = flip_bits(rng, d2[[5, 77, 300]], 10)
    F1, F2 = Frame.front(k1, d1, 640, 480), Frame.front(k2, d2, 640, 480)
    pm_g = np.ascontiguousarray(np.stack([k1["x"], k1["y"]], 1), np.float32)
    pm_o = pm_g.copy()
    n_g, m_g = ORBmatcher(0.9, True).SearchForInitialization(F1, F2, pm_g, 100)
    n_o, m_o = oracle.search_for_initialization(F1, F2, pm_o, 100, 0.9, True)
    assert n_g == n_o == 3 and np.array_equal(m_g, m_o)


@pytest.mark.parametrize("ratio,ori,win", [(0.9, True, 10), (0.9, False, 10), (0.8, True, 25)])
def test_birdview_match(oracle, bird_pair, ratio, ori, win):
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    F1, F2 = bird_pair
    n_g, d_g = ORBmatcher(ratio, ori).BirdviewMatch(F2, F1.kps, F1.desc, win)
    n_o, d_o = oracle.birdview_match(F1.kps, F1.desc, F2, win, ratio, ori)
    assert n_g == n_o and np.array_equal(d_g, d_o) and n_g > 20


def test_bird_map_point_match_c3(oracle, bird_pair):
    """Config C3: 20 000 projected bird map points against the keypoints of one bird frame."""
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher, PIXEL2METER, REAR_AXLE_TO_CENTER
    _, FB = bird_pair
    rng = np.random.default_rng(5)
    nmp = 20000
    src = rng.integers(0, FB.N, nmp)
    mp_desc = FB.desc[src].copy()
    nflip = rng.integers(0, 41, nmp)
    bitpos = rng.integers(0, 256, (nmp, 40))
    for j in range(40):
        sel = nflip > j
        mp_desc[sel, bitpos[sel, j] >> 3] ^= (1 << (bitpos[sel, j] & 7)).astype(np.uint8)
    # world points: bird pixel -> base XY (Converter::BirdPixel2BaseXY) -> world through Tbw^-1, +-0.1 m jitter
    px, py = FB.kps["x"][src], FB.kps["y"][src]
    bx = (192 - py) * PIXEL2METER + REAR_AXLE_TO_CENTER
    by = (192 - px) * PIXEL2METER
    base = np.stack([bx, by, np.zeros(nmp)], 1) + rng.uniform(-0.1, 0.1, (nmp, 3)) * [1, 1, 0.5]
    yaw = np.deg2rad(2.0)
    Tbw = np.eye(4, dtype=np.float32)
    Tbw[:3, :3] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]
    Tbw[:3, 3] = [0.3, 0.1, 0.0]
    world = ((base - Tbw[:3, 3]) @ Tbw[:3, :3]).astype(np.float32)           # R^T (p - t)
    world[rng.random(nmp) < 0.03, 0] = np.nan                                # NULL MapPointBird*
    Tcw = np.eye(4, dtype=np.float32)
    cam_xyz = rng.normal(0, 1, (FB.N, 3)).astype(np.float32)
    m = ORBmatcher(0.9, True)
    inl, assigned, m12 = m.BirdMapPointMatch(FB, world, mp_desc, Tbw, Tcw, cam_xyz, 384, 384, 10, 0.05)
    # oracle on the same host-projected pixels
    from fishbirdeyevisualslam_b200.matcher import BaseXY2BirdPixel, transform_points
    local = transform_points(Tbw, np.nan_to_num(world))
    pix = BaseXY2BirdPixel(local, 384, 384)
    skip = np.isnan(world[:, 0]) | (np.abs(local[:, 2]) > 0.2) | (pix[:, 0] < 0) | (pix[:, 0] >= 384) | (pix[:, 1] < 0) | (pix[:, 1] >= 384)
    pix[skip, 0] = np.nan
    n_o, m_o = oracle.bird_map_point_match(pix, mp_desc, FB, 10, 0.9)
    assert np.array_equal(m12, m_o) and (m12 >= 0).sum() == n_o and n_o > 1000


def test_search_by_projection_last(oracle, front_pair):
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    F1, F2 = front_pair
    rng = np.random.default_rng(6)
    proj = np.stack([F1.kps["x"], F1.kps["y"]], 1).astype(np.float32) + rng.normal(0, 3, (F1.N, 2)).astype(np.float32)
    proj[rng.random(F1.N) < 0.2, 0] = np.nan
    taken = (rng.random(F2.N) < 0.1).astype(np.uint8)
    obs = (rng.random(F1.N) < 0.7).astype(np.uint8)
    for th, ori, ho in [(15, True, None), (30, True, None), (15, False, None), (15, True, obs)]:
        n_g, c_g = ORBmatcher(0.9, ori).SearchByProjectionLast(F2, F1.kps, proj, F1.desc, th, cur_taken=taken, last_has_obs=ho)
        n_o, c_o = oracle.search_by_projection_last(F2, F1.kps, proj, F1.desc, F2.scale_factors, th, ori, cur_taken=taken, last_has_obs=ho)
        assert n_g == n_o and np.array_equal(c_g, c_o) and n_g > 100


def test_search_by_projection_reloc_and_loop(oracle, front_pair):
    """a-18: the relocalisation (:1473-1600) and loop-closing (:291-404) overloads."""
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    F1, F2 = front_pair
    rng = np.random.default_rng(16)
    proj = np.stack([F1.kps["x"], F1.kps["y"]], 1).astype(np.float32) + rng.normal(0, 3, (F1.N, 2)).astype(np.float32)
    proj[rng.random(F1.N) < 0.2, 0] = np.nan
    level = np.clip(F1.kps["octave"] + rng.integers(-1, 2, F1.N), 0, 7).astype(np.int32)
    mdesc = flip_bits(rng, F1.desc, 40)
    taken = (rng.random(F2.N) < 0.1).astype(np.uint8)
    for th, orbdist, ori in [(10, 100, True), (3, 64, True), (10, 100, False)]:
        n_g, c_g = ORBmatcher(0.9, ori).SearchByProjectionReloc(F2, F1.kps, proj, level, mdesc, th, orbdist, cur_taken=taken)
        n_o, c_o = oracle.search_by_projection_kf(F2, F1.kps, proj, level, mdesc, F2.scale_factors, th, orbdist, 1, ori, cur_taken=taken)
        assert n_g == n_o and np.array_equal(c_g, c_o) and n_g > 100
    for th in (10, 4):
        n_g, c_g = ORBmatcher(0.75, True).SearchByProjectionLoop(F2, proj, level, mdesc, th, kf_matched=taken)
        n_o, c_o = oracle.search_by_projection_kf(F2, F1.kps, proj, level, mdesc, F2.scale_factors, th, 50, 0, False, cur_taken=taken)
        assert n_g == n_o and np.array_equal(c_g, c_o) and n_g > 50


def test_search_by_projection_map(oracle, front_pair):
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    _, F2 = front_pair
    rng = np.random.default_rng(7)
    nmap = 6000
    src = rng.integers(0, F2.N, nmap)
    mdesc = flip_bits(rng, F2.desc[src], 60)
    mproj = np.stack([F2.kps["x"][src], F2.kps["y"][src]], 1).astype(np.float32) + rng.normal(0, 2, (nmap, 2)).astype(np.float32)
    mlevel = np.clip(F2.kps["octave"][src] + rng.integers(-1, 2, nmap), 0, 7).astype(np.int32)
    mcos = rng.uniform(0.99, 1.0, nmap).astype(np.float32)
    taken = (rng.random(F2.N) < 0.1).astype(np.uint8)
    for th, ratio in [(1.0, 0.8), (3.0, 0.8), (5.0, 0.6)]:
        n_g, c_g = ORBmatcher(ratio, True).SearchByProjectionMap(F2, mproj, mlevel, mcos, mdesc, th, cur_taken=taken)
        n_o, c_o = oracle.search_by_projection_map(F2, F2.scale_factors, mproj, mlevel, mcos, mdesc, th, ratio, cur_taken=taken)
        assert n_g == n_o and np.array_equal(c_g, c_o) and n_g > 300


def test_search_by_bow(oracle, front_pair):
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    F1, F2 = front_pair

    def nodes(k, shift):
        return (np.floor((k["x"] - shift[0]) / 80).astype(int) * 16 + np.floor((k["y"] - shift[1]) / 80).astype(int)) * 8 + k["octave"]
    kfv, ffv = featvec(nodes(F1.kps, (0, 0))), featvec(nodes(F2.kps, (3, 2)))
    has_mp = (np.random.default_rng(8).random(F1.N) < 0.8).astype(np.uint8)
    for ratio, ori in [(0.7, True), (0.75, False)]:
        n_g, f_g = ORBmatcher(ratio, ori).SearchByBoW(F1.kps, F1.desc, has_mp, kfv, F2, ffv)
        n_o, f_o = oracle.search_by_bow(F1.kps, F1.desc, has_mp, kfv, F2.kps, F2.desc, ffv, ratio, ori)
        assert n_g == n_o and np.array_equal(f_g, f_o) and n_g > 50
    # key frame vs key frame (loop closing): strict threshold, matched key-frame-2 features blocked, result indexed by kf1
    has_mp2 = (np.random.default_rng(9).random(F2.N) < 0.85).astype(np.uint8)
    one1, one2 = featvec(np.zeros(F1.N, int)), featvec(np.zeros(F2.N, int))
    for fv1, fv2 in ((kfv, ffv), (one1, one2)):
        for ratio, ori in [(0.75, True), (0.9, False)]:
            n_g, m_g = ORBmatcher(ratio, ori).SearchByBoWKF(F1.kps, F1.desc, has_mp, fv1, F2.kps, F2.desc, has_mp2, fv2)
            n_o, m_o = oracle.search_by_bow_kf(F1.kps, F1.desc, has_mp, fv1, F2.kps, F2.desc, has_mp2, fv2, ratio, ori)
            assert n_g == n_o and np.array_equal(m_g, m_o) and n_g > 50


def test_search_for_triangulation(oracle):
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    from test_oracle_vs_refmatch import triangulation_args
    rng = np.random.default_rng(31)
    F1, F2 = frame_pair(rng, 1500)

    def nodes(k, shift):
        return (np.floor((k["x"] - shift[0]) / 80).astype(int) * 16 + np.floor((k["y"] - shift[1]) / 80).astype(int)) * 8 + k["octave"]
    has1, has2 = (rng.random(F1.N) < 0.5).astype(np.uint8), (rng.random(F2.N) < 0.5).astype(np.uint8)
    for node2 in (nodes(F2.kps, (3, 2)), np.zeros(F2.N, int)):
        node1 = nodes(F1.kps, (0, 0)) if node2.any() else np.zeros(F1.N, int)
        args = triangulation_args(F1, F2, has1, has2, node1, node2, 3)
        for only_stereo, ori in [(False, True), (False, False), (True, True)]:
            n_g, m_g = ORBmatcher(0.6, ori).SearchForTriangulation(*args, bOnlyStereo=only_stereo)
            n_o, m_o = oracle.search_for_triangulation(*args, only_stereo, 0.6, ori)
            assert n_g == n_o and np.array_equal(m_g, m_o)
            assert n_g > (5 if only_stereo else 40)


def test_fuse_search(oracle):
    """The candidate search of both Fuse overloads (with / without the reprojection gate, mono + stereo key-frame features)."""
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    from test_oracle_vs_refmatch import fuse_args
    rng = np.random.default_rng(41)
    _, KF = frame_pair(rng, 1500)
    kf, uright, inv, proj, level, desc = fuse_args(KF, 5, n=4000)[:6]
    sf = KF.scale_factors
    pur = (proj[:, 0] - np.float32(0.5)).astype(np.float32)
    for th, chi2, ur in [(3.0, True, uright), (3.0, True, None), (4.0, False, None), (0.5, True, uright)]:
        radius = (np.float32(th) * sf[level]).astype(np.float32)
        bi_g, bd_g = ORBmatcher(0.6, True).FuseSearch(kf, ur, inv, proj, pur, level, radius, desc, chi2)
        bi_o, bd_o = oracle.fuse_search(kf, ur, inv, proj, pur, level, radius, desc, chi2)
        assert np.array_equal(bi_g, bi_o) and np.array_equal(bd_g, bd_o)
        assert (bi_o >= 0).sum() > (100 if th < 1 else 1500)


def test_search_by_sim3(oracle):
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    from test_oracle_vs_refmatch import sim3_args
    rng = np.random.default_rng(51)
    F1, F2 = frame_pair(rng, 1500)
    _, _, pos1, lvl1, d1, pos2, lvl2, d2, pre = sim3_args(F1, F2, 9)
    pre[:] = -1
    t = np.float32([-3.0, -2.0])
    p12 = (pos1 - t).astype(np.float32); p21 = (pos2 + t).astype(np.float32)
    for p in (p12, p21):
        p[~((p[:, 0] >= 0) & (p[:, 0] < 640) & (p[:, 1] >= 0) & (p[:, 1] < 480)), 0] = np.nan
    for th in (7.5, 3.0):
        n_g, m_g = ORBmatcher(0.75, True).SearchBySim3(F1, F2, p12, lvl1, d1, p21, lvl2, d2, pre, th)
        n_o, m_o = oracle.search_by_sim3(F1, F2, pos1, lvl1, d1, pos2, lvl2, d2, pre, t, th)
        assert n_g == n_o and np.array_equal(m_g, m_o) and n_g > 300


def test_ransac_scoring(oracle):
    """Initializer::CheckHomography / CheckFundamental for 200 hypotheses x ~1000 matches: scores bit-identical (sequential
    float accumulation order reproduced), inlier masks identical; n > one shared-memory chunk is covered by the 5000-match case."""
    from fishbirdeyevisualslam_b200.matcher import CheckFundamental, CheckHomography
    from test_oracle_vs_refmatch import model_args
    rng = np.random.default_rng(61)
    for npts in (1500, 7000):
        F1, F2 = frame_pair(rng, npts)
        prev = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
        _, m12 = oracle.search_for_initialization(F1, F2, prev, 100, 0.9, False)
        (k1, k2, matches), models = model_args(F1, F2, m12, 3, K=200)
        assert len(matches) > (500 if npts == 1500 else 2100)
        for sigma in (1.0, 0.7):
            H21, H12, _ = models["H"]
            s_g, i_g = CheckHomography(k1, k2, matches, H21, H12, sigma)
            s_o, i_o = oracle.check_models(k1, k2, matches, H21, H12, sigma, True)
            assert np.array_equal(s_g.view(np.int32), s_o.view(np.int32)) and np.array_equal(i_g, i_o)
            F21 = models["F"][0]
            s_g, i_g = CheckFundamental(k1, k2, matches, F21, sigma)
            s_o, i_o = oracle.check_models(k1, k2, matches, F21, None, sigma, False)
            assert np.array_equal(s_g.view(np.int32), s_o.view(np.int32)) and np.array_equal(i_g, i_o)
            assert 0.05 < i_o.mean() < 0.95
    s, i = CheckHomography(k1, k2, np.zeros((0, 2), np.int32), H21[:3], H12[:3], 1.0)
    assert s.tolist() == [0.0, 0.0, 0.0]


def test_distinctive_descriptors(oracle):
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    from test_oracle_vs_refmatch import distinct_lists
    rng = np.random.default_rng(21)
    sizes = tuple(int(x) for x in rng.integers(0, 45, 400)) + (0, 1, 2, 300, 129, 64, 33, 32, 31)
    dd, st = distinct_lists(rng, sizes)
    b_g, m_g = ORBmatcher(0.9, True).ComputeDistinctiveDescriptors(dd, st)
    b_o, m_o = oracle.distinctive_descriptors(dd, st)
    assert np.array_equal(b_g, b_o) and np.array_equal(m_g, m_o)
    e_b, e_m = ORBmatcher(0.9, True).ComputeDistinctiveDescriptors(np.zeros((0, 32), np.uint8), np.zeros(3, np.int32))
    assert e_b.tolist() == [-1, -1] and e_m.tolist() == [0, 0]


@pytest.mark.parametrize("seed", range(3))
def test_random_scenes_all_matchers(oracle, seed):
    """Synthetic scenes with quantised coordinates (exact cell-boundary and window-boundary hits) and duplicates (ties)."""
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    rng = np.random.default_rng(100 + seed)
    F1, F2 = frame_pair(rng, 400)
    pm_g = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
    pm_o = pm_g.copy()
    n_g, m_g = ORBmatcher(0.9, True).SearchForInitialization(F1, F2, pm_g, 100)
    n_o, m_o = oracle.search_for_initialization(F1, F2, pm_o, 100, 0.9, True)
    assert n_g == n_o and np.array_equal(m_g, m_o) and np.array_equal(pm_g, pm_o)
    B1, B2 = frame_pair(rng, 400, 384, 384, bird=True)
    n_g, d_g = ORBmatcher(0.9, True).BirdviewMatch(B2, B1.kps, B1.desc, 10)
    n_o, d_o = oracle.birdview_match(B1.kps, B1.desc, B2, 10, 0.9, True)
    assert n_g == n_o and np.array_equal(d_g, d_o)


def test_window_wider_than_32_grid_columns(oracle):
    """The window walk flattens the grid columns of a window 32 at a time (csrc/match.cu: window_walk); windows of 41-64 columns
    need a second pass whose ranks continue the first pass's -- front (64 columns of 10 px) and bird (48 columns of 8 px) grids."""
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    rng = np.random.default_rng(77)
    F1, F2 = frame_pair(rng, 500)
    for win in (200, 330, 700):
        pm_g = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
        pm_o = pm_g.copy()
        n_g, m_g = ORBmatcher(0.9, True).SearchForInitialization(F1, F2, pm_g, win)
        n_o, m_o = oracle.search_for_initialization(F1, F2, pm_o, win, 0.9, True)
        assert n_g == n_o and np.array_equal(m_g, m_o) and np.array_equal(pm_g, pm_o), win
    B1, B2 = frame_pair(rng, 500, 384, 384, bird=True)
    for win in (150, 400):
        n_g, d_g = ORBmatcher(0.9, True).BirdviewMatch(B2, B1.kps, B1.desc, win)
        n_o, d_o = oracle.birdview_match(B1.kps, B1.desc, B2, win, 0.9, True)
        assert n_g == n_o and np.array_equal(d_g, d_o), win


@pytest.mark.parametrize("kind", ["ties", "steals"])
@pytest.mark.parametrize("seed", range(4))
def test_resolve_under_contention(oracle, seed, kind):
    """The sequential accept / steal chain is replayed in speculative waves of 16 queries (csrc/match.cu: k_resolve); this is the
    input that makes the speculation fail as often as possible: many octave-0 queries packed into one window, few targets,
    descriptors drawn from a handful of prototypes a few bits apart ("ties": queries of one wave keep choosing the same best / second
    best target and distances tie) or every target its own prototype with ~17 noisy copies among the queries ("steals": nearly every
    query is accepted, targets are stolen again and again, and a wave of 16 queries over 40-115 targets collides all the time).
    Every search mode that goes through k_resolve."""
    from fishbirdeyevisualslam_b200.matcher import Frame, ORBmatcher
    rng = np.random.default_rng(500 + seed)
    nproto, n1, n2 = 6, 700, 40 + 25 * seed
    if kind == "ties":
        proto = rng.integers(0, 256, (nproto, 32), dtype=np.uint8)
        d1 = flip_bits(rng, proto[rng.integers(0, nproto, n1)], 3)
        d2 = flip_bits(rng, proto[rng.integers(0, nproto, n2)], 2)
    else:
        d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        src = d2[rng.integers(0, n2, n1)]
        d1 = np.concatenate([flip_bits(rng, src[:200], 20), flip_bits(rng, src[200:450], 8), flip_bits(rng, src[450:], 2)])
        d1 = d1[rng.permutation(n1)]
    x1 = rng.uniform(300, 340, n1).astype(np.float32); y1 = rng.uniform(220, 260, n1).astype(np.float32)
    x2 = rng.uniform(300, 340, n2).astype(np.float32); y2 = rng.uniform(220, 260, n2).astype(np.float32)
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    a1 = rng.choice([10.0, 40.0, 200.0], n1).astype(np.float32); a2 = rng.choice([10.0, 40.0, 200.0], n2).astype(np.float32)
    F1 = Frame.front(make_kps(x1, y1, np.zeros(n1, np.int32), a1), d1, 640, 480, sf)
    F2 = Frame.front(make_kps(x2, y2, np.zeros(n2, np.int32), a2), d2, 640, 480, sf)
    for ratio, ori in [(0.9, True), (0.9, False), (1.0, False), (0.6, True)]:
        pm_g = np.ascontiguousarray(np.stack([x1, y1], 1), np.float32)
        pm_o = pm_g.copy()
        n_g, m_g = ORBmatcher(ratio, ori).SearchForInitialization(F1, F2, pm_g, 100)
        n_o, m_o = oracle.search_for_initialization(F1, F2, pm_o, 100, ratio, ori)
        assert n_g == n_o and np.array_equal(m_g, m_o) and np.array_equal(pm_g, pm_o), (ratio, ori)
    assert n_o > (0 if kind == "ties" else n2 // 2)
    # last-frame projection search: `taken` grows with every accept (has_obs) or never (no obs: later queries overwrite earlier ones)
    proj = np.stack([x1, y1], 1).astype(np.float32)
    obs = (rng.random(n1) < 0.5).astype(np.uint8)
    for th, ori, ho in [(30, True, None), (30, False, obs), (30, True, np.zeros(n1, np.uint8))]:
        n_g, c_g = ORBmatcher(0.9, ori).SearchByProjectionLast(F2, F1.kps, proj, d1, th, last_has_obs=ho)
        n_o, c_o = oracle.search_by_projection_last(F2, F1.kps, proj, d1, sf, th, ori, last_has_obs=ho)
        assert n_g == n_o and np.array_equal(c_g, c_o), (th, ori)
    # map-point projection search: the second-best candidate's LEVEL enters the ratio gate
    lv2 = rng.integers(0, 3, n2).astype(np.int32)
    F2l = Frame.front(make_kps(x2, y2, lv2, a2), d2, 640, 480, sf)
    mlevel = rng.integers(0, 3, n1).astype(np.int32)
    mcos = rng.uniform(0.99, 1.0, n1).astype(np.float32)
    for th, ratio in [(5.0, 0.8), (5.0, 0.95), (3.0, 0.6)]:
        n_g, c_g = ORBmatcher(ratio, True).SearchByProjectionMap(F2l, proj, mlevel, mcos, d1, th)
        n_o, c_o = oracle.search_by_projection_map(F2l, sf, proj, mlevel, mcos, d1, th, ratio)
        assert n_g == n_o and np.array_equal(c_g, c_o), (th, ratio)


def test_quirk_cases_on_device(oracle):
    from fishbirdeyevisualslam_b200.matcher import Frame, ORBmatcher
    rng = np.random.default_rng(0)
    # Q8: match to keypoint 0 is counted but not emitted
    d = rng.integers(0, 256, (3, 32), dtype=np.uint8)
    cur = Frame.bird(make_kps(np.float32([100, 200, 300]), np.float32([100, 200, 300])), d, 384, 384)
    ref = make_kps(np.float32([101, 201]), np.float32([101, 201]))
    n, dm = ORBmatcher(0.9, False).BirdviewMatch(cur, ref, d[:2], 10)
    assert n == 2 and dm.tolist() == [[1, 1, 0]]
    # Q2: exclusive upper cell bound of the bird query hides an in-window keypoint
    k = make_kps(np.float32([3.0, 105.0]), np.float32([3.0, 100.0]))
    dd = rng.integers(0, 256, (2, 32), dtype=np.uint8)
    cur = Frame.bird(k, dd, 384, 384)
    q = make_kps(np.float32([98.0]), np.float32([100.0]))
    n, dm = ORBmatcher(0.9, False).BirdviewMatch(cur, q, dd[1:2], 10)
    assert n == 0 and len(dm) == 0
    # Q10: gate (<=) and steal
    base = rng.integers(0, 256, 32, dtype=np.uint8)
    far = base.copy(); far[0] ^= 0xFF
    near = base.copy(); near[1] ^= 0x01
    F2 = Frame.front(make_kps(np.float32([100]), np.float32([100])), base[None], 640, 480)
    F1 = Frame.front(make_kps(np.float32([100, 101, 102]), np.float32([100, 100, 100])), np.stack([far, near, far]), 640, 480)
    pm = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
    n, m = ORBmatcher(0.9, False).SearchForInitialization(F1, F2, pm, 50)
    assert n == 1 and m.tolist() == [-1, 0, -1]
    # empty frames
    E = Frame.front(make_kps(np.float32([]), np.float32([])), np.zeros((0, 32), np.uint8), 640, 480)
    n, m = ORBmatcher(0.9, True).SearchForInitialization(F1, E, pm, 50)
    assert n == 0 and m.tolist() == [-1, -1, -1]
    n, m = ORBmatcher(0.9, True).SearchForInitialization(E, F2, np.zeros((0, 2), np.float32), 50)
    assert n == 0 and len(m) == 0


def test_bruteforce_top2_c5_size(oracle):
    """C5 matching leg: 8000 x 8000 brute-force Hamming, ties -> lowest index, second counts duplicates."""
    from fishbirdeyevisualslam_b200.matcher import ORBmatcher
    rng = np.random.default_rng(9)
    q = rng.integers(0, 256, (8000, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (8000, 32), dtype=np.uint8)
    t[100] = q[5]; t[4000] = q[5]; t[7999] = q[6]
    g = ORBmatcher(0.9, True).BruteForceTop2(q, t)
    # oracle on a slice (the scalar loop needs ~1 s per 1000 queries) + vectorised numpy check of the distances
    o = oracle.bruteforce_top2(q[:600], t)
    assert all(np.array_equal(a[:600], b) for a, b in zip(g, o))
    assert g[0][5] == 100 and g[1][5] == 0 and g[2][5] == 0 and g[0][6] == 7999
    idx = rng.integers(0, 8000, 300)
    dist = np.unpackbits(q[idx] ^ t[g[0][idx]], axis=1).sum(1)
    assert np.array_equal(dist, g[1][idx])
    assert (g[2] >= g[1]).all()


def test_widened_entry_points_edge_cases(oracle):
    """Empty and degenerate inputs of the rows added after the core path: they return cleanly with empty / -1 outputs."""
    import ctypes as C
    from fishbirdeyevisualslam_b200 import _lib
    from fishbirdeyevisualslam_b200.matcher import CheckFundamental, ORBmatcher
    rng = np.random.default_rng(71)
    F1, F2 = frame_pair(rng, 200)
    m = ORBmatcher(0.6, True)
    empty_fv = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    one1, one2 = featvec(np.zeros(F1.N, int)), featvec(np.zeros(F2.N, int))
    z1, z2 = np.zeros(F1.N, np.uint8), np.zeros(F2.N, np.uint8)
    sf = F2.scale_factors
    Fm = np.float32([[0, 0, -2], [0, 0, 3], [2, -3, 0]])
    # no shared vocabulary node / no eligible feature on either side
    n, m12 = m.SearchForTriangulation(F1.kps, F1.desc, z1, z1, empty_fv, F2.kps, F2.desc, z2, z2, one2, Fm, (0, 0), sf, sf * sf)
    assert n == 0 and (m12 == -1).all()
    n, m12 = m.SearchForTriangulation(F1.kps, F1.desc, z1 + 1, z1, one1, F2.kps, F2.desc, z2, z2, one2, Fm, (0, 0), sf, sf * sf)
    assert n == 0 and (m12 == -1).all()
    n, m12 = m.SearchByBoWKF(F1.kps, F1.desc, z1 + 1, one1, F2.kps, F2.desc, z2, one2)          # key frame 2 has no map points
    assert n == 0 and (m12 == -1).all()
    # degenerate fundamental matrix (den == 0 for every feature): CheckDistEpipolarLine returns false (:153-154)
    n, m12 = m.SearchForTriangulation(F1.kps, F1.desc, z1, z1, one1, F2.kps, F2.desc, z2, z2, one2, np.zeros((3, 3), np.float32), (-1e4, -1e4), sf, sf * sf)
    n_o, m_o = oracle.search_for_triangulation(F1.kps, F1.desc, z1, z1, one1, F2.kps, F2.desc, z2, z2, one2, np.zeros((3, 3), np.float32), (-1e4, -1e4), sf, sf * sf, False, 0.6, True)
    assert n == n_o == 0 and np.array_equal(m12, m_o)
    # Fuse search: every projection skipped / radius so small that no feature falls inside
    nanp = np.full((5, 2), np.nan, np.float32)
    bi, bd = m.FuseSearch(F2, None, 1 / (sf * sf), nanp, None, np.zeros(5, np.int32), np.ones(5, np.float32), np.zeros((5, 32), np.uint8), True)
    assert (bi == -1).all() and (bd == np.iinfo(np.int32).max).all()
    proj = np.stack([F2.kps["x"][:50] + 0.4, F2.kps["y"][:50] + 0.4], 1).astype(np.float32)
    bi, bd = m.FuseSearch(F2, None, 1 / (sf * sf), proj, None, F2.kps["octave"][:50].astype(np.int32), np.full(50, 1e-3, np.float32), F2.desc[:50], False)
    bi_o, bd_o = oracle.fuse_search(F2, None, 1 / (sf * sf), proj, None, F2.kps["octave"][:50].astype(np.int32), np.full(50, 1e-3, np.float32), F2.desc[:50], False)
    assert np.array_equal(bi, bi_o) and np.array_equal(bd, bd_o)
    # scoring with zero hypotheses; invalid match index is rejected, not read
    s, i = CheckFundamental(F1.kps, F2.kps, np.zeros((3, 2), np.int32), np.zeros((0, 3, 3), np.float32))
    assert len(s) == 0
    L = _lib.load()
    bad = np.array([[0, -1]], np.int32)
    sc = np.zeros(1, np.float32)
    rc = L.fbe_check_fundamental(F1.kps.ctypes.data_as(C.c_void_p), F2.kps.ctypes.data_as(C.c_void_p), bad.ctypes.data_as(C.c_void_p), 1,
                                 Fm.ctypes.data_as(C.c_void_p), 1, C.c_float(1.0), 0, sc.ctypes.data_as(C.c_void_p), None)
    assert rc != 0



def test_frame_cache_is_content_addressed(oracle):
    """Frames handed to the matcher stay on the device between calls and are recognised by CONTENT: repeated searches on the
    same frames hit the cache and give identical results; changing one keypoint in place (same buffer address) is a miss and
    gives the result of the changed frame."""
    import numpy as np
    from fishbirdeyevisualslam_b200 import synth
    from fishbirdeyevisualslam_b200.extractor import ORBextractor
    from fishbirdeyevisualslam_b200.matcher import Frame, ORBmatcher
    h, w = 240, 320
    a, b = synth.frame_pair_in_time(h, w, 21)
    ex = ORBextractor(500, 1.2, 6, 15, 5)
    ka, da = ex(a)
    kb, db = ex(b)
    F1, F2 = Frame.front(ka, da, w, h), Frame.front(kb, db, w, h)
    m = ORBmatcher(0.9, True)

    def run():
        pm = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
        n, m12 = m.SearchForInitialization(F1, F2, pm, 100)
        pmo = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
        no, m12o = oracle.search_for_initialization(F1, F2, pmo, 100, 0.9, True)
        assert n == no and np.array_equal(m12, m12o) and np.array_equal(pm, pmo)
        return n

    n0 = run()
    assert m.cache_stats() == (0, 2)
    assert run() == n0 and run() == n0
    assert m.cache_stats() == (4, 2)
    # SearchByProjection-style reuse with another window radius: still the same two frames
    pm = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
    m.SearchForInitialization(F1, F2, pm, 50)
    assert m.cache_stats() == (6, 2)
    # in-place change of the target frame (same address, different content) must not be served from the cache
    F2.kps["x"][5] += 40.0
    F2.kps["y"][5] += 25.0
    run()
    assert m.cache_stats() == (7, 3)
    # six different frames cycle through the four entries without confusing contents
    frames = []
    for s in range(6):
        k, d = ex(synth.frame(h, w, 40 + s))
        frames.append(Frame.front(k, d, w, h))
    for rep in range(2):
        for f in frames:
            pm = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
            pmo = pm.copy()
            n, m12 = m.SearchForInitialization(F1, f, pm, 100)
            no, m12o = oracle.search_for_initialization(F1, f, pmo, 100, 0.9, True)
            assert n == no and np.array_equal(m12, m12o)
    m.close()
