"""oracle/match_oracle.cpp (C++) against tests/pyref.py (pure Python, written separately from the reference source) on
random small scenes, plus hand-built cases for each quirk of SURVEY Appendix B.  CPU only."""
import numpy as np
import pytest

import pyref
from conftest import make_kps
from fishbirdeyevisualslam_b200.matcher import Frame
from scenes import featvec, flip_bits, frame_pair


@pytest.mark.parametrize("seed", range(4))
def test_grid_and_area_queries(oracle, seed):
    rng = np.random.default_rng(seed)
    for bird in (False, True):
        F1, F2 = frame_pair(rng, 250, 384 if bird else 640, 384 if bird else 480, bird=bird)
        s, it = oracle.grid_assign(F2.kps, F2.min_x, F2.min_y, F2.inv_w, F2.inv_h, F2.gcols, F2.grows)
        g = pyref.build_grid(F2)
        flat = [j for ix in range(F2.gcols) for iy in range(F2.grows) for j in g[ix][iy]]
        assert it.tolist() == flat
        assert s.tolist() == np.concatenate([[0], np.cumsum([len(g[ix][iy]) for ix in range(F2.gcols) for iy in range(F2.grows)])]).tolist()
        for _ in range(60):
            x, y = rng.uniform(-30, 700), rng.uniform(-30, 520)
            r = float(rng.choice([3, 10, 15.5, 40, 100]))
            lv = [(-1, -1), (0, 0), (1, 2), (0, -1), (2, -1)][int(rng.integers(0, 5))]
            got = oracle.features_in_area(F2, x, y, r, lv[0], lv[1], not bird)
            assert got.tolist() == pyref.area(F2, g, x, y, r, lv[0], lv[1], not bird)


@pytest.mark.parametrize("seed", range(4))
def test_search_for_initialization(oracle, seed):
    rng = np.random.default_rng(10 + seed)
    F1, F2 = frame_pair(rng, 220)
    for ratio, ori, win in [(0.9, True, 100), (0.9, False, 30), (0.6, True, 60)]:
        prev = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
        n_p, m_p, prev_p = pyref.search_for_initialization(F1, F2, prev, win, ratio, ori)
        n_o, m_o = oracle.search_for_initialization(F1, F2, prev, win, ratio, ori)
        assert (n_o, m_o.tolist()) == (n_p, m_p.tolist()) and np.array_equal(prev, prev_p)


@pytest.mark.parametrize("seed", range(4))
def test_bird_matchers(oracle, seed):
    rng = np.random.default_rng(20 + seed)
    F1, F2 = frame_pair(rng, 260, 384, 384, bird=True, flips=25)
    for ratio, ori, win in [(0.9, True, 10), (0.9, False, 10), (0.8, True, 22)]:
        n_p, d_p = pyref.birdview_match(F1.kps, F1.desc, F2, win, ratio, ori)
        n_o, d_o = oracle.birdview_match(F1.kps, F1.desc, F2, win, ratio, ori)
        assert n_o == n_p and np.array_equal(d_o, d_p)
    pix = np.stack([F1.kps["x"], F1.kps["y"]], 1).astype(np.float32) + rng.normal(0, 2, (len(F1.kps), 2)).astype(np.float32) + np.float32([3, 2])
    pix[rng.random(len(pix)) < 0.1, 0] = np.nan
    n_p, m_p = pyref.bird_map_point_match(pix, F1.desc, F2, 10, 0.9)
    n_o, m_o = oracle.bird_map_point_match(pix, F1.desc, F2, 10, 0.9)
    assert n_o == n_p and np.array_equal(m_o, m_p)


@pytest.mark.parametrize("seed", range(3))
def test_projection_searches(oracle, seed):
    rng = np.random.default_rng(30 + seed)
    F1, F2 = frame_pair(rng, 240)
    sf = F2.scale_factors
    proj = np.stack([F1.kps["x"], F1.kps["y"]], 1).astype(np.float32) + np.float32([3, 2])
    proj[rng.random(len(proj)) < 0.15, 0] = np.nan
    taken = (rng.random(F2.N) < 0.15).astype(np.uint8)
    obs = (rng.random(F1.N) < 0.7).astype(np.uint8)
    for th, ori, ho in [(15, True, None), (30, False, None), (15, True, obs)]:
        n_p, c_p = pyref.search_by_projection_last(F2, F1.kps, proj, F1.desc, sf, th, ori, taken, ho)
        n_o, c_o = oracle.search_by_projection_last(F2, F1.kps, proj, F1.desc, sf, th, ori, taken, ho)
        assert n_o == n_p and np.array_equal(c_o, c_p)
    level = np.clip(F1.kps["octave"] + rng.integers(-1, 2, F1.N), 0, 7).astype(np.int32)
    for th, thd, up, ori in [(10, 100, 1, True), (3, 64, 1, False), (10, 50, 0, False)]:      # reloc x2, loop closing
        n_p, c_p = pyref.search_by_projection_kf(F2, F1.kps, proj, level, F1.desc, sf, th, thd, up, ori, taken)
        n_o, c_o = oracle.search_by_projection_kf(F2, F1.kps, proj, level, F1.desc, sf, th, thd, up, ori, taken)
        assert n_o == n_p and np.array_equal(c_o, c_p)
    nmp = 400
    src = rng.integers(0, F2.N, nmp)
    mproj = np.stack([F2.kps["x"][src], F2.kps["y"][src]], 1).astype(np.float32) + rng.normal(0, 2, (nmp, 2)).astype(np.float32)
    mlev = np.clip(F2.kps["octave"][src] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
    mcos = rng.uniform(0.99, 1.0, nmp).astype(np.float32)
    mdesc = flip_bits(rng, F2.desc[src], 50)
    for th, ratio in [(1.0, 0.8), (3.0, 0.8), (5.0, 0.5)]:
        n_p, c_p = pyref.search_by_projection_map(F2, sf, mproj, mlev, mcos, mdesc, th, ratio, taken, None)
        n_o, c_o = oracle.search_by_projection_map(F2, sf, mproj, mlev, mcos, mdesc, th, ratio, taken, None)
        assert n_o == n_p and np.array_equal(c_o, c_p)


@pytest.mark.parametrize("seed", range(3))
def test_search_by_bow(oracle, seed):
    rng = np.random.default_rng(40 + seed)
    F1, F2 = frame_pair(rng, 260)
    node1 = rng.integers(0, 12, F1.N) * 3 + 1
    node2 = rng.integers(0, 14, F2.N) * 3 + 1          # some nodes exist on one side only
    has_mp = (rng.random(F1.N) < 0.8).astype(np.uint8)
    for ratio, ori in [(0.7, True), (0.9, False)]:
        n_p, f_p = pyref.search_by_bow(F1.kps, F1.desc, has_mp, featvec(node1), F2.kps, F2.desc, featvec(node2), ratio, ori)
        n_o, f_o = oracle.search_by_bow(F1.kps, F1.desc, has_mp, featvec(node1), F2.kps, F2.desc, featvec(node2), ratio, ori)
        assert n_o == n_p and np.array_equal(f_o, f_p)
    has_mp2 = (rng.random(F2.N) < 0.7).astype(np.uint8)
    for n1, n2 in ((node1, node2), (np.zeros(F1.N, int), np.zeros(F2.N, int))):     # second: one node, full contention
        for ratio, ori in [(0.75, True), (0.95, False)]:
            n_p, m_p = pyref.search_by_bow_kf(F1.kps, F1.desc, has_mp, featvec(n1), F2.kps, F2.desc, has_mp2, featvec(n2), ratio, ori)
            n_o, m_o = oracle.search_by_bow_kf(F1.kps, F1.desc, has_mp, featvec(n1), F2.kps, F2.desc, has_mp2, featvec(n2), ratio, ori)
            assert n_o == n_p and np.array_equal(m_o, m_p)
            hit = m_o[m_o >= 0]
            assert len(np.unique(hit)) == len(hit) and has_mp2[hit].all() and has_mp[m_o >= 0].all()


def test_distinctive_descriptors_against_numpy(oracle):
    """Row medians by numpy sort (MapPoint.cc:283-299): sorted_row[int(0.5*(N-1))], first least median wins."""
    from test_oracle_vs_refmatch import distinct_lists
    rng = np.random.default_rng(77)
    dd, st = distinct_lists(rng, sizes=(0, 1, 2, 3, 6, 6, 9, 30, 64, 101))
    best, med = oracle.distinctive_descriptors(dd, st)
    for p in range(len(st) - 1):
        d = dd[st[p]:st[p + 1]]
        if len(d) == 0:
            assert best[p] == -1
            continue
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2)
        m = np.sort(D, axis=1)[:, int(0.5 * (len(d) - 1))]
        assert best[p] == int(np.argmin(m)) and med[p] == int(m.min())


# ---- hand-built quirk cases ------------------------------------------------------------------
def test_q1_grid_uses_round_and_drops_last_half_cell(oracle):
    # 640x480 front grid: cell width 10 px.  x = 14.9 -> round(1.49) = 1; x = 15.0 -> round(1.5) = 2 (half away from zero);
    # x = 635 -> round(63.5) = 64 -> dropped
    k = make_kps(np.float32([14.9, 15.0, 635.0, 634.9]), np.float32([10, 10, 10, 10]))
    F = Frame.front(k, np.zeros((4, 32), np.uint8), 640, 480)
    s, it = oracle.grid_assign(k, F.min_x, F.min_y, F.inv_w, F.inv_h, 64, 48)
    cell = {int(i): c for c in range(64 * 48) for i in it[s[c]:s[c + 1]]}
    assert cell[0] // 48 == 1 and cell[1] // 48 == 2 and cell[3] // 48 == 63 and 2 not in cell


def test_q2_bird_area_upper_bound_is_exclusive(oracle):
    # bird 384x384: 12-px cells.  Query x=98, r=10: nMinCellX = floor(88/12) = 7, nMaxCellX = ceil(108/12) = 9.
    # The bird loop (`ix < nMaxCellX`) visits columns 7,8; the front-style loop (`<=`) also visits 9.
    # Keypoint 1 at x=105 lies in column round(105/12)=9 and inside the window (|dx| = 7 < 10): bird query misses it.
    k = make_kps(np.float32([100.0, 105.0]), np.float32([100.0, 100.0]))
    F = Frame.bird(k, np.zeros((2, 32), np.uint8), 384, 384)
    assert oracle.features_in_area(F, 98.0, 100.0, 10.0, -1, -1, True).tolist() == [0, 1]
    assert oracle.features_in_area(F, 98.0, 100.0, 10.0, -1, -1, False).tolist() == [0]
    # nMinCell == nMaxCell (window inside one clamped border cell): the bird query searches nothing at all
    k2 = make_kps(np.float32([380.0]), np.float32([380.0]))       # cell round(31.67) = 32 -> not even in the grid
    k3 = make_kps(np.float32([377.0]), np.float32([377.0]))       # cell round(31.4) = 31
    for kk, inc in ((k2, []), (k3, [0])):
        Fb = Frame.bird(kk, np.zeros((1, 32), np.uint8), 384, 384)
        assert oracle.features_in_area(Fb, 378.0, 378.0, 5.0, -1, -1, True).tolist() == inc
        assert oracle.features_in_area(Fb, 378.0, 378.0, 5.0, -1, -1, False).tolist() == []


def test_q3_q4_level_check_and_strict_window(oracle):
    k = make_kps(np.float32([50, 60, 50]), np.float32([50, 50, 50]), octave=np.int32([0, 2, 3]))
    F = Frame.front(k, np.zeros((3, 32), np.uint8), 640, 480)
    # minLevel == 0 and maxLevel < 0: bCheckLevels is false; |dx| == r is NOT inside the (strict) window
    assert oracle.features_in_area(F, 50, 50, 10.0, 0, -1, True).tolist() == [0, 2]
    assert oracle.features_in_area(F, 50, 50, 10.001, 0, -1, True).tolist() == [0, 2, 1]     # ix outer: cell 5 before cell 6
    assert oracle.features_in_area(F, 50, 50, 20.0, 1, -1, True).tolist() == [2, 1]          # minLevel > 0 enables the check
    assert oracle.features_in_area(F, 50, 50, 20.0, 0, 0, True).tolist() == [0]              # maxLevel >= 0 enables it too


def test_q8_match_to_index_zero_is_dropped(oracle):
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (3, 32), dtype=np.uint8)
    cur = Frame.bird(make_kps(np.float32([100, 200, 300]), np.float32([100, 200, 300])), d, 384, 384)
    ref = make_kps(np.float32([101, 201]), np.float32([101, 201]))
    n, dm = oracle.birdview_match(ref, d[:2], cur, 10, 0.9, False)
    assert n == 2                       # both counted ...
    assert dm.tolist() == [[1, 1, 0]]   # ... but the match to keypoint 0 is not emitted


def test_q7_ratio_failures_still_vote_in_the_histogram(oracle):
    # 12 queries: 4 pass the ratio test at rotation bin 0, 8 fail it at bin 3 (two identical targets) but still vote,
    # making bin 3 the maximum; with 10% rule bin 0 survives as second maximum -> all 4 matches kept.
    rng = np.random.default_rng(1)
    xs = np.arange(12, dtype=np.float32) * 30 + 20
    qd = rng.integers(0, 256, (12, 32), dtype=np.uint8)
    ref = make_kps(xs, np.full(12, 50, np.float32), angle=np.float32([0] * 4 + [95] * 8))
    tx, td, ta = [0.0], [rng.integers(0, 256, 32, dtype=np.uint8)], [0.0]    # index 0 is a dummy
    for i in range(12):
        tx += [xs[i] + 1] * (1 if i < 4 else 2)
        td += [qd[i]] * (1 if i < 4 else 2)
        ta += [0.0] * (1 if i < 4 else 2)
    cur = Frame.bird(make_kps(np.float32(tx), np.full(len(tx), 51, np.float32), angle=np.float32(ta)), np.array(td, np.uint8), 384, 384)
    n, dm = oracle.birdview_match(ref, qd, cur, 10, 0.9, True)
    n2, dm2 = pyref.birdview_match(ref, qd, cur, 10, 0.9, True)
    assert n == n2 == 4 and np.array_equal(dm, dm2) and len(dm) == 4


def test_q10_gate_and_steal_in_initialization(oracle):
    rng = np.random.default_rng(2)
    base = rng.integers(0, 256, 32, dtype=np.uint8)
    d_far = base.copy(); d_far[0] ^= 0xFF                 # distance 8 from base
    d_near = base.copy(); d_near[1] ^= 0x01               # distance 1
    F2 = Frame.front(make_kps(np.float32([100]), np.float32([100])), base[None], 640, 480)
    # query 0 matches with distance 8, query 1 steals with distance 1, query 2 (distance 8 again) is gated out
    F1 = Frame.front(make_kps(np.float32([100, 101, 102]), np.float32([100, 100, 100])), np.stack([d_far, d_near, d_far]), 640, 480)
    prev = np.ascontiguousarray(np.stack([F1.kps["x"], F1.kps["y"]], 1), np.float32)
    n, m = oracle.search_for_initialization(F1, F2, prev, 50, 0.9, False)
    assert n == 1 and m.tolist() == [-1, 0, -1]
    # equal distance cannot steal (<=)
    F1b = Frame.front(make_kps(np.float32([100, 101]), np.float32([100, 100])), np.stack([d_near, d_near]), 640, 480)
    prev = np.ascontiguousarray(np.stack([F1b.kps["x"], F1b.kps["y"]], 1), np.float32)
    n, m = oracle.search_for_initialization(F1b, F2, prev, 50, 0.9, False)
    assert n == 1 and m.tolist() == [0, -1]
