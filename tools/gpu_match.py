import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import oracle as O
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200.extractor import ORBextractor
from fishbirdeyevisualslam_b200.matcher import ORBmatcher, Frame, grid_assign, BaseXY2BirdPixel, transform_points

def check(name, ok): print(('OK  ' if ok else 'FAIL'), name)

# --- front pair, SearchForInitialization
h, w = 480, 640
a, b = synth.frame_pair_in_time(h, w, 11)
ex = ORBextractor(1000, 1.2, 8, 15, 5)
ka, da = ex(a); kb, db = ex(b)
F1 = Frame.front(ka, da, w, h, ex.GetScaleFactors()); F2 = Frame.front(kb, db, w, h, ex.GetScaleFactors())
s_g, i_g = F2.AssignFeaturesToGrid(); s_o, i_o = O.grid_assign(kb, F2.min_x, F2.min_y, F2.inv_w, F2.inv_h, 64, 48)
check('grid front', np.array_equal(s_g, s_o) and np.array_equal(i_g, i_o))
for ratio, ori, win in [(0.9, True, 100), (0.9, False, 100), (0.7, True, 30), (0.9, True, 400)]:
    m = ORBmatcher(ratio, ori)
    pm_g = np.ascontiguousarray(np.stack([ka['x'], ka['y']], 1), np.float32); pm_o = pm_g.copy()
    n_g, m_g = m.SearchForInitialization(F1, F2, pm_g, win)
    n_o, m_o = O.search_for_initialization(F1, F2, pm_o, win, ratio, ori)
    check('SearchForInitialization r=%.1f ori=%d win=%d n=%d/%d' % (ratio, ori, win, n_g, n_o), n_g == n_o and np.array_equal(m_g, m_o) and np.array_equal(pm_g, pm_o))
    # second call with updated prev (as Tracking does)
    n_g, m_g = m.SearchForInitialization(F1, F2, pm_g, win); n_o, m_o = O.search_for_initialization(F1, F2, pm_o, win, ratio, ori)
    check('  repeat n=%d/%d' % (n_g, n_o), n_g == n_o and np.array_equal(m_g, m_o) and np.array_equal(pm_g, pm_o))

# --- bird pair, BirdviewMatch
bh = bw = 384
ba, bb = synth.frame_pair_in_time(bh, bw, 21)
exb = ORBextractor(1000, 1.2, 8, 15, 5)
kba, dba = exb(ba); kbb, dbb = exb(bb)
FB = Frame.bird(kbb, dbb, bw, bh)
s_g, i_g = FB.AssignFeaturesToGrid(); s_o, i_o = O.grid_assign(kbb, 0, 0, FB.inv_w, FB.inv_h, 32, 32)
check('grid bird', np.array_equal(s_g, s_o) and np.array_equal(i_g, i_o))
for ratio, ori, win in [(0.9, True, 10), (0.9, False, 10), (0.8, True, 25)]:
    m = ORBmatcher(ratio, ori)
    n_g, d_g = m.BirdviewMatch(FB, kba, dba, win); n_o, d_o = O.birdview_match(kba, dba, FB, win, ratio, ori)
    check('BirdviewMatch r=%.1f ori=%d win=%d n=%d/%d dm=%d' % (ratio, ori, win, n_g, n_o, len(d_g)), n_g == n_o and np.array_equal(d_g, d_o))

# --- BirdMapPointMatch: 20k map points vs bird keypoints (config C3 shape)
rng = np.random.default_rng(5)
nmp = 20000
src = rng.integers(0, len(kbb), nmp)
mp_desc = dbb[src].copy()
flips = rng.integers(0, 41, nmp)
for i in range(nmp):
    bits = rng.choice(256, flips[i], replace=False)
    for bt in bits: mp_desc[i, bt >> 3] ^= (1 << (bt & 7))
pix = np.stack([kbb['x'][src], kbb['y'][src]], 1).astype(np.float32) + rng.normal(0, 2.5, (nmp, 2)).astype(np.float32)
pix[rng.random(nmp) < 0.05, 0] = np.nan
m = ORBmatcher(0.9, True)
v = FB.view()
import ctypes as C
from fishbirdeyevisualslam_b200._lib import ptr, check as chk
m12 = np.full(nmp, -1, np.int32); n = C.c_int32()
chk(m._L.fbe_bird_map_point_match(m._h, ptr(pix), ptr(mp_desc), nmp, C.byref(v), 10, ptr(m12), C.byref(n)))
n_o, m_o = O.bird_map_point_match(pix, mp_desc, FB, 10, 0.9)
check('BirdMapPointMatch n=%d/%d' % (n.value, n_o), n.value == n_o and np.array_equal(m12, m_o))

# --- SearchByProjection (last frame): identity motion, project = last keypoint position + jitter
sf = ex.GetScaleFactors()
proj = np.stack([ka['x'], ka['y']], 1).astype(np.float32) + rng.normal(0, 3, (len(ka), 2)).astype(np.float32)
proj[rng.random(len(ka)) < 0.2, 0] = np.nan
taken = (rng.random(len(kb)) < 0.1).astype(np.uint8)
for th, ori in [(15, True), (30, True), (15, False)]:
    m = ORBmatcher(0.9, ori)
    n_g, c_g = m.SearchByProjectionLast(F2, ka, proj, da, th, cur_taken=taken)
    n_o, c_o = O.search_by_projection_last(F2, ka, proj, da, sf, th, ori, cur_taken=taken)
    check('SearchByProjection(last) th=%d ori=%d n=%d/%d' % (th, ori, n_g, n_o), n_g == n_o and np.array_equal(c_g, c_o))
# with some map points lacking observations (do not block)
obs = (rng.random(len(ka)) < 0.7).astype(np.uint8)
m = ORBmatcher(0.9, True)
n_g, c_g = m.SearchByProjectionLast(F2, ka, proj, da, 15, cur_taken=taken, last_has_obs=obs)
n_o, c_o = O.search_by_projection_last(F2, ka, proj, da, sf, 15, True, cur_taken=taken, last_has_obs=obs)
check('SearchByProjection(last, no-obs) n=%d/%d' % (n_g, n_o), n_g == n_o and np.array_equal(c_g, c_o))

# --- SearchByProjection (local map)
nmap = 6000
src = rng.integers(0, len(kb), nmap)
mdesc = db[src].copy()
for i in range(nmap):
    for bt in rng.choice(256, rng.integers(0, 60), replace=False): mdesc[i, bt >> 3] ^= (1 << (bt & 7))
mproj = np.stack([kb['x'][src], kb['y'][src]], 1).astype(np.float32) + rng.normal(0, 2, (nmap, 2)).astype(np.float32)
mlevel = np.clip(kb['octave'][src] + rng.integers(-1, 2, nmap), 0, 7).astype(np.int32)
mcos = rng.uniform(0.99, 1.0, nmap).astype(np.float32)
for th, ratio in [(1.0, 0.8), (3.0, 0.8), (5.0, 0.6)]:
    m = ORBmatcher(ratio, True)
    n_g, c_g = m.SearchByProjectionMap(F2, mproj, mlevel, mcos, mdesc, th, cur_taken=taken)
    n_o, c_o = O.search_by_projection_map(F2, sf, mproj, mlevel, mcos, mdesc, th, ratio, cur_taken=taken)
    check('SearchByProjection(map) th=%.0f r=%.1f n=%d/%d' % (th, ratio, n_g, n_o), n_g == n_o and np.array_equal(c_g, c_o))

# --- SearchByBoW with synthetic node assignment
def featvec(n, nnodes, seed):
    r = np.random.default_rng(seed)
    node = r.integers(0, nnodes, n)
    ids = np.unique(node)
    start = [0]; items = []
    for t in ids:
        it = np.nonzero(node == t)[0]; items.extend(it.tolist()); start.append(len(items))
    return ids.astype(np.int32) * 3 + 1, np.array(start, np.int32), np.array(items, np.int32)
# node of a frame feature correlates with its kf counterpart: use position-hash so that true matches share nodes
def featvec_pos(k, shift):
    node = ((np.floor((k['x'] - shift[0]) / 80).astype(int) * 16 + np.floor((k['y'] - shift[1]) / 80).astype(int)) * 8 + k['octave'])
    ids = np.unique(node); start = [0]; items = []
    for t in ids:
        it = np.nonzero(node == t)[0]; items.extend(it.tolist()); start.append(len(items))
    return ids.astype(np.int32), np.array(start, np.int32), np.array(items, np.int32)
kfv = featvec_pos(ka, (0, 0)); ffv = featvec_pos(kb, (3, 2))
has_mp = (rng.random(len(ka)) < 0.8).astype(np.uint8)
for ratio, ori in [(0.7, True), (0.75, False)]:
    m = ORBmatcher(ratio, ori)
    n_g, f_g = m.SearchByBoW(ka, da, has_mp, kfv, F2, ffv)
    n_o, f_o = O.search_by_bow(ka, da, has_mp, kfv, kb, db, ffv, ratio, ori)
    check('SearchByBoW r=%.2f ori=%d n=%d/%d' % (ratio, ori, n_g, n_o), n_g == n_o and np.array_equal(f_g, f_o))

# --- brute force
q = rng.integers(0, 256, (3000, 32), dtype=np.uint8); t = rng.integers(0, 256, (5000, 32), dtype=np.uint8)
t[100] = q[5]; t[4000] = q[5]
m = ORBmatcher(0.9, True)
g = m.BruteForceTop2(q, t); o = O.bruteforce_top2(q, t)
check('bruteforce', all(np.array_equal(x, y) for x, y in zip(g, o)))
check('hamming', all(ORBmatcher.DescriptorDistance(q[i], t[i]) == O.hamming256(q[i], t[i]) for i in range(50)))
