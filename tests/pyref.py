"""Second, independent restatement (pure Python, small cases only) of the reference's grid + matcher loops, written
directly from /root/reference/src/Frame.cc and src/ORBmatcher.cc.  It exists to cross-check oracle/match_oracle.cpp:
the reference ships no tests for these paths and its translation units cannot be compiled here, so two restatements
that were written separately and agree on random inputs are the best available pin ("parity unpinned")."""
import math

import numpy as np

F32 = np.float32
INT_MAX = 2**31 - 1
TH_HIGH, TH_LOW, HISTO = 100, 50, 30


def c_round(v):           # C round(): half away from zero
    v = float(v)
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


def build_grid(f):
    g = [[[] for _ in range(f.grows)] for _ in range(f.gcols)]
    for i in range(len(f.kps)):
        px = c_round((F32(f.kps["x"][i]) - F32(f.min_x)) * F32(f.inv_w))      # Frame.cc:550-551
        py = c_round((F32(f.kps["y"][i]) - F32(f.min_y)) * F32(f.inv_h))
        if px < 0 or px >= f.gcols or py < 0 or py >= f.grows:
            continue
        g[px][py].append(i)
    return g


def area(f, g, x, y, r, minL=-1, maxL=-1, inclusive=True):
    x, y, r = F32(x), F32(y), F32(r)
    out = []
    x0 = max(0, int(math.floor((x - F32(f.min_x) - r) * F32(f.inv_w))))
    if x0 >= f.gcols:
        return out
    x1 = min(f.gcols - 1, int(math.ceil((x - F32(f.min_x) + r) * F32(f.inv_w))))
    if x1 < 0:
        return out
    y0 = max(0, int(math.floor((y - F32(f.min_y) - r) * F32(f.inv_h))))
    if y0 >= f.grows:
        return out
    y1 = min(f.grows - 1, int(math.ceil((y - F32(f.min_y) + r) * F32(f.inv_h))))
    if y1 < 0:
        return out
    check = (minL > 0) or (maxL >= 0)
    e = 1 if inclusive else 0
    for ix in range(x0, x1 + e):
        for iy in range(y0, y1 + e):
            for j in g[ix][iy]:
                o = int(f.kps["octave"][j])
                if check:
                    if o < minL:
                        continue
                    if maxL >= 0 and o > maxL:
                        continue
                if abs(F32(f.kps["x"][j]) - x) < r and abs(F32(f.kps["y"][j]) - y) < r:
                    out.append(j)
    return out


def ham(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def three_maxima(h):
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i in range(HISTO):
        s = len(h[i])
        if s > m1:
            m3, m2, m1 = m2, m1, s
            i3, i2, i1 = i2, i1, i
        elif s > m2:
            m3, m2 = m2, s
            i3, i2 = i2, i
        elif s > m3:
            m3, i3 = s, i
    if F32(m2) < F32(0.1) * F32(m1):
        i2 = i3 = -1
    elif F32(m3) < F32(0.1) * F32(m1):
        i3 = -1
    return i1, i2, i3


def rot_bin(a1, a2):
    rot = F32(a1) - F32(a2)
    if rot < 0:
        rot = F32(rot + F32(360.0))
    b = c_round(F32(rot * F32(F32(1.0) / F32(HISTO))))
    return 0 if b == HISTO else b


def search_for_initialization(F1, F2, prev, window, ratio, ori):
    n1, n2 = len(F1.kps), len(F2.kps)
    m12 = [-1] * n1
    m21 = [-1] * n2
    md = [INT_MAX] * n2
    hist = [[] for _ in range(HISTO)]
    g2 = build_grid(F2)
    nm = 0
    for i1 in range(n1):
        if F1.kps["octave"][i1] > 0:
            continue
        cand = area(F2, g2, prev[i1, 0], prev[i1, 1], window, 0, 0, True)
        if not cand:
            continue
        best = best2 = INT_MAX
        bi = -1
        for i2 in cand:
            d = ham(F1.desc[i1], F2.desc[i2])
            if md[i2] <= d:
                continue
            if d < best:
                best2, best, bi = best, d, i2
            elif d < best2:
                best2 = d
        if best <= TH_LOW and F32(best) < F32(best2) * F32(ratio):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1
                nm -= 1
            m12[i1], m21[bi], md[bi] = bi, i1, best
            nm += 1
            if ori:
                hist[rot_bin(F1.kps["angle"][i1], F2.kps["angle"][bi])].append(i1)
    if ori:
        keep = three_maxima(hist)
        for i in range(HISTO):
            if i in keep:
                continue
            for i1 in hist[i]:
                if m12[i1] >= 0:
                    m12[i1] = -1
                    nm -= 1
    prev = prev.copy()
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev[i1] = (F2.kps["x"][m12[i1]], F2.kps["y"][m12[i1]])
    return nm, np.array(m12, np.int32), prev


def birdview_match(ref_kps, ref_desc, cur, window, ratio, ori):
    n = len(ref_kps)
    m12 = [-1] * n
    md = [INT_MAX] * n
    hist = [[] for _ in range(HISTO)]
    g = build_grid(cur)
    nm = 0
    for i1 in range(n):
        if ref_kps["octave"][i1] > 0:
            continue
        cand = area(cur, g, ref_kps["x"][i1], ref_kps["y"][i1], window, 0, 0, False)
        if not cand:
            continue
        best = best2 = INT_MAX
        bi = -1
        for i2 in cand:
            d = ham(ref_desc[i1], cur.desc[i2])
            if d < best:
                best2, best, bi = best, d, i2
            elif d < best2:
                best2 = d
        if best <= TH_LOW:
            if F32(best) < F32(best2) * F32(ratio):
                m12[i1], md[i1] = bi, best
                nm += 1
            if ori:
                hist[rot_bin(ref_kps["angle"][i1], cur.kps["angle"][bi])].append(i1)
    if ori:
        keep = three_maxima(hist)
        for i in range(HISTO):
            if i in keep:
                continue
            for i1 in hist[i]:
                if m12[i1] >= 0:
                    m12[i1] = -1
                    nm -= 1
    dm = [(i, m12[i], md[i]) for i in range(n) if m12[i] > 0]
    return nm, np.array(dm, np.int32).reshape(-1, 3)


def bird_map_point_match(pix, desc, cur, window, ratio):
    g = build_grid(cur)
    m12 = [-1] * len(pix)
    nm = 0
    for i1 in range(len(pix)):
        if np.isnan(pix[i1, 0]):
            continue
        cand = area(cur, g, pix[i1, 0], pix[i1, 1], window, -1, -1, False)
        if not cand:
            continue
        best = best2 = INT_MAX
        bi = -1
        for i2 in cand:
            d = ham(desc[i1], cur.desc[i2])
            if d < best:
                best2, best, bi = best, d, i2
            elif d < best2:
                best2 = d
        if best <= TH_LOW and F32(best) < F32(best2) * F32(ratio):
            m12[i1] = bi
            nm += 1
    return nm, np.array(m12, np.int32)


def search_by_projection_kf(cur, q_kps, proj, level, mp_desc, sf, th, th_dist, level_up, ori, taken0=None):
    """ORBmatcher.cc:1473-1600 (reloc: level_up=1, ORBdist, ori) and :291-404 (loop: level_up=0, TH_LOW, no ori)."""
    g = build_grid(cur)
    n = len(cur.kps)
    cur_mp = [-1] * n
    taken = [0] * n if taken0 is None else list(taken0)
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for i in range(len(level)):
        if np.isnan(proj[i, 0]):
            continue
        pl = int(level[i])
        radius = F32(th) * F32(sf[pl])
        cand = area(cur, g, proj[i, 0], proj[i, 1], radius, pl - 1, pl + level_up, True)
        if not cand:
            continue
        best, bi = 256, -1
        for i2 in cand:
            if taken[i2]:
                continue
            d = ham(mp_desc[i], cur.desc[i2])
            if d < best:
                best, bi = d, i2
        if best <= th_dist and bi >= 0:
            cur_mp[bi] = i
            taken[bi] = 1
            nm += 1
            if ori:
                hist[rot_bin(q_kps["angle"][i], cur.kps["angle"][bi])].append(bi)
    if ori:
        keep = three_maxima(hist)
        for b in range(HISTO):
            if b in keep:
                continue
            for k in hist[b]:
                cur_mp[k] = -2
                nm -= 1
    return nm, np.array(cur_mp, np.int32)


def search_by_projection_last(cur, last_kps, proj, mp_desc, sf, th, ori, taken0=None, has_obs=None):
    g = build_grid(cur)
    n = len(cur.kps)
    cur_mp = [-1] * n
    taken = [0] * n if taken0 is None else list(taken0)
    hist = [[] for _ in range(HISTO)]
    nm = 0
    for i in range(len(last_kps)):
        if np.isnan(proj[i, 0]):
            continue
        o = int(last_kps["octave"][i])
        radius = F32(th) * F32(sf[o])
        cand = area(cur, g, proj[i, 0], proj[i, 1], radius, o - 1, o + 1, True)
        if not cand:
            continue
        best, bi = 256, -1
        for i2 in cand:
            if taken[i2]:
                continue
            d = ham(mp_desc[i], cur.desc[i2])
            if d < best:
                best, bi = d, i2
        if best <= TH_HIGH:
            cur_mp[bi] = i
            if has_obs is None or has_obs[i]:
                taken[bi] = 1
            nm += 1
            if ori:
                hist[rot_bin(last_kps["angle"][i], cur.kps["angle"][bi])].append(bi)
    if ori:
        keep = three_maxima(hist)
        for b in range(HISTO):
            if b in keep:
                continue
            for k in hist[b]:
                cur_mp[k] = -2
                nm -= 1
    return nm, np.array(cur_mp, np.int32)


def search_by_projection_map(cur, sf, proj, level, viewcos, desc, th, ratio, taken0=None, has_obs=None):
    g = build_grid(cur)
    n = len(cur.kps)
    cur_mp = [-1] * n
    taken = [0] * n if taken0 is None else list(taken0)
    nm = 0
    for i in range(len(level)):
        lv = int(level[i])
        r = F32(2.5) if float(viewcos[i]) > 0.998 else F32(4.0)
        if th != 1.0:
            r = F32(r * F32(th))
        cand = area(cur, g, proj[i, 0], proj[i, 1], F32(r * F32(sf[lv])), lv - 1, lv, True)
        if not cand:
            continue
        best = best2 = 256
        bl = bl2 = -1
        bi = -1
        for k in cand:
            if taken[k]:
                continue
            d = ham(desc[i], cur.desc[k])
            if d < best:
                best2, best, bl2, bl, bi = best, d, bl, int(cur.kps["octave"][k]), k
            elif d < best2:
                bl2, best2 = int(cur.kps["octave"][k]), d
        if best <= TH_HIGH:
            if bl == bl2 and F32(best) > F32(ratio) * F32(best2):
                continue
            cur_mp[bi] = i
            if has_obs is None or has_obs[i]:
                taken[bi] = 1
            nm += 1
    return nm, np.array(cur_mp, np.int32)


def search_by_bow(kf_kps, kf_desc, kf_has_mp, kfv, f_kps, f_desc, ffv, ratio, ori):
    (ka, ks, ki), (fa, fs, fi) = kfv, ffv
    f_mp = [-1] * len(f_kps)
    hist = [[] for _ in range(HISTO)]
    nm = 0
    fmap = {int(n): j for j, n in enumerate(fa)}
    for a, node in enumerate(ka):                    # ordered merge == dictionary lookup on ascending ids
        if int(node) not in fmap:
            continue
        b = fmap[int(node)]
        for p in range(ks[a], ks[a + 1]):
            ikf = int(ki[p])
            if not kf_has_mp[ikf]:
                continue
            best = best2 = 256
            bi = -1
            for q in range(fs[b], fs[b + 1]):
                jf = int(fi[q])
                if f_mp[jf] >= 0:
                    continue
                d = ham(kf_desc[ikf], f_desc[jf])
                if d < best:
                    best2, best, bi = best, d, jf
                elif d < best2:
                    best2 = d
            if best <= TH_LOW and F32(best) < F32(ratio) * F32(best2):
                f_mp[bi] = ikf
                if ori:
                    hist[rot_bin(kf_kps["angle"][ikf], f_kps["angle"][bi])].append(bi)
                nm += 1
    if ori:
        keep = three_maxima(hist)
        for b in range(HISTO):
            if b in keep:
                continue
            for k in hist[b]:
                f_mp[k] = -2
                nm -= 1
    return nm, np.array(f_mp, np.int32)


def search_by_bow_kf(k1, d1, has1, fv1, k2, d2, has2, fv2, ratio, ori):
    """key frame vs key frame (ORBmatcher.cc:523-656): strict TH_LOW, matched key-frame-2 features are blocked."""
    (a1, s1, i1), (a2, s2, i2) = fv1, fv2
    m12 = [-1] * len(k1)
    used2 = [False] * len(k2)
    hist = [[] for _ in range(HISTO)]
    nm = 0
    map2 = {int(n): j for j, n in enumerate(a2)}
    for a, node in enumerate(a1):
        b = map2.get(int(node))
        if b is None:
            continue
        for p in range(s1[a], s1[a + 1]):
            x1 = int(i1[p])
            if not has1[x1]:
                continue
            best = best2 = 256
            bi = -1
            for q in range(s2[b], s2[b + 1]):
                x2 = int(i2[q])
                if used2[x2] or not has2[x2]:
                    continue
                d = ham(d1[x1], d2[x2])
                if d < best:
                    best2, best, bi = best, d, x2
                elif d < best2:
                    best2 = d
            if best < TH_LOW and F32(best) < F32(ratio) * F32(best2):
                m12[x1] = bi
                used2[bi] = True
                if ori:
                    hist[rot_bin(k1["angle"][x1], k2["angle"][bi])].append(x1)
                nm += 1
    if ori:
        keep = three_maxima(hist)
        for b in range(HISTO):
            if b not in keep:
                for x1 in hist[b]:
                    m12[x1] = -1
                    nm -= 1
    return nm, np.array(m12, np.int32)
