"""ORACLE loader (test infrastructure).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this module; the product package never does."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(quiet: bool = True) -> None:
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_ext_create.restype = C.c_void_p
        _lib.orc_ext_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_cv_round_f.argtypes = [C.c_float]
        _lib.orc_cv_round_d.argtypes = [C.c_double]
        for name in ("orc_ext_destroy", "orc_ext_tables", "orc_ext_run", "orc_ext_result", "orc_ext_level_size",
                     "orc_ext_level_padded", "orc_ext_level_blurred", "orc_ext_candidates", "orc_ext_level_nkeys",
                     "orc_ext_cell_stats"):
            getattr(_lib, name).argtypes = None
    return _lib


def host_runs_timing_build() -> bool:
    """The -O3 -march=x86-64-v3 timing variants (oracle/Makefile `timing`) need AVX2, FMA and BMI2 on the host they RUN on."""
    try:
        flags = set()
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = set(line.split(":", 1)[1].split())
                break
        return {"avx2", "fma", "bmi2"} <= flags
    except OSError:
        return False


_ref_o3 = None


def ref(timing: bool = False):
    """The reference's own ORBextractor.cc compiled verbatim, or None when oracle/_ref was not built.
    timing=True: the -O3 -march=x86-64-v3 variant for bench.py's CPU legs (None when absent or the host lacks AVX2)."""
    global _ref, _ref_o3
    if timing:
        if _ref_o3 is None:
            path = os.path.join(HERE, "_ref", "libfbe_ref_o3.so")
            if not os.path.exists(path) or not host_runs_timing_build():
                return None
            _ref_o3 = C.CDLL(path)
            _ref_o3.ref_extractor_create.restype = C.c_void_p
            _ref_o3.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        return _ref_o3
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libfbe_ref.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        _ref.ref_extractor_create.restype = C.c_void_p
        _ref.ref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    return _ref


class OracleExtractor:
    """Restated ORBextractor (oracle/orb_oracle.cpp)."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=15, min_th=5):
        self.L = lib()
        self.nlevels = nlevels
        self.h = C.c_void_p(self.L.orc_ext_create(nfeatures, scale, nlevels, ini_th, min_th))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_ext_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        per = np.zeros(n, np.int32)
        umax = np.zeros(16, np.int32)
        self.L.orc_ext_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(per), _p(umax))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, per_level=per, umax=umax)

    def __call__(self, img: np.ndarray):
        img = np.ascontiguousarray(img, np.uint8)
        n = self.L.orc_ext_run(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0])
        kps = np.zeros(n, KP_DTYPE)
        desc = np.zeros((n, 32), np.uint8)
        self.boundary = np.zeros(n, np.uint8)
        if n:
            self.L.orc_ext_result(self.h, _p(kps), _p(desc), _p(self.boundary))
        return kps, desc

    def level_size(self, l):
        w, h = C.c_int32(), C.c_int32()
        self.L.orc_ext_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_padded(self, l):
        w, h = self.level_size(l)
        out = np.zeros((h + 38, w + 38), np.uint8)
        self.L.orc_ext_level_padded(self.h, l, _p(out))
        return out

    def level_blurred(self, l):
        w, h = self.level_size(l)
        out = np.zeros((h, w), np.uint8)
        ok = self.L.orc_ext_level_blurred(self.h, l, _p(out))
        return out if ok else None

    def candidates(self, l):
        n = self.L.orc_ext_candidates(self.h, l, None, 0)
        out = np.zeros((n, 3), np.int32)
        if n:
            self.L.orc_ext_candidates(self.h, l, _p(out), n)
        return out

    def level_nkeys(self, l):
        return self.L.orc_ext_level_nkeys(self.h, l)

    def cell_stats(self, l):
        a, b = C.c_int32(), C.c_int32()
        self.L.orc_ext_cell_stats(self.h, l, C.byref(a), C.byref(b))
        return a.value, b.value


class RefExtractor:
    """The reference's ORBextractor (verbatim TU) through oracle/_ref/libfbe_ref.so."""

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=15, min_th=5, timing=False):
        self.R = ref(timing)
        if self.R is None:
            raise RuntimeError("oracle/_ref/libfbe_ref%s.so not built" % ("_o3" if timing else ""))
        self.nlevels = nlevels
        self.cap = max(4 * nfeatures, 4096)
        self.h = C.c_void_p(self.R.ref_extractor_create(nfeatures, scale, nlevels, ini_th, min_th))

    def __del__(self):
        if getattr(self, "h", None):
            self.R.ref_extractor_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        self.R.ref_extractor_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2)

    def __call__(self, img: np.ndarray):
        img = np.ascontiguousarray(img, np.uint8)
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = self.R.ref_extract(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], _p(kps), _p(desc), self.cap)
        assert n <= self.cap
        return kps[:n].copy(), desc[:n].copy()

    def pyramid_level(self, img: np.ndarray, level: int):
        img = np.ascontiguousarray(img, np.uint8)
        r, c = C.c_int(), C.c_int()
        self.R.ref_pyramid_level(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], level, None, 0, C.byref(r), C.byref(c))
        out = np.zeros((r.value + 38, c.value + 38), np.uint8)
        self.R.ref_pyramid_level(self.h, _p(img), img.shape[0], img.shape[1], img.strides[0], level, _p(out), out.strides[0], C.byref(r), C.byref(c))
        return out


class OFrameView(C.Structure):
    _fields_ = [("kps", C.c_void_p), ("desc", C.c_void_p), ("n", C.c_int32),
                ("min_x", C.c_float), ("min_y", C.c_float), ("inv_w", C.c_float), ("inv_h", C.c_float),
                ("gcols", C.c_int32), ("grows", C.c_int32)]


def _view(f):
    """f: any object with kps, desc, min_x, min_y, inv_w, inv_h, gcols, grows (e.g. the package's matcher.Frame)."""
    return OFrameView(f.kps.ctypes.data if len(f.kps) else None, f.desc.ctypes.data if len(f.kps) else None, len(f.kps),
                      f.min_x, f.min_y, f.inv_w, f.inv_h, f.gcols, f.grows)


def _fp(x):
    return C.c_float(float(x))


def hamming256(a, b):
    return lib().orc_hamming256(_p(np.ascontiguousarray(a, np.uint8)), _p(np.ascontiguousarray(b, np.uint8)))


def grid_assign(kps, min_x, min_y, inv_w, inv_h, gcols, grows):
    kps = np.ascontiguousarray(kps)
    start = np.zeros(gcols * grows + 1, np.int32)
    items = np.zeros(max(len(kps), 1), np.int32)
    n = lib().orc_grid_assign(_p(kps), len(kps), _fp(min_x), _fp(min_y), _fp(inv_w), _fp(inv_h), gcols, grows, _p(start), _p(items))
    return start, items[:n].copy()


def features_in_area(f, x, y, r, min_level=-1, max_level=-1, upper_inclusive=True):
    v = _view(f)
    out = np.zeros(max(len(f.kps), 1), np.int32)
    n = lib().orc_features_in_area(C.byref(v), _fp(x), _fp(y), _fp(r), min_level, max_level, int(upper_inclusive), _p(out), len(out))
    return out[:n].copy()


def search_for_initialization(f1, f2, prev_matched, window, nn_ratio, check_ori):
    v1, v2 = _view(f1), _view(f2)
    m12 = np.full(len(f1.kps), -1, np.int32)
    n = lib().orc_search_for_initialization(C.byref(v1), C.byref(v2), _p(prev_matched), _p(m12), int(window), _fp(nn_ratio), int(check_ori))
    return n, m12


def birdview_match(ref_kps, ref_desc, cur, window, nn_ratio, check_ori):
    v = _view(cur)
    ref_kps = np.ascontiguousarray(ref_kps); ref_desc = np.ascontiguousarray(ref_desc)
    dm = np.zeros((max(len(ref_kps), 1), 3), np.int32)
    nd = C.c_int32()
    n = lib().orc_birdview_match(_p(ref_kps), _p(ref_desc), len(ref_kps), C.byref(v), int(window), _fp(nn_ratio), int(check_ori), _p(dm), C.byref(nd))
    return n, dm[:nd.value].copy()


def bird_map_point_match(mp_pix, mp_desc, cur, window, nn_ratio):
    v = _view(cur)
    mp_pix = np.ascontiguousarray(mp_pix, np.float32); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    m12 = np.full(len(mp_pix), -1, np.int32)
    n = lib().orc_bird_map_point_match(_p(mp_pix), _p(mp_desc), len(mp_pix), C.byref(v), int(window), _fp(nn_ratio), _p(m12))
    return n, m12


def search_by_projection_last(cur, last_kps, last_proj, last_mp_desc, scale_factors, th, check_ori, cur_taken=None, last_has_obs=None):
    v = _view(cur)
    last_kps = np.ascontiguousarray(last_kps); last_proj = np.ascontiguousarray(last_proj, np.float32)
    last_mp_desc = np.ascontiguousarray(last_mp_desc, np.uint8); sf = np.ascontiguousarray(scale_factors, np.float32)
    cur_mp = np.full(max(len(cur.kps), 1), -1, np.int32)
    tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
    ho = None if last_has_obs is None else np.ascontiguousarray(last_has_obs, np.uint8)
    n = lib().orc_search_by_projection_last(C.byref(v), _p(last_kps), _p(last_proj), _p(last_mp_desc), len(last_kps), _p(sf),
                                            None if tk is None else _p(tk), None if ho is None else _p(ho), _fp(th), int(check_ori), _p(cur_mp))
    return n, cur_mp[:len(cur.kps)]


def search_by_projection_kf(cur, q_kps, proj, level, mp_desc, scale_factors, th, th_dist, level_up, check_ori, cur_taken=None):
    """reloc (level_up=1, th_dist=ORBdist, ori) / loop-closing (level_up=0, th_dist=TH_LOW, no ori) SearchByProjection."""
    v = _view(cur)
    q_kps = np.ascontiguousarray(q_kps); proj = np.ascontiguousarray(proj, np.float32)
    level = np.ascontiguousarray(level, np.int32)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8); sf = np.ascontiguousarray(scale_factors, np.float32)
    cur_mp = np.full(max(len(cur.kps), 1), -1, np.int32)
    tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
    n = lib().orc_search_by_projection_kf(C.byref(v), _p(q_kps), _p(proj), _p(level), _p(mp_desc), len(level), _p(sf),
                                          None if tk is None else _p(tk), _fp(th), int(th_dist), int(level_up), int(check_ori), _p(cur_mp))
    return n, cur_mp[:len(cur.kps)]


def search_by_projection_map(cur, scale_factors, mp_proj, mp_level, mp_viewcos, mp_desc, th, nn_ratio, cur_taken=None, mp_has_obs=None):
    v = _view(cur)
    sf = np.ascontiguousarray(scale_factors, np.float32); mp_proj = np.ascontiguousarray(mp_proj, np.float32)
    mp_level = np.ascontiguousarray(mp_level, np.int32); mp_viewcos = np.ascontiguousarray(mp_viewcos, np.float32)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    cur_mp = np.full(max(len(cur.kps), 1), -1, np.int32)
    tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
    ho = None if mp_has_obs is None else np.ascontiguousarray(mp_has_obs, np.uint8)
    n = lib().orc_search_by_projection_map(C.byref(v), _p(sf), _p(mp_proj), _p(mp_level), _p(mp_viewcos), _p(mp_desc), len(mp_level),
                                           None if tk is None else _p(tk), None if ho is None else _p(ho), _fp(th), _fp(nn_ratio), _p(cur_mp))
    return n, cur_mp[:len(cur.kps)]


def search_by_bow(kf_kps, kf_desc, kf_has_mp, kf_fv, f_kps, f_desc, f_fv, nn_ratio, check_ori):
    kf_kps = np.ascontiguousarray(kf_kps); kf_desc = np.ascontiguousarray(kf_desc, np.uint8); kf_has_mp = np.ascontiguousarray(kf_has_mp, np.uint8)
    f_kps = np.ascontiguousarray(f_kps); f_desc = np.ascontiguousarray(f_desc, np.uint8)
    ka, kb, kc = (np.ascontiguousarray(a, np.int32) for a in kf_fv)
    fa, fb, fc = (np.ascontiguousarray(a, np.int32) for a in f_fv)
    f_mp = np.full(max(len(f_kps), 1), -1, np.int32)
    n = lib().orc_search_by_bow(_p(kf_kps), _p(kf_desc), len(kf_kps), _p(kf_has_mp), _p(ka), _p(kb), _p(kc), len(ka),
                                _p(f_kps), _p(f_desc), len(f_kps), _p(fa), _p(fb), _p(fc), len(fa), _fp(nn_ratio), int(check_ori), _p(f_mp))
    return n, f_mp[:len(f_kps)]


def search_by_bow_kf(k1, d1, has1, fv1, k2, d2, has2, fv2, nn_ratio, check_ori, _L=None, _fn="orc_search_by_bow_kf"):
    """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (loop closing) -> (nmatches, matches12[n1] = key-frame-2 feature or -1)."""
    k1 = np.ascontiguousarray(k1); d1 = np.ascontiguousarray(d1, np.uint8); has1 = np.ascontiguousarray(has1, np.uint8)
    k2 = np.ascontiguousarray(k2); d2 = np.ascontiguousarray(d2, np.uint8); has2 = np.ascontiguousarray(has2, np.uint8)
    a1, b1, c1 = (np.ascontiguousarray(a, np.int32) for a in fv1)
    a2, b2, c2 = (np.ascontiguousarray(a, np.int32) for a in fv2)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = getattr(_L or lib(), _fn)(_p(k1), _p(d1), len(k1), _p(has1), _p(a1), _p(b1), _p(c1), len(a1), _p(k2), _p(d2), len(k2), _p(has2),
                                 _p(a2), _p(b2), _p(c2), len(a2), _fp(nn_ratio), int(check_ori), _p(m12))
    return n, m12[:len(k1)]


class FrustumViewC(C.Structure):
    _fields_ = [("Rcw", C.c_float * 9), ("tcw", C.c_float * 3), ("Ow", C.c_float * 3), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float),
                ("max_y", C.c_float), ("mbf", C.c_float), ("log_scale_factor", C.c_float), ("n_levels", C.c_int32)]


def frustum_view(Rcw, tcw, Ow, K, bounds, mbf, log_scale_factor, n_levels):
    v = FrustumViewC()
    v.Rcw[:] = [float(x) for x in np.asarray(Rcw, np.float32).ravel()]
    v.tcw[:] = [float(x) for x in np.asarray(tcw, np.float32).ravel()]
    v.Ow[:] = [float(x) for x in np.asarray(Ow, np.float32).ravel()]
    v.fx, v.fy, v.cx, v.cy = (float(np.float32(x)) for x in K)
    v.min_x, v.max_x, v.min_y, v.max_y = (float(np.float32(x)) for x in bounds)
    v.mbf, v.log_scale_factor, v.n_levels = float(np.float32(mbf)), float(np.float32(log_scale_factor)), int(n_levels)
    return v


def is_in_frustum(view, pos, normal, min_dist, max_dist, cos_limit):
    """Frame::isInFrustum for n map points -> dict(in_view, proj, proj_xr, level, view_cos, level_boundary)."""
    pos = np.ascontiguousarray(pos, np.float32); normal = np.ascontiguousarray(normal, np.float32)
    mn = np.ascontiguousarray(min_dist, np.float32); mx = np.ascontiguousarray(max_dist, np.float32)
    n = len(pos)
    o = dict(in_view=np.zeros(n, np.uint8), proj=np.zeros((n, 2), np.float32), proj_xr=np.zeros(n, np.float32),
             level=np.zeros(n, np.int32), view_cos=np.zeros(n, np.float32), level_boundary=np.zeros(n, np.uint8))
    lib().orc_is_in_frustum(C.byref(view), _p(pos), _p(normal), _p(mn), _p(mx), n, _fp(cos_limit), _p(o["in_view"]), _p(o["proj"]),
                            _p(o["proj_xr"]), _p(o["level"]), _p(o["view_cos"]), _p(o["level_boundary"]))
    return o


def search_for_triangulation(k1, d1, has_mp1, stereo1, fv1, k2, d2, has_mp2, stereo2, fv2, F12, epipole, scale2, sigma2, only_stereo,
                             nn_ratio, check_ori, _L=None):
    """SearchForTriangulation -> (nmatches, matches12[n1])."""
    k1 = np.ascontiguousarray(k1); d1 = np.ascontiguousarray(d1, np.uint8); k2 = np.ascontiguousarray(k2); d2 = np.ascontiguousarray(d2, np.uint8)
    h1 = np.ascontiguousarray(has_mp1, np.uint8); h2 = np.ascontiguousarray(has_mp2, np.uint8)
    s1 = np.ascontiguousarray(stereo1, np.uint8); s2 = np.ascontiguousarray(stereo2, np.uint8)
    a1, b1, c1 = (np.ascontiguousarray(a, np.int32) for a in fv1)
    a2, b2, c2 = (np.ascontiguousarray(a, np.int32) for a in fv2)
    F = np.ascontiguousarray(F12, np.float32).ravel(); sc = np.ascontiguousarray(scale2, np.float32); sg = np.ascontiguousarray(sigma2, np.float32)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    if _L is None:
        skip1 = (h1 != 0) | ((s1 == 0) if only_stereo else False)
        skip2 = (h2 != 0) | ((s2 == 0) if only_stereo else False)
        skip1 = np.ascontiguousarray(skip1, np.uint8); skip2 = np.ascontiguousarray(skip2, np.uint8)
        n = lib().orc_search_for_triangulation(_p(k1), _p(d1), len(k1), _p(skip1), _p(s1), _p(a1), _p(b1), _p(c1), len(a1), _p(k2), _p(d2),
                                               len(k2), _p(skip2), _p(s2), _p(a2), _p(b2), _p(c2), len(a2), _p(F), _fp(epipole[0]),
                                               _fp(epipole[1]), _p(sc), _p(sg), int(check_ori), _p(m12))
    else:
        n = _L.refm_search_for_triangulation(_p(k1), _p(d1), len(k1), _p(h1), _p(s1), _p(a1), _p(b1), _p(c1), len(a1), _p(k2), _p(d2), len(k2),
                                             _p(h2), _p(s2), _p(a2), _p(b2), _p(c2), len(a2), _p(F), _fp(epipole[0]), _fp(epipole[1]),
                                             _p(sc), _p(sg), int(only_stereo), _fp(nn_ratio), int(check_ori), _p(m12))
    return n, m12[:len(k1)]


def fuse_search(kf, uright, inv_sigma2, proj, proj_ur, level, radius, mp_desc, check_chi2):
    """The candidate search of ORBmatcher::Fuse -> (best_idx[n], best_dist[n])."""
    proj = np.ascontiguousarray(proj, np.float32); level = np.ascontiguousarray(level, np.int32); radius = np.ascontiguousarray(radius, np.float32)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8); inv = np.ascontiguousarray(inv_sigma2, np.float32)
    ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    pur = None if proj_ur is None else np.ascontiguousarray(proj_ur, np.float32)
    n = len(level)
    bi = np.full(max(n, 1), -1, np.int32); bd = np.zeros(max(n, 1), np.int32)
    v = _view(kf)
    lib().orc_fuse_search(C.byref(v), None if ur is None else _p(ur), _p(inv), _p(proj), None if pur is None else _p(pur), _p(level), _p(radius),
                          _p(mp_desc), n, int(check_chi2), _p(bi), _p(bd))
    return bi[:n], bd[:n]


def fuse(kf, uright, inv_sigma2, proj, level, mp_desc, mp_nobs, mp_bad, mp_in_kf, occ, occ_nobs, occ_bad, th, overload, _L=None):
    """ORBmatcher::Fuse overload 1 / 2 on an abstract map state (bf = 0: ur = u) -> (nFused, act[n], slot[n])."""
    proj = np.ascontiguousarray(proj, np.float32); level = np.ascontiguousarray(level, np.int32); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    inv = np.ascontiguousarray(inv_sigma2, np.float32); ur = None if uright is None else np.ascontiguousarray(uright, np.float32)
    nobs = np.ascontiguousarray(mp_nobs, np.int32); bad = np.ascontiguousarray(mp_bad, np.uint8); inkf = np.ascontiguousarray(mp_in_kf, np.uint8)
    occ = np.ascontiguousarray(occ, np.int32); onobs = np.ascontiguousarray(occ_nobs, np.int32); obad = np.ascontiguousarray(occ_bad, np.uint8)
    n = len(level)
    act = np.zeros(max(n, 1), np.int32); slot = np.full(max(n, 1), -1, np.int32)
    v = _view(kf)
    sf = np.ascontiguousarray(kf.scale_factors, np.float32)
    if _L is None:
        radius = (np.float32(th) * sf[level]).astype(np.float32)
        pur = np.ascontiguousarray(proj[:, 0])
        nf = lib().orc_fuse(C.byref(v), None if ur is None else _p(ur), _p(inv), _p(proj), _p(pur), _p(level), _p(radius), _p(mp_desc), _p(nobs),
                            _p(bad), _p(inkf), n, _p(occ), _p(onobs), _p(obad), int(overload), _p(act), _p(slot))
    else:
        nf = _L.refm_fuse(C.byref(v), None if ur is None else _p(ur), _p(sf), _p(inv), _p(proj), _p(level), _p(mp_desc), _p(nobs), _p(bad), _p(inkf),
                          n, _p(occ), _p(onobs), _p(obad), len(onobs), _fp(th), int(overload), _p(act), _p(slot))
    return nf, act[:n], slot[:n]


def search_by_sim3(F1, F2, pos1, lvl1, desc1, pos2, lvl2, desc2, pre12, t12, th, _L=None):
    """SearchBySim3 with identity poses, s12 = 1, R12 = I, t12 = (tx, ty, 0): pos1 / pos2 are the image positions (u, v) of the
    map points hanging on the features of key frame 1 / 2 (NaN = none), lvl their scale level.  -> (nFound, matches12[n1])."""
    pos1 = np.ascontiguousarray(pos1, np.float32); pos2 = np.ascontiguousarray(pos2, np.float32)
    lvl1 = np.ascontiguousarray(lvl1, np.int32); lvl2 = np.ascontiguousarray(lvl2, np.int32)
    desc1 = np.ascontiguousarray(desc1, np.uint8); desc2 = np.ascontiguousarray(desc2, np.uint8)
    pre12 = np.ascontiguousarray(pre12, np.int32)
    sf = np.ascontiguousarray(F1.scale_factors, np.float32)
    m12 = np.full(max(F1.N, 1), -1, np.int32)
    v1, v2 = _view(F1), _view(F2)
    tx, ty = np.float32(t12[0]), np.float32(t12[1])
    if _L is not None:
        n = _L.refm_search_by_sim3(C.byref(v1), C.byref(v2), _p(sf), _p(pos1), _p(lvl1), _p(desc1), _p(pos2), _p(lvl2), _p(desc2), _p(pre12),
                                   _fp(tx), _fp(ty), _fp(th), _p(m12))
        return n, m12[:F1.N]
    # the projections of :1160-1176 / :1233-1249 for this camera set-up: key frame 2 sees p - t12, key frame 1 sees p + t12
    pre = np.where((pre12 >= 0) & ~np.isnan(pos2[np.maximum(pre12, 0), 0]), pre12, -1).astype(np.int32)
    p12 = (pos1 - np.array([tx, ty], np.float32)).astype(np.float32)
    p21 = (pos2 + np.array([tx, ty], np.float32)).astype(np.float32)
    w, h = np.float32(F1.gcols / F1.inv_w), np.float32(F1.grows / F1.inv_h)
    for p in (p12, p21):
        out = ~((p[:, 0] >= F1.min_x) & (p[:, 0] < F1.min_x + w) & (p[:, 1] >= F1.min_y) & (p[:, 1] < F1.min_y + h))
        p[out, 0] = np.nan
    p12[pre >= 0, 0] = np.nan
    p21[pre[pre >= 0], 0] = np.nan
    n = lib().orc_search_by_sim3(C.byref(v1), C.byref(v2), _p(sf), _p(p12), _p(lvl1), _p(desc1), _p(p21), _p(lvl2), _p(desc2), _p(pre), _fp(th),
                                 _p(m12))
    return n, m12[:F1.N]


def check_models(k1, k2, matches, A, B, sigma, homography, _L=None):
    """Initializer::CheckHomography (A = H21, B = H12) / CheckFundamental (A = F21) for K hypotheses -> (scores[K], inliers[K, n])."""
    k1 = np.ascontiguousarray(k1); k2 = np.ascontiguousarray(k2); matches = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 9); K, n = len(A), len(matches)
    Bp = None if B is None else np.ascontiguousarray(B, np.float32).reshape(-1, 9)
    scores = np.zeros(max(K, 1), np.float32); inl = np.zeros((max(K, 1), max(n, 1)), np.uint8)
    inl_flat = np.zeros(max(K * n, 1), np.uint8)
    if _L is None:
        lib().orc_check_models(_p(k1), _p(k2), _p(matches), n, _p(A), None if Bp is None else _p(Bp), K, _fp(sigma), int(homography), _p(scores),
                               _p(inl_flat))
    else:
        _L.refm_check_models(_p(k1), len(k1), _p(k2), len(k2), _p(matches), n, _p(A), None if Bp is None else _p(Bp), K, _fp(sigma),
                             int(homography), _p(scores), _p(inl_flat))
    return scores[:K], inl_flat[:K * n].reshape(K, n)


def bow_transform(L, parent, is_word, ndesc, nweight, desc, levelsup):
    """Per-feature part of DBoW2 TemplatedVocabulary::transform -> (word_id[n], node_id[n], weight[n])."""
    parent = np.ascontiguousarray(parent, np.int32); is_word = np.ascontiguousarray(is_word, np.uint8)
    ndesc = np.ascontiguousarray(ndesc, np.uint8); nweight = np.ascontiguousarray(nweight, np.float64)
    desc = np.ascontiguousarray(desc, np.uint8)
    n = len(desc)
    w = np.zeros(max(n, 1), np.int32); nd = np.zeros(max(n, 1), np.int32); wt = np.zeros(max(n, 1), np.float64)
    lib().orc_bow_transform(int(L), _p(parent), _p(is_word), _p(ndesc), _p(nweight), len(parent), _p(desc), n, int(levelsup), _p(w), _p(nd), _p(wt))
    return w[:n], nd[:n], wt[:n]


_refv = None


def refvoc():
    """ctypes handle of the verbatim DBoW2 vocabulary build (oracle/_ref/libfbe_refvoc.so) or None."""
    global _refv
    if _refv is None:
        path = os.path.join(HERE, "_ref", "libfbe_refvoc.so")
        if not os.path.exists(path):
            return None
        _refv = C.CDLL(path)
        _refv.refv_load_text.restype = C.c_void_p
        _refv.refv_load_text.argtypes = [C.c_char_p]
        _refv.refv_free.argtypes = [C.c_void_p]
        _refv.refv_size.argtypes = [C.c_void_p]
    return _refv


def ref_voc_transform(text_path, desc, levelsup):
    """The reference's own loadFromTextFile + transform(features, BowVector&, FeatureVector&, levelsup)
    -> (bow ids, bow values, (fv node ids, start, items), number of words in the vocabulary)."""
    L_ = refvoc()
    h = L_.refv_load_text(text_path.encode())
    assert h
    desc = np.ascontiguousarray(desc, np.uint8)
    n = len(desc)
    ids = np.zeros(max(n, 1), np.int32); vals = np.zeros(max(n, 1), np.float64)
    fid = np.zeros(max(n, 1), np.int32); fst = np.zeros(max(n, 1) + 1, np.int32); fit = np.zeros(max(n, 1), np.int32)
    nn = C.c_int32()
    k = L_.refv_transform(C.c_void_p(h), _p(desc), n, int(levelsup), _p(ids), _p(vals), _p(fid), _p(fst), _p(fit), C.byref(nn))
    size = L_.refv_size(C.c_void_p(h))
    L_.refv_free(C.c_void_p(h))
    return ids[:k].copy(), vals[:k].copy(), (fid[:nn.value].copy(), fst[:nn.value + 1].copy(), fit[:fst[nn.value]].copy()), size


def distinctive_descriptors(desc, start, _L=None):
    """MapPoint::ComputeDistinctiveDescriptors for CSR lists of observed descriptors -> (best index per point, its median)."""
    desc = np.ascontiguousarray(desc, np.uint8); start = np.ascontiguousarray(start, np.int32)
    npts = len(start) - 1
    best = np.full(max(npts, 1), -1, np.int32); med = np.zeros(max(npts, 1), np.int32)
    if _L is None:
        lib().orc_distinctive_descriptors(_p(desc), _p(start), npts, _p(best), _p(med))
        return best[:npts], med[:npts]
    _L.refm_distinctive_descriptors(_p(desc), _p(start), npts, _p(best))
    return best[:npts], None


def bruteforce_top2(q, t):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    bi, bd, sd = (np.zeros(max(len(q), 1), np.int32) for _ in range(3))
    lib().orc_bruteforce_top2(_p(q), len(q), _p(t), len(t), _p(bi), _p(bd), _p(sd))
    return bi[:len(q)], bd[:len(q)], sd[:len(q)]


def fisheye_undistort(pts, K, D):
    """cv::fisheye::undistortPoints(pts, K, D, R=I, P=K) restated (OpenCV 4.13 semantics); pts float32 [n,2]."""
    pts = np.ascontiguousarray(pts, np.float32)
    K = np.ascontiguousarray(K, np.float32); D = np.ascontiguousarray(D, np.float32)
    out = np.empty_like(pts)
    lib().orc_fisheye_undistort(_p(pts), len(pts), _p(K), _p(D), _p(out))
    return out


def bird_near_edges(contour, xy):
    """Frame::nearEdges (Frame.cc:717-739) for points xy float32 [n,2] -> u8[n]."""
    contour = np.asarray(contour); xy = np.ascontiguousarray(xy, np.float32)
    assert contour.dtype == np.uint8 and contour.strides[1] == 1
    keep = np.zeros(len(xy), np.uint8)
    lib().orc_bird_near_edges(_p(contour), contour.shape[0], contour.shape[1], contour.strides[0], _p(xy), len(xy), _p(keep))
    return keep


def corner_subpix(img, xy, half_win=(5, 5), max_iter=40, eps=0.001):
    """cv::cornerSubPix(img, xy, half_win, (-1,-1), (EPS+MAX_ITER, max_iter, eps)) restated (OpenCV 4.13) -> (xy', iterations)."""
    img = np.asarray(img); out = np.array(xy, np.float32, order="C")
    assert img.dtype == np.uint8 and img.strides[1] == 1
    it = np.zeros(len(out), np.int32)
    lib().orc_corner_subpix(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(out), len(out), int(half_win[0]), int(half_win[1]),
                            int(max_iter), C.c_double(eps), _p(it))
    return out, it


def resize_linear_exact(img, dw, dh):
    """cv::resize(img u8, (dw, dh), INTER_LINEAR_EXACT) restated (the resize of cv::ORB's image / mask pyramid)."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((dh, dw), np.uint8)
    lib().orc_resize_linear_exact_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw)
    return out


def retain_best(response, n_points):
    """KeyPointsFilter::retainBest on a response array -> (order [n] after std::nth_element + std::partition, number kept)."""
    r = np.ascontiguousarray(response, np.float32)
    order = np.zeros(max(len(r), 1), np.int32)
    kept = lib().orc_retain_best(_p(r), len(r), int(n_points), _p(order))
    return order[:len(r)], kept


def antiselect(n, nth):
    """Responses on which this libstdc++'s std::nth_element(.., nth, .., greater) exhausts introselect's depth budget and falls back
    to heap select (McIlroy's adversary run against the real algorithm, oracle/cvorb_oracle.cpp)."""
    out = np.zeros(max(n, 1), np.float32)
    lib().orc_antiselect(int(n), int(nth), _p(out))
    return out[:n]


def cvorb_blur(img):
    """The GaussianBlur cv::ORB applies to a pyramid level (float sepFilter2D path with FMA, see oracle/cvorb_oracle.cpp)."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros_like(img)
    lib().orc_cvorb_blur(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(out))
    return out


def cvorb_detect(img, mask=None, nfeatures=2000):
    """cv::ORB::create(nfeatures)->detect(img, keypoints, mask) restated -> keypoints (KP_DTYPE) in cv::ORB's output order."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.strides[1] == 1
    if mask is not None:
        mask = np.asarray(mask)
        assert mask.dtype == np.uint8 and mask.shape == img.shape and mask.strides[1] == 1
    cap = 4 * nfeatures + 4096
    while True:
        out = np.zeros(cap, KP_DTYPE)
        n = lib().orc_cvorb_detect(_p(img), img.shape[0], img.shape[1], img.strides[0], None if mask is None else _p(mask),
                                   0 if mask is None else mask.strides[0], int(nfeatures), _p(out), cap, None)
        if n <= cap:
            return out[:n].copy()
        cap = n


def cvorb_compute(img, kps):
    """cv::ORB::create(...)->compute(img, keypoints, descriptors) restated -> (surviving keypoints, descriptors [n, 32])."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.strides[1] == 1
    k = np.array(kps, KP_DTYPE, order="C")
    d = np.zeros((max(len(k), 1), 32), np.uint8)
    n = lib().orc_cvorb_compute(_p(img), img.shape[0], img.shape[1], img.strides[0], _p(k), len(k), _p(d))
    assert n >= 0
    return k[:n].copy(), d[:n].copy()


def rect_subpix(img, cx, cy, ww, wh):
    """cv::getRectSubPix(img u8, (ww, wh), (cx, cy), patchType=CV_32F) restated."""
    img = np.asarray(img); out = np.zeros((wh, ww), np.float32)
    lib().orc_rect_subpix(_p(img), img.shape[0], img.shape[1], img.strides[0], C.c_float(cx), C.c_float(cy), ww, wh, _p(out))
    return out


# ---- the reference's OWN matcher (src/ORBmatcher.cc compiled verbatim, oracle/_ref/libfbe_refmatch.so) -----------------
_refm = None


_refm_o3 = None


def refmatch(timing: bool = False):
    """ctypes handle of the verbatim matcher build, or None when oracle/_ref was not built (no /root/reference).
    timing=True: the -O3 -march=x86-64-v3 variant for bench.py's CPU legs."""
    global _refm, _refm_o3
    if timing:
        if _refm_o3 is None:
            path = os.path.join(HERE, "_ref", "libfbe_refmatch_o3.so")
            if not os.path.exists(path) or not host_runs_timing_build():
                return None
            _refm_o3 = C.CDLL(path)
        return _refm_o3
    if _refm is None:
        path = os.path.join(HERE, "_ref", "libfbe_refmatch.so")
        if not os.path.exists(path):
            return None
        _refm = C.CDLL(path)
    return _refm


_dropm = None


def dropinmatch():
    """ctypes handle of the DROP-IN matcher build (the reference's ORBmatcher class with host/ORBmatcher_fbe.cc's bodies on
    libfbe_b200.so, oracle/Makefile target dropinmatch), or None when it was not built.  Needs a GPU to be called."""
    global _dropm
    if _dropm is None:
        path = os.path.join(HERE, "_ref", "libfbe_dropinmatch.so")
        if not os.path.exists(path):
            return None
        _dropm = C.CDLL(path)
    return _dropm


class RefMatch:
    """Same call shapes as the restated functions above, executed by the reference's compiled code (or, with
    lib=dropinmatch(), by the reference's class with the drop-in bodies)."""

    def __init__(self, lib=None):
        self.L = lib if lib is not None else refmatch()
        assert self.L is not None

    def grid_assign(self, kps, min_x, min_y, inv_w, inv_h, gcols, grows):
        kps = np.ascontiguousarray(kps)
        start = np.zeros(gcols * grows + 1, np.int32)
        items = np.zeros(max(len(kps), 1), np.int32)
        n = self.L.refm_grid_assign(_p(kps), len(kps), _fp(min_x), _fp(min_y), _fp(inv_w), _fp(inv_h), gcols, grows, _p(start), _p(items))
        return start, items[:n].copy()

    def features_in_area(self, f, x, y, r, min_level=-1, max_level=-1, upper_inclusive=True):
        v = _view(f)
        out = np.zeros(max(len(f.kps), 1), np.int32)
        n = self.L.refm_features_in_area(C.byref(v), _fp(x), _fp(y), _fp(r), int(min_level), int(max_level), int(upper_inclusive), _p(out), len(out))
        return out[:n].copy()

    def search_for_initialization(self, f1, f2, prev_matched, window, nn_ratio, check_ori):
        v1, v2 = _view(f1), _view(f2)
        m12 = np.full(max(len(f1.kps), 1), -1, np.int32)
        n = self.L.refm_search_for_initialization(C.byref(v1), C.byref(v2), _p(prev_matched), _p(m12), int(window), _fp(nn_ratio), int(check_ori))
        return n, m12[:len(f1.kps)]

    def birdview_match(self, ref_kps, ref_desc, cur, window, nn_ratio, check_ori):
        v = _view(cur)
        ref_kps = np.ascontiguousarray(ref_kps); ref_desc = np.ascontiguousarray(ref_desc, np.uint8)
        dm = np.zeros((max(len(ref_kps), 1), 3), np.int32)
        nd = C.c_int32()
        n = self.L.refm_birdview_match(_p(ref_kps), _p(ref_desc), len(ref_kps), C.byref(v), int(window), _fp(nn_ratio), int(check_ori), _p(dm), C.byref(nd))
        return n, dm[:nd.value].copy()

    def bird_map_point_match(self, mp_base, mp_desc, cur, window, nn_ratio):
        """-> (inliers, pixels the reference searched around [n,2] (NaN = rejected), assigned[cur.N])"""
        v = _view(cur)
        mp_base = np.ascontiguousarray(mp_base, np.float32); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        pix = np.zeros((len(mp_base), 2), np.float32)
        assigned = np.full(max(len(cur.kps), 1), -1, np.int32)
        n = self.L.refm_bird_map_point_match(_p(mp_base), _p(mp_desc), len(mp_base), C.byref(v), int(window), _fp(nn_ratio), _p(pix), _p(assigned))
        return n, pix, assigned[:len(cur.kps)]

    def search_by_projection_last(self, cur, last_kps, last_proj, last_mp_desc, scale_factors, th, check_ori, cur_taken=None, last_has_obs=None):
        v = _view(cur)
        last_kps = np.ascontiguousarray(last_kps); last_proj = np.ascontiguousarray(last_proj, np.float32)
        last_mp_desc = np.ascontiguousarray(last_mp_desc, np.uint8); sf = np.ascontiguousarray(scale_factors, np.float32)
        cur_mp = np.full(max(len(cur.kps), 1), -1, np.int32)
        tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
        ho = None if last_has_obs is None else np.ascontiguousarray(last_has_obs, np.uint8)
        n = self.L.refm_search_by_projection_last(C.byref(v), _p(last_kps), _p(last_proj), _p(last_mp_desc), len(last_kps), _p(sf),
                                                  None if tk is None else _p(tk), None if ho is None else _p(ho), _fp(th), int(check_ori), _p(cur_mp))
        return n, cur_mp[:len(cur.kps)]

    def search_by_projection_map(self, cur, scale_factors, mp_proj, mp_level, mp_viewcos, mp_desc, th, nn_ratio, cur_taken=None, mp_has_obs=None):
        v = _view(cur)
        sf = np.ascontiguousarray(scale_factors, np.float32); mp_proj = np.ascontiguousarray(mp_proj, np.float32)
        mp_level = np.ascontiguousarray(mp_level, np.int32); mp_viewcos = np.ascontiguousarray(mp_viewcos, np.float32)
        mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        cur_mp = np.full(max(len(cur.kps), 1), -1, np.int32)
        tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
        ho = None if mp_has_obs is None else np.ascontiguousarray(mp_has_obs, np.uint8)
        n = self.L.refm_search_by_projection_map(C.byref(v), _p(sf), _p(mp_proj), _p(mp_level), _p(mp_viewcos), _p(mp_desc), len(mp_level),
                                                 None if tk is None else _p(tk), None if ho is None else _p(ho), _fp(th), _fp(nn_ratio), _p(cur_mp))
        return n, cur_mp[:len(cur.kps)]

    def search_by_projection_kf(self, cur, q_kps, proj, level, mp_desc, scale_factors, th, th_dist, level_up, check_ori, cur_taken=None):
        v = _view(cur)
        q_kps = np.ascontiguousarray(q_kps); proj = np.ascontiguousarray(proj, np.float32); level = np.ascontiguousarray(level, np.int32)
        mp_desc = np.ascontiguousarray(mp_desc, np.uint8); sf = np.ascontiguousarray(scale_factors, np.float32)
        cur_mp = np.full(max(len(cur.kps), 1), -1, np.int32)
        tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
        n = self.L.refm_search_by_projection_kf(C.byref(v), _p(q_kps), _p(proj), _p(level), _p(mp_desc), len(level), _p(sf),
                                                None if tk is None else _p(tk), _fp(th), int(th_dist), int(level_up), int(check_ori), _p(cur_mp))
        return n, cur_mp[:len(cur.kps)]

    def search_by_bow(self, kf_kps, kf_desc, kf_has_mp, kf_fv, f_kps, f_desc, f_fv, nn_ratio, check_ori):
        kf_kps = np.ascontiguousarray(kf_kps); kf_desc = np.ascontiguousarray(kf_desc, np.uint8)
        kf_has_mp = np.ascontiguousarray(kf_has_mp, np.uint8)
        f_kps = np.ascontiguousarray(f_kps); f_desc = np.ascontiguousarray(f_desc, np.uint8)
        ka, kb, kc = (np.ascontiguousarray(a, np.int32) for a in kf_fv)
        fa, fb, fc = (np.ascontiguousarray(a, np.int32) for a in f_fv)
        f_mp = np.full(max(len(f_kps), 1), -1, np.int32)
        n = self.L.refm_search_by_bow(_p(kf_kps), _p(kf_desc), len(kf_kps), _p(kf_has_mp), _p(ka), _p(kb), _p(kc), len(ka),
                                      _p(f_kps), _p(f_desc), len(f_kps), _p(fa), _p(fb), _p(fc), len(fa), _fp(nn_ratio), int(check_ori), _p(f_mp))
        return n, f_mp[:len(f_kps)]

    def search_by_bow_kf(self, k1, d1, has1, fv1, k2, d2, has2, fv2, nn_ratio, check_ori):
        return search_by_bow_kf(k1, d1, has1, fv1, k2, d2, has2, fv2, nn_ratio, check_ori, _L=self.L, _fn="refm_search_by_bow_kf")

    def distinctive_descriptors(self, desc, start):
        return distinctive_descriptors(desc, start, _L=self.L)

    def search_for_triangulation(self, *a):
        return search_for_triangulation(*a, _L=self.L)

    def fuse(self, *a):
        return fuse(*a, _L=self.L)

    def search_by_sim3(self, *a):
        return search_by_sim3(*a, _L=self.L)

    def check_models(self, *a):
        return check_models(*a, _L=self.L)

    def hamming256(self, a, b):
        return self.L.refm_hamming256(_p(np.ascontiguousarray(a, np.uint8)), _p(np.ascontiguousarray(b, np.uint8)))
