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


def ref():
    """The reference's own ORBextractor.cc compiled verbatim, or None when oracle/_ref was not built."""
    global _ref
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

    def __init__(self, nfeatures=1000, scale=1.2, nlevels=8, ini_th=15, min_th=5):
        self.R = ref()
        if self.R is None:
            raise RuntimeError("oracle/_ref/libfbe_ref.so not built")
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
