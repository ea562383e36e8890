"""Python mirror of the reference's ORBextractor (include/ORBextractor.h:45-111) over the C-ABI.

Same constructor arguments, same call shape -- `keypoints, descriptors = extractor(image, mask)` for
`operator()(image, mask, keypoints, descriptors)` -- and the same getters.  All compute happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, ExtractorCfg, check, ptr


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int,
                 max_batch: int = 1, device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        cfg = ExtractorCfg(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_batch, device)
        check(self._L.fbe_extractor_create(C.byref(cfg), C.byref(self._h)))
        self.nfeatures, self.nlevels, self.max_batch = nfeatures, nlevels, max_batch
        self._scaleFactor = scaleFactor
        n = C.c_int32()
        ps = [C.POINTER(C.c_float)() for _ in range(4)]
        check(self._L.fbe_extractor_tables(self._h, C.byref(n), *[C.byref(p) for p in ps]))
        self._tables = [np.ctypeslib.as_array(p, shape=(n.value,)).copy() for p in ps]
        pl = C.POINTER(C.c_int32)()
        check(self._L.fbe_extractor_features_per_level(self._h, C.byref(pl)))
        self.mnFeaturesPerLevel = np.ctypeslib.as_array(pl, shape=(n.value,)).copy()
        self._cap = {}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.fbe_extractor_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # --- reference getters (include/ORBextractor.h:63-83) -------------------------------------------------
    def GetLevels(self) -> int:
        return self.nlevels

    def GetScaleFactor(self) -> float:
        return self._scaleFactor

    def GetScaleFactors(self):
        return self._tables[0].copy()

    def GetInverseScaleFactors(self):
        return self._tables[1].copy()

    def GetScaleSigmaSquares(self):
        return self._tables[2].copy()

    def GetInverseScaleSigmaSquares(self):
        return self._tables[3].copy()

    def max_keypoints(self, rows: int, cols: int) -> int:
        key = (rows, cols)
        if key not in self._cap:
            c = C.c_int32()
            check(self._L.fbe_extractor_max_keypoints(self._h, rows, cols, C.byref(c)))
            self._cap[key] = c.value
        return self._cap[key]

    # --- operator() ------------------------------------------------------------------------------------------
    def __call__(self, image: np.ndarray, mask=None):
        """Returns (keypoints[KP_DTYPE], descriptors[n,32] u8).  `mask` is ignored, as in the reference."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 image expected (reference asserts the same)"
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        rows, cols = image.shape
        cap = self.max_keypoints(rows, cols)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int32()
        check(self._L.fbe_extract(self._h, ptr(image), rows, cols, image.strides[0], ptr(kps), ptr(desc), cap, C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_with_pyramid(self, image: np.ndarray):
        """operator() as the reference runs it: keypoints, descriptors AND mvImagePyramid of the same image in one call
        (fbe_extract_pyramid: the level copies ride behind detection).  Returns (keypoints, descriptors, [padded level images])."""
        assert image.dtype == np.uint8 and image.ndim == 2
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        rows, cols = image.shape
        cap = self.max_keypoints(rows, cols)
        nl = self.GetLevels()
        lr, lc = (C.c_int32 * nl)(), (C.c_int32 * nl)()
        check(self._L.fbe_pyramid_geometry(self._h, rows, cols, lr, lc))
        levels = [np.zeros((lr[l] + 38, lc[l] + 38), np.uint8) for l in range(nl)]
        dst = (C.c_void_p * nl)(*[a.ctypes.data for a in levels])
        steps = (C.c_size_t * nl)(*[a.strides[0] for a in levels])
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int32()
        check(self._L.fbe_extract_pyramid(self._h, ptr(image), rows, cols, image.strides[0], ptr(kps), ptr(desc), cap, C.byref(n), dst, steps))
        return kps[:n.value].copy(), desc[:n.value].copy(), levels

    def extract_batch(self, images):
        """List/array of equally sized u8 images -> list of (keypoints, descriptors)."""
        imgs = [np.ascontiguousarray(im, np.uint8) for im in images]
        rows, cols = imgs[0].shape
        cap = self.max_keypoints(rows, cols)
        nimg = len(imgs)
        kps = np.zeros((nimg, cap), KP_DTYPE)
        desc = np.zeros((nimg, cap, 32), np.uint8)
        n = np.zeros(nimg, np.int32)
        arr = (C.c_void_p * nimg)(*[im.ctypes.data for im in imgs])
        check(self._L.fbe_extract_batch(self._h, arr, nimg, rows, cols, cols, ptr(kps), ptr(desc), cap, ptr(n)))
        return [(kps[i, :n[i]].copy(), desc[i, :n[i]].copy()) for i in range(nimg)]

    # --- mvImagePyramid + stage taps -------------------------------------------------------------------------
    def pyramid_level(self, level: int, slot: int = 0) -> np.ndarray:
        """Padded level image ((rows+38) x (cols+38)) of the last call; the ROI is out[19:-19, 19:-19]."""
        r, c = C.c_int32(), C.c_int32()
        check(self._L.fbe_pyramid_level(self._h, slot, level, None, 0, C.byref(r), C.byref(c)))
        out = np.zeros((r.value + 38, c.value + 38), np.uint8)
        check(self._L.fbe_pyramid_level(self._h, slot, level, ptr(out), out.strides[0], C.byref(r), C.byref(c)))
        return out

    def debug_blurred(self, level: int, slot: int = 0) -> np.ndarray:
        r, c = C.c_int32(), C.c_int32()
        check(self._L.fbe_debug_blurred(self._h, slot, level, None, C.byref(r), C.byref(c)))
        out = np.zeros((r.value, c.value), np.uint8)
        check(self._L.fbe_debug_blurred(self._h, slot, level, ptr(out), C.byref(r), C.byref(c)))
        return out

    def debug_candidates(self, level: int, slot: int = 0) -> np.ndarray:
        n = C.c_int32()
        check(self._L.fbe_debug_candidates(self._h, slot, level, None, 0, C.byref(n)))
        out = np.zeros((n.value, 3), np.int32)
        if n.value:
            check(self._L.fbe_debug_candidates(self._h, slot, level, ptr(out), n.value, C.byref(n)))
        return out


def debug_octree(xys: np.ndarray, min_x: int, max_x: int, min_y: int, max_y: int, nfeat: int) -> np.ndarray:
    L = _lib.load()
    xys = np.ascontiguousarray(xys, np.int32).reshape(-1, 3)
    cap = max(nfeat, 4 * max(1, round((max_x - min_x) / (max_y - min_y)))) + 16
    sel = np.zeros(cap, np.int32)
    n = C.c_int32()
    check(L.fbe_debug_octree(ptr(xys), len(xys), min_x, max_x, min_y, max_y, nfeat, ptr(sel), cap, C.byref(n)))
    return sel[:n.value].copy()
