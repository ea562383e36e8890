"""The bird-view feature path the reference ships (src/Frame.cc:336-355) over the C-ABI: cv::ORB::create(2000) detect / compute in
OpenCV 4.13's arithmetic on the GPU, and the whole block (detect -> GuidenceKeyBirdPts -> cornerSubPix -> compute) as one
device-resident call.  Mirrors the cv2 calls a user of the reference would write:

    orb = BirdORB(2000, rows, cols)            # cv::ORB::create(2000)
    kps = orb.detect(img, mask)                # extractorBird->detect(mBirdviewImg, preKeysBird, mBirdviewMask)
    kps, desc = orb.compute(img, kps)          # extractorBird->compute(mBirdviewImg, mvKeysBird, mDescriptorsBird)
    kps, desc, ndet = orb.features(img, mask, contour)     # the whole block
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, ptr


class BirdORB:
    def __init__(self, nfeatures: int = 2000, rows: int = 384, cols: int = 384, max_batch: int = 1, device: int = 0):
        self._L = _lib.load()
        self.rows, self.cols, self.max_batch = rows, cols, max_batch
        self._h = C.c_void_p()
        check(self._L.fbe_bird_orb_create(int(nfeatures), int(rows), int(cols), int(max_batch), int(device), C.byref(self._h)))
        cap = C.c_int32()
        check(self._L.fbe_bird_orb_max_keypoints(self._h, C.byref(cap)))
        self.cap = cap.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.fbe_bird_orb_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _images(self, a, name):
        a = np.asarray(a)
        if a.ndim == 2:
            a = a[None]
        assert a.dtype == np.uint8 and a.shape[1:] == (self.rows, self.cols) and len(a) <= self.max_batch, name
        return np.ascontiguousarray(a)

    def detect_batch(self, imgs, masks=None):
        """-> list of keypoint arrays (KP_DTYPE), one per frame, in cv::ORB's output order."""
        imgs = self._images(imgs, "imgs")
        B = len(imgs)
        masks = None if masks is None else self._images(masks, "masks")
        kps = np.zeros((B, self.cap), KP_DTYPE)
        n = np.zeros(B, np.int32)
        sz = C.c_size_t
        check(self._L.fbe_bird_orb_detect(self._h, ptr(imgs), sz(self.cols), sz(self.rows * self.cols), None if masks is None else ptr(masks),
                                          sz(self.cols), sz(self.rows * self.cols), B, ptr(kps), ptr(n)))
        return [kps[b, :n[b]].copy() for b in range(B)]

    def detect(self, img, mask=None):
        return self.detect_batch(img, mask)[0]

    def compute_batch(self, imgs, kps_list):
        """-> list of (surviving keypoints, descriptors [n, 32]) per frame."""
        imgs = self._images(imgs, "imgs")
        B = len(imgs)
        assert len(kps_list) == B
        kps = np.zeros((B, self.cap), KP_DTYPE)
        n = np.zeros(B, np.int32)
        for b, k in enumerate(kps_list):
            assert len(k) <= self.cap
            kps[b, :len(k)] = k
            n[b] = len(k)
        desc = np.zeros((B, self.cap, 32), np.uint8)
        sz = C.c_size_t
        check(self._L.fbe_bird_orb_compute(self._h, ptr(imgs), sz(self.cols), sz(self.rows * self.cols), B, ptr(kps), ptr(n), ptr(desc)))
        return [(kps[b, :n[b]].copy(), desc[b, :n[b]].copy()) for b in range(B)]

    def compute(self, img, kps):
        return self.compute_batch(img, [kps])[0]

    def features_batch(self, imgs, masks=None, contours=None):
        """The reference's bird block per frame -> list of (mvKeysBird, mDescriptorsBird, |preKeysBird|)."""
        imgs = self._images(imgs, "imgs")
        B = len(imgs)
        masks = None if masks is None else self._images(masks, "masks")
        contours = None if contours is None else self._images(contours, "contours")
        kps = np.zeros((B, self.cap), KP_DTYPE)
        n = np.zeros(B, np.int32)
        ndet = np.zeros(B, np.int32)
        desc = np.zeros((B, self.cap, 32), np.uint8)
        sz = C.c_size_t
        st, sd = sz(self.cols), sz(self.rows * self.cols)
        check(self._L.fbe_bird_features(self._h, ptr(imgs), st, sd, None if masks is None else ptr(masks), st, sd,
                                        None if contours is None else ptr(contours), st, sd, B, ptr(kps), ptr(n), ptr(desc), ptr(ndet)))
        return [(kps[b, :n[b]].copy(), desc[b, :n[b]].copy(), int(ndet[b])) for b in range(B)]

    def features(self, img, mask=None, contour=None):
        return self.features_batch(img, mask, contour)[0]


def retain_best(response: np.ndarray, n_points: int, device: int = 0, with_flag: bool = False):
    """Parity-test tap: KeyPointsFilter::retainBest replayed on the device -> (order [n], number kept[, heap select used])."""
    L = _lib.load()
    r = np.ascontiguousarray(response, np.float32)
    order = np.zeros(max(len(r), 1), np.int32)
    kept = C.c_int32()
    heap = C.c_int32()
    check(L.fbe_debug_retain_best(ptr(r), len(r), int(n_points), int(device), ptr(order), C.byref(kept), C.byref(heap)))
    if with_flag:
        return order[:len(r)], kept.value, bool(heap.value)
    return order[:len(r)], kept.value
