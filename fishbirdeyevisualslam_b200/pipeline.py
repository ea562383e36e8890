"""Device-resident front+bird batch pipeline (configs C2/C4) over the C-ABI: extract both views, bucket, match each pair
against the previous one.  What Tracking does per frame with `Frame` + `ORBmatcher`, batched and kept on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, PAIR_RESULT_DTYPE, ExtractorCfg, PipelineCfg, PipelineFeatures, check, ptr


class FrontBirdPipeline:
    def __init__(self, batch: int, front_shape=(720, 1280), bird_shape=(384, 384), front_features=2000, bird_features=1000,
                 scale=1.2, nlevels=8, ini_th=15, min_th=5, nn_ratio=0.9, check_orientation=True, front_window=100,
                 bird_window=10, device=0, front_fisheye=None, front_row_cap=0):
        """front_fisheye = (K, D) with K = (fx, fy, cx, cy), D = (k1, k2, p1, p2): undistort the front keypoints on the device
        before the grid (Frame::UndistortKeyPoints), as the reference's Frame constructor does for its fisheye camera."""
        self._L = _lib.load()
        self.batch = batch
        self.front_shape, self.bird_shape = tuple(front_shape), tuple(bird_shape)
        cfg = PipelineCfg(ExtractorCfg(front_features, scale, nlevels, ini_th, min_th, batch + 1, device),
                          ExtractorCfg(bird_features, scale, nlevels, ini_th, min_th, batch + 1, device),
                          front_shape[0], front_shape[1], bird_shape[0], bird_shape[1], batch, nn_ratio,
                          int(check_orientation), front_window, bird_window, device)
        cfg.front_row_cap = int(front_row_cap)       # 0 = 256 candidates per front search window
        if front_fisheye is not None:
            cfg.front_fisheye = 1
            cfg.front_K[:] = [float(x) for x in front_fisheye[0]]
            cfg.front_D[:] = [float(x) for x in front_fisheye[1]]
        self._h = C.c_void_p()
        check(self._L.fbe_pipeline_create(C.byref(cfg), C.byref(self._h)))
        fc, bc = C.c_int32(), C.c_int32()
        check(self._L.fbe_pipeline_caps(self._h, C.byref(fc), C.byref(bc)))
        self.front_cap, self.bird_cap = fc.value, bc.value
        s = C.c_void_p()
        check(self._L.fbe_pipeline_stream(self._h, C.byref(s)))
        self.stream_ptr = s.value or 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.fbe_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def step_dev(self, d_front_ptr: int, d_bird_ptr: int):
        """Asynchronous step on device-resident inputs ([batch,rows,cols] u8 each)."""
        check(self._L.fbe_pipeline_step_dev(self._h, C.c_void_p(d_front_ptr), C.c_void_p(d_bird_ptr)))

    def join(self):
        """Order the pipeline's public stream after all submitted steps (stream-level wait, no host synchronisation)."""
        check(self._L.fbe_pipeline_join(self._h))

    def sync(self):
        check(self._L.fbe_pipeline_sync(self._h))

    def step_host(self, h_front_ptr: int, h_bird_ptr: int, res: np.ndarray, fm: np.ndarray | None = None, bm: np.ndarray | None = None):
        check(self._L.fbe_pipeline_step_host(self._h, C.c_void_p(h_front_ptr), C.c_void_p(h_bird_ptr), ptr(res),
                                             None if fm is None else ptr(fm), None if bm is None else ptr(bm)))

    def submit_host(self, h_front_ptr: int, h_bird_ptr: int, res: np.ndarray, fm: np.ndarray | None = None, bm: np.ndarray | None = None,
                    features=None) -> int:
        """Asynchronous step through host (pinned) buffers; returns a ticket for wait().  Up to three steps may be in flight.
        features = (front_kps [B, front_cap] KP_DTYPE, front_desc [B, front_cap, 32] u8, bird_kps, bird_desc) host arrays that
        receive what ORBextractor::operator() returns for every frame of the batch (any entry may be None)."""
        t = C.c_int32()
        if features is None:
            check(self._L.fbe_pipeline_submit_host(self._h, C.c_void_p(h_front_ptr), C.c_void_p(h_bird_ptr), ptr(res),
                                                   None if fm is None else ptr(fm), None if bm is None else ptr(bm), C.byref(t)))
        else:
            f = PipelineFeatures(*[None if a is None else a.ctypes.data for a in features])
            check(self._L.fbe_pipeline_submit_host_features(self._h, C.c_void_p(h_front_ptr), C.c_void_p(h_bird_ptr), ptr(res),
                                                            None if fm is None else ptr(fm), None if bm is None else ptr(bm),
                                                            C.byref(f), C.byref(t)))
        return t.value

    def device_results(self):
        """Device addresses (ints) of the last step's fbe_pair_result[B], front matches12 [B, front_cap] i32 and bird matches12
        [B, bird_cap] i32 -- for a caller that moves them between GPUs itself (shard.gather_matches)."""
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._L.fbe_pipeline_device_results(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def wait(self, ticket: int):
        check(self._L.fbe_pipeline_wait(self._h, C.c_int32(ticket)))

    def fetch(self, with_matches=True):
        res = np.zeros(self.batch, PAIR_RESULT_DTYPE)
        fm = np.zeros((self.batch, self.front_cap), np.int32) if with_matches else None
        bm = np.zeros((self.batch, self.bird_cap), np.int32) if with_matches else None
        check(self._L.fbe_pipeline_fetch(self._h, ptr(res), None if fm is None else ptr(fm), None if bm is None else ptr(bm)))
        return res, fm, bm

    def fetch_pair(self, pair: int):
        fk = np.zeros(self.front_cap, KP_DTYPE); fd = np.zeros((self.front_cap, 32), np.uint8)
        bk = np.zeros(self.bird_cap, KP_DTYPE); bd = np.zeros((self.bird_cap, 32), np.uint8)
        check(self._L.fbe_pipeline_fetch_pair(self._h, pair, ptr(fk), ptr(fd), ptr(bk), ptr(bd)))
        return fk, fd, bk, bd

    def fetch_front_undistorted(self, pair: int):
        """(mvKeysUn of one front frame of the last step, (mnMinX, mnMaxX, mnMinY, mnMaxY))."""
        fk = np.zeros(self.front_cap, KP_DTYPE)
        b = np.zeros(4, np.float32)
        check(self._L.fbe_pipeline_fetch_front_undistorted(self._h, pair, ptr(fk), ptr(b)))
        return fk, tuple(float(x) for x in b)

    def last_step_ms(self) -> float:
        ms = C.c_float()
        check(self._L.fbe_pipeline_last_step_ms(self._h, C.byref(ms)))
        return ms.value


class PinnedBuffer:
    """Page-locked host memory from the library (cudaMallocHost) exposed as a numpy array."""

    def __init__(self, shape, dtype=np.uint8):
        self._L = _lib.load()
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = C.c_void_p()
        check(self._L.fbe_host_alloc(C.byref(self._p), C.c_size_t(self.nbytes)))
        buf = (C.c_uint8 * self.nbytes).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)
        self.ptr = self._p.value

    def close(self):
        if getattr(self, "_p", None) and self._p.value:
            self.array = None
            self._L.fbe_host_free(self._p)
            self._p = C.c_void_p()

    __del__ = close
