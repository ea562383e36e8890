"""Python mirror of the reference's ORBmatcher (include/ORBmatcher.h:36-108) and of the slice of Frame it reads,
over the C-ABI.  Method names, argument meaning and return values follow the reference; pointer-valued arguments
(MapPoint*, KeyFrame*) are replaced by the flat arrays the C++ drop-in shim extracts from them.
All searches run in the CUDA library; this module only marshals and performs the reference's host-side cv::Mat
arithmetic (projections, bird pixel conversion, the BirdMapPointMatch distance filter).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, FrameView, check, ptr

FRAME_GRID_ROWS, FRAME_GRID_COLS, FRAME_GRID_BIRD = 48, 64, 32          # include/Frame.h:38-40
METER2PIXEL, PIXEL2METER, REAR_AXLE_TO_CENTER = 25.1, 0.03984, 1.393     # src/Frame.cc:39-44


@dataclass
class Frame:
    """The fields of ORB_SLAM2::Frame the matchers read (src/Frame.cc): keypoints, descriptors, grid geometry."""
    kps: np.ndarray                      # KP_DTYPE[n]   (mvKeysUn / mvKeysBird)
    desc: np.ndarray                     # u8[n,32]      (mDescriptors / mDescriptorsBird)
    min_x: float = 0.0                   # mnMinX
    min_y: float = 0.0                   # mnMinY
    inv_w: float = 0.0                   # mfGridElementWidthInv
    inv_h: float = 0.0                   # mfGridElementHeightInv
    gcols: int = FRAME_GRID_COLS
    grows: int = FRAME_GRID_ROWS
    scale_factors: np.ndarray = field(default_factory=lambda: np.ones(8, np.float32))   # mvScaleFactors

    @staticmethod
    def front(kps, desc, cols, rows, scale_factors=None):
        """k1 == 0 image bounds (src/Frame.cc:741-795): mnMinX = 0, mnMaxX = cols, ..."""
        f = Frame(np.ascontiguousarray(kps), np.ascontiguousarray(desc), 0.0, 0.0,
                  float(np.float32(FRAME_GRID_COLS) / np.float32(cols)), float(np.float32(FRAME_GRID_ROWS) / np.float32(rows)),
                  FRAME_GRID_COLS, FRAME_GRID_ROWS)
        if scale_factors is not None:
            f.scale_factors = np.ascontiguousarray(scale_factors, np.float32)
        return f

    @staticmethod
    def bird(kps, desc, bird_cols, bird_rows):
        return Frame(np.ascontiguousarray(kps), np.ascontiguousarray(desc), 0.0, 0.0,
                     float(np.float32(FRAME_GRID_BIRD) / np.float32(bird_cols)),
                     float(np.float32(FRAME_GRID_BIRD) / np.float32(bird_rows)), FRAME_GRID_BIRD, FRAME_GRID_BIRD)

    @property
    def N(self):
        return len(self.kps)

    def view(self) -> FrameView:
        return FrameView(self.kps.ctypes.data if len(self.kps) else None, self.desc.ctypes.data if len(self.kps) else None,
                         len(self.kps), self.min_x, self.min_y, self.inv_w, self.inv_h, self.gcols, self.grows)

    def AssignFeaturesToGrid(self):
        """Frame::AssignFeaturesToGrid (src/Frame.cc:381-411) -> CSR (cell_start[gcols*grows+1], cell_items)."""
        return grid_assign(self.kps, self.min_x, self.min_y, self.inv_w, self.inv_h, self.gcols, self.grows)


def UndistortKeyPoints(kps: np.ndarray, K, D, device: int = 0) -> np.ndarray:
    """Frame::UndistortKeyPoints (Frame.cc:638-669): keypoints with pt replaced by cv::fisheye::undistortPoints(pt, K, D, P=K)."""
    L = _lib.load()
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    out = np.empty_like(kps)
    K = np.ascontiguousarray(K, np.float32); D = np.ascontiguousarray(D, np.float32)
    check(L.fbe_undistort_keypoints(ptr(kps), len(kps), ptr(K), ptr(D), C.c_int32(device), ptr(out)))
    return out


def BirdGuideRefine(contour, img, kps: np.ndarray, half_win=(5, 5), max_iter: int = 40, eps: float = 0.001, device: int = 0):
    """Frame::GuidenceKeyBirdPts (Frame.cc:671-684, nearEdges :717-739) + cv::cornerSubPix (Frame.cc:345-352) on the device.
    contour / img: 8-bit images of one size (either may be None to skip that step).
    -> (keep u8[n], kept keypoints in input order with refined pt, iterations per kept point)."""
    L = _lib.load()
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    ref = contour if contour is not None else img
    rows, cols = ref.shape
    n = len(kps)
    keep = np.zeros(n, np.uint8); out = np.empty_like(kps); iters = np.zeros(n, np.int32); n_out = C.c_int32(0)
    cp = C.c_void_p(contour.ctypes.data) if contour is not None else None
    ip = C.c_void_p(img.ctypes.data) if img is not None else None
    check(L.fbe_bird_refine(cp, C.c_size_t(contour.strides[0] if contour is not None else 0), ip,
                            C.c_size_t(img.strides[0] if img is not None else 0), C.c_int32(rows), C.c_int32(cols), ptr(kps),
                            C.c_int32(n), C.c_int32(half_win[0]), C.c_int32(half_win[1]), C.c_int32(max_iter), C.c_double(eps),
                            C.c_int32(device), ptr(keep), ptr(out), C.byref(n_out), ptr(iters)))
    return keep, out[:n_out.value].copy(), iters[:n_out.value].copy()


def BirdGuideRefineBatch(contours, imgs, kps: np.ndarray, n, half_win=(5, 5), max_iter: int = 40, eps: float = 0.001, device: int = 0):
    """fbe_bird_refine_batch: the two steps of BirdGuideRefine for B frames in one call.  contours / imgs u8 [B, rows, cols]
    (either may be None), kps KP_DTYPE [B, cap], n int32 [B].  -> (keep u8 [B, cap], out KP_DTYPE [B, cap], n_out [B], iters [B, cap])."""
    L = _lib.load()
    kps = np.ascontiguousarray(kps, KP_DTYPE); n = np.ascontiguousarray(n, np.int32)
    B, cap = kps.shape
    ref = contours if contours is not None else imgs
    assert ref.ndim == 3 and ref.shape[0] == B and ref.strides[2] == 1
    rows, cols = ref.shape[1:]
    keep = np.zeros((B, cap), np.uint8); out = np.zeros_like(kps); iters = np.zeros((B, cap), np.int32); n_out = np.zeros(B, np.int32)
    cp = C.c_void_p(contours.ctypes.data) if contours is not None else None
    ip = C.c_void_p(imgs.ctypes.data) if imgs is not None else None
    cs = contours.strides if contours is not None else (0, 0, 0)
    is_ = imgs.strides if imgs is not None else (0, 0, 0)
    check(L.fbe_bird_refine_batch(cp, C.c_size_t(cs[1]), C.c_size_t(cs[0]), ip, C.c_size_t(is_[1]), C.c_size_t(is_[0]), C.c_int32(rows),
                                  C.c_int32(cols), C.c_int32(B), ptr(kps), ptr(n), C.c_int32(cap), C.c_int32(half_win[0]),
                                  C.c_int32(half_win[1]), C.c_int32(max_iter), C.c_double(eps), C.c_int32(device), ptr(keep), ptr(out),
                                  ptr(n_out), ptr(iters)))
    return keep, out, n_out, iters


def isInFrustum(view, pos, normal, min_dist, max_dist, viewing_cos_limit: float, device: int = 0):
    """Frame::isInFrustum (Frame.cc:435-491) for n map points at once.  `view` is a _lib.FrustumView (or anything with the
    same ctypes layout); pos / normal n x 3, min_dist / max_dist = mfMinDistance / mfMaxDistance.
    -> dict(in_view u8, proj n x 2, proj_xr, level, view_cos): mbTrackInView, mTrackProjX/Y, mTrackProjXR, mnTrackScaleLevel,
    mTrackViewCos per point."""
    L = _lib.load()
    pos = np.ascontiguousarray(pos, np.float32); normal = np.ascontiguousarray(normal, np.float32)
    mn = np.ascontiguousarray(min_dist, np.float32); mx = np.ascontiguousarray(max_dist, np.float32)
    n = len(pos)
    o = dict(in_view=np.zeros(n, np.uint8), proj=np.zeros((n, 2), np.float32), proj_xr=np.zeros(n, np.float32),
             level=np.zeros(n, np.int32), view_cos=np.zeros(n, np.float32))
    check(L.fbe_is_in_frustum(C.byref(view), ptr(pos), ptr(normal), ptr(mn), ptr(mx), C.c_int32(n), C.c_float(float(viewing_cos_limit)),
                              C.c_int32(device), ptr(o["in_view"]), ptr(o["proj"]), ptr(o["proj_xr"]), ptr(o["level"]), ptr(o["view_cos"])))
    return o


def _check_models(kps1, kps2, matches, A, B, sigma, device):
    L = _lib.load()
    k1 = np.ascontiguousarray(kps1, KP_DTYPE); k2 = np.ascontiguousarray(kps2, KP_DTYPE)
    m = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
    A = np.ascontiguousarray(A, np.float32).reshape(-1, 9)
    K, n = len(A), len(m)
    scores = np.zeros(max(K, 1), np.float32); inl = np.zeros(max(K * n, 1), np.uint8)
    if B is not None:
        Bm = np.ascontiguousarray(B, np.float32).reshape(-1, 9)
        check(L.fbe_check_homography(ptr(k1), ptr(k2), ptr(m), n, ptr(A), ptr(Bm), K, C.c_float(float(sigma)), C.c_int32(device), ptr(scores), ptr(inl)))
    else:
        check(L.fbe_check_fundamental(ptr(k1), ptr(k2), ptr(m), n, ptr(A), K, C.c_float(float(sigma)), C.c_int32(device), ptr(scores), ptr(inl)))
    return scores[:K], inl[:K * n].reshape(K, n)


def CheckHomography(kps1, kps2, matches, H21, H12, sigma: float = 1.0, device: int = 0):
    """Initializer::CheckHomography (Initializer.cc:391-474) for K hypotheses (H21, H12: K x 3 x 3) -> (scores[K], inliers[K, n])."""
    return _check_models(kps1, kps2, matches, H21, H12, sigma, device)


def CheckFundamental(kps1, kps2, matches, F21, sigma: float = 1.0, device: int = 0):
    """Initializer::CheckFundamental (Initializer.cc:476-554) for K hypotheses (F21: K x 3 x 3) -> (scores[K], inliers[K, n])."""
    return _check_models(kps1, kps2, matches, F21, None, sigma, device)


def ComputeImageBounds(cols: int, rows: int, K, D, device: int = 0):
    """Frame::ComputeImageBounds (Frame.cc:741-795) -> (mnMinX, mnMaxX, mnMinY, mnMaxY)."""
    L = _lib.load()
    K = np.ascontiguousarray(K, np.float32); D = np.ascontiguousarray(D, np.float32)
    b = np.zeros(4, np.float32)
    check(L.fbe_image_bounds(C.c_int32(cols), C.c_int32(rows), ptr(K), ptr(D), C.c_int32(device), ptr(b)))
    return tuple(float(v) for v in b)


def grid_assign(kps, min_x, min_y, inv_w, inv_h, gcols, grows):
    L = _lib.load()
    kps = np.ascontiguousarray(kps)
    start = np.zeros(gcols * grows + 1, np.int32)
    items = np.zeros(max(len(kps), 1), np.int32)
    n = C.c_int32()
    check(L.fbe_grid_assign(ptr(kps) if len(kps) else None, len(kps), min_x, min_y, inv_w, inv_h, gcols, grows,
                            ptr(start), ptr(items), C.byref(n)))
    return start, items[:n.value].copy()


def BaseXY2BirdPixel(local_xyz: np.ndarray, bird_cols: int, bird_rows: int) -> np.ndarray:
    """Converter::BaseXY2BirdPixel (src/Converter.cc:304-310), fp32, integer cols/2 (quirk Q12)."""
    x = np.float32(bird_cols // 2) - local_xyz[:, 1].astype(np.float32) * np.float32(METER2PIXEL)
    y = np.float32(bird_rows // 2) - (local_xyz[:, 0].astype(np.float32) - np.float32(REAR_AXLE_TO_CENTER)) * np.float32(METER2PIXEL)
    return np.stack([x, y], axis=1).astype(np.float32)


def transform_points(T: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """R*x + t with a row-major 4x4; fp32 products accumulated in fp64 then rounded (cv::gemm on CV_32F)."""
    R = T[:3, :3].astype(np.float64)
    t = T[:3, 3].astype(np.float32)
    return ((pts.astype(np.float64) @ R.T).astype(np.float32) + t).astype(np.float32)


class ORBmatcher:
    TH_LOW = 50
    TH_HIGH = 100
    HISTO_LENGTH = 30

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        check(self._L.fbe_matcher_create(C.c_float(nnratio), int(bool(checkOri)), device, C.byref(self._h)))
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.fbe_matcher_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def cache_stats(self):
        """(hits, misses) of the matcher's device-resident frame cache (frames are recognised by content)."""
        h, m = C.c_uint64(), C.c_uint64()
        check(self._L.fbe_matcher_cache_stats(self._h, C.byref(h), C.byref(m)))
        return h.value, m.value

    @staticmethod
    def DescriptorDistance(a: np.ndarray, b: np.ndarray) -> int:
        """ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1951-1967): host inline, never a device round trip."""
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return int(_lib.load().fbe_hamming256(ptr(a), ptr(b)))

    def SearchForInitialization(self, F1: Frame, F2: Frame, vbPrevMatched: np.ndarray, windowSize: int = 10):
        """-> (nmatches, vnMatches12[int32 N1]); vbPrevMatched (float32 [N1,2]) is updated in place."""
        assert vbPrevMatched.dtype == np.float32 and vbPrevMatched.flags.c_contiguous
        m12 = np.full(F1.N, -1, np.int32)
        n = C.c_int32()
        v1, v2 = F1.view(), F2.view()
        check(self._L.fbe_search_for_initialization(self._h, C.byref(v1), C.byref(v2), ptr(vbPrevMatched), ptr(m12),
                                                    int(windowSize), C.byref(n)))
        return n.value, m12

    def BirdviewMatch(self, CurF: Frame, vRefKeysBird: np.ndarray, DescriptorsBird: np.ndarray, windowSize: int = 10):
        """isProject == 0 path. -> (nmatches, DMatch array int32 [k,3] = (queryIdx, trainIdx, distance))."""
        ref_k = np.ascontiguousarray(vRefKeysBird)
        ref_d = np.ascontiguousarray(DescriptorsBird)
        dm = np.zeros((max(len(ref_k), 1), 3), np.int32)
        nd, n = C.c_int32(), C.c_int32()
        v = CurF.view()
        check(self._L.fbe_birdview_match(self._h, ptr(ref_k) if len(ref_k) else None, ptr(ref_d) if len(ref_k) else None,
                                         len(ref_k), C.byref(v), int(windowSize), ptr(dm), C.byref(nd), C.byref(n)))
        return n.value, dm[:nd.value].copy()

    def BirdMapPointMatch(self, CurF: Frame, mp_world: np.ndarray, mp_desc: np.ndarray, Tbw: np.ndarray, Tcw: np.ndarray,
                          cur_cam_xyz: np.ndarray, bird_cols: int, bird_rows: int, windowSize: int = 10, filterSize: float = 0.05):
        """mp_world: float32 [n,3] (NaN row = NULL MapPointBird*).  -> (InlierMatches, vnMatches12, assigned[cur.N])."""
        n_mp = len(mp_world)
        local = transform_points(Tbw, np.nan_to_num(mp_world))
        pix = BaseXY2BirdPixel(local, bird_cols, bird_rows)
        skip = np.isnan(mp_world[:, 0]) | (np.abs(local[:, 2]) > 0.2) | (pix[:, 0] < 0) | (pix[:, 0] >= bird_cols) | \
            (pix[:, 1] < 0) | (pix[:, 1] >= bird_rows)
        pix[skip, 0] = np.nan
        pix = np.ascontiguousarray(pix, np.float32)
        mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        m12 = np.full(n_mp, -1, np.int32)
        n = C.c_int32()
        v = CurF.view()
        check(self._L.fbe_bird_map_point_match(self._h, ptr(pix), ptr(mp_desc), n_mp, C.byref(v), int(windowSize), ptr(m12), C.byref(n)))
        return (*bird_map_second_pass(m12, mp_world, Tcw, cur_cam_xyz, CurF.N, filterSize), m12)

    def SearchByProjectionLast(self, CurrentFrame: Frame, last_kps: np.ndarray, last_proj: np.ndarray, last_mp_desc: np.ndarray,
                               th: float, cur_taken=None, last_has_obs=None):
        """SearchByProjection(CurrentFrame, LastFrame, th, bMono=true) with host-projected (u,v) (NaN u = skipped).
        -> (nmatches, cur_mp[int32 N] = last-frame index assigned to each keypoint or -1)."""
        last_kps = np.ascontiguousarray(last_kps)
        last_proj = np.ascontiguousarray(last_proj, np.float32)
        last_mp_desc = np.ascontiguousarray(last_mp_desc, np.uint8)
        cur_mp = np.full(max(CurrentFrame.N, 1), -1, np.int32)
        n = C.c_int32()
        v = CurrentFrame.view()
        sf = np.ascontiguousarray(CurrentFrame.scale_factors, np.float32)
        tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
        ho = None if last_has_obs is None else np.ascontiguousarray(last_has_obs, np.uint8)
        check(self._L.fbe_search_by_projection_last(self._h, C.byref(v), ptr(last_kps), ptr(last_proj), ptr(last_mp_desc), len(last_kps),
                                                    ptr(sf), len(sf), None if tk is None else ptr(tk), None if ho is None else ptr(ho),
                                                    C.c_float(th), ptr(cur_mp), C.byref(n)))
        return n.value, cur_mp[:CurrentFrame.N]

    def SearchByProjectionMap(self, F: Frame, mp_proj, mp_level, mp_viewcos, mp_desc, th: float = 1.0, cur_taken=None, mp_has_obs=None):
        """SearchByProjection(F, vpMapPoints, th) on the fields Frame::isInFrustum fills."""
        mp_proj = np.ascontiguousarray(mp_proj, np.float32)
        mp_level = np.ascontiguousarray(mp_level, np.int32)
        mp_viewcos = np.ascontiguousarray(mp_viewcos, np.float32)
        mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        cur_mp = np.full(max(F.N, 1), -1, np.int32)
        n = C.c_int32()
        v = F.view()
        sf = np.ascontiguousarray(F.scale_factors, np.float32)
        tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
        ho = None if mp_has_obs is None else np.ascontiguousarray(mp_has_obs, np.uint8)
        check(self._L.fbe_search_by_projection_map(self._h, C.byref(v), ptr(sf), len(sf), ptr(mp_proj), ptr(mp_level), ptr(mp_viewcos),
                                                   ptr(mp_desc), len(mp_level), None if tk is None else ptr(tk),
                                                   None if ho is None else ptr(ho), C.c_float(th), ptr(cur_mp), C.byref(n)))
        return n.value, cur_mp[:F.N]

    def SearchByProjectionReloc(self, CurrentFrame: Frame, kf_kps, mp_proj, mp_level, mp_desc, th: float, orb_dist: int, cur_taken=None):
        """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (relocalisation, ORBmatcher.cc:1473-1600) on
        host-projected map points (NaN u = rejected) with their predicted levels."""
        kf_kps = np.ascontiguousarray(kf_kps)
        mp_proj = np.ascontiguousarray(mp_proj, np.float32)
        mp_level = np.ascontiguousarray(mp_level, np.int32)
        mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        cur_mp = np.full(max(CurrentFrame.N, 1), -1, np.int32)
        n = C.c_int32()
        v = CurrentFrame.view()
        sf = np.ascontiguousarray(CurrentFrame.scale_factors, np.float32)
        tk = None if cur_taken is None else np.ascontiguousarray(cur_taken, np.uint8)
        check(self._L.fbe_search_by_projection_reloc(self._h, C.byref(v), ptr(kf_kps), ptr(mp_proj), ptr(mp_level), ptr(mp_desc),
                                                     len(mp_level), ptr(sf), len(sf), None if tk is None else ptr(tk), C.c_float(th),
                                                     C.c_int32(orb_dist), ptr(cur_mp), C.byref(n)))
        return n.value, cur_mp[:CurrentFrame.N]

    def SearchByProjectionLoop(self, KF: Frame, mp_proj, mp_level, mp_desc, th: int, kf_matched=None):
        """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (loop closing, ORBmatcher.cc:291-404)."""
        mp_proj = np.ascontiguousarray(mp_proj, np.float32)
        mp_level = np.ascontiguousarray(mp_level, np.int32)
        mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        kf_mp = np.full(max(KF.N, 1), -1, np.int32)
        n = C.c_int32()
        v = KF.view()
        sf = np.ascontiguousarray(KF.scale_factors, np.float32)
        tk = None if kf_matched is None else np.ascontiguousarray(kf_matched, np.uint8)
        check(self._L.fbe_search_by_projection_loop(self._h, C.byref(v), ptr(mp_proj), ptr(mp_level), ptr(mp_desc), len(mp_level),
                                                    ptr(sf), len(sf), None if tk is None else ptr(tk), C.c_int32(th), ptr(kf_mp), C.byref(n)))
        return n.value, kf_mp[:KF.N]

    def SearchByBoW(self, kf_kps, kf_desc, kf_has_mp, kf_featvec, F: Frame, f_featvec):
        """SearchByBoW(KeyFrame*, Frame&, matches).  Feature vectors: (node_ids[nn], start[nn+1], items).
        -> (nmatches, f_mp[int32 N] = key-frame keypoint index per frame keypoint or -1)."""
        kf_kps = np.ascontiguousarray(kf_kps)
        kf_desc = np.ascontiguousarray(kf_desc, np.uint8)
        kf_has_mp = np.ascontiguousarray(kf_has_mp, np.uint8)
        ka, kb, kc = (np.ascontiguousarray(a, np.int32) for a in kf_featvec)
        fa, fb, fc = (np.ascontiguousarray(a, np.int32) for a in f_featvec)
        f_mp = np.full(max(F.N, 1), -1, np.int32)
        n = C.c_int32()
        check(self._L.fbe_search_by_bow(self._h, ptr(kf_kps), ptr(kf_desc), len(kf_kps), ptr(kf_has_mp), ptr(ka), ptr(kb), ptr(kc), len(ka),
                                        ptr(F.kps), ptr(F.desc), F.N, ptr(fa), ptr(fb), ptr(fc), len(fa), ptr(f_mp), C.byref(n)))
        return n.value, f_mp[:F.N]

    def SearchByBoWKF(self, kf1_kps, kf1_desc, kf1_has_mp, kf1_featvec, kf2_kps, kf2_desc, kf2_has_mp, kf2_featvec):
        """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) (loop closing, ORBmatcher.cc:523-656).
        -> (nmatches, matches12[int32 n1] = key-frame-2 feature index per key-frame-1 feature or -1)."""
        k1 = np.ascontiguousarray(kf1_kps); d1 = np.ascontiguousarray(kf1_desc, np.uint8); h1 = np.ascontiguousarray(kf1_has_mp, np.uint8)
        k2 = np.ascontiguousarray(kf2_kps); d2 = np.ascontiguousarray(kf2_desc, np.uint8); h2 = np.ascontiguousarray(kf2_has_mp, np.uint8)
        a1, b1, c1 = (np.ascontiguousarray(a, np.int32) for a in kf1_featvec)
        a2, b2, c2 = (np.ascontiguousarray(a, np.int32) for a in kf2_featvec)
        m12 = np.full(max(len(k1), 1), -1, np.int32)
        n = C.c_int32()
        check(self._L.fbe_search_by_bow_kf(self._h, ptr(k1), ptr(d1), len(k1), ptr(h1), ptr(a1), ptr(b1), ptr(c1), len(a1),
                                           ptr(k2), ptr(d2), len(k2), ptr(h2), ptr(a2), ptr(b2), ptr(c2), len(a2), ptr(m12), C.byref(n)))
        return n.value, m12[:len(k1)]

    def SearchForTriangulation(self, kf1_kps, kf1_desc, kf1_has_mp, kf1_stereo, kf1_featvec, kf2_kps, kf2_desc, kf2_has_mp, kf2_stereo,
                               kf2_featvec, F12, epipole, kf2_scale_factors, kf2_level_sigma2, bOnlyStereo=False):
        """SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, bOnlyStereo) (ORBmatcher.cc:658-824).  kfN_stereo = mvuRight >= 0;
        epipole = (ex, ey) of :666-672.  -> (nmatches, matches12[n1]); vMatchedPairs = [(i, m) for i, m in enumerate(matches12) if m >= 0]."""
        k1 = np.ascontiguousarray(kf1_kps); d1 = np.ascontiguousarray(kf1_desc, np.uint8)
        k2 = np.ascontiguousarray(kf2_kps); d2 = np.ascontiguousarray(kf2_desc, np.uint8)
        s1 = np.ascontiguousarray(kf1_stereo, np.uint8); s2 = np.ascontiguousarray(kf2_stereo, np.uint8)
        skip1 = np.ascontiguousarray((np.asarray(kf1_has_mp) != 0) | ((s1 == 0) if bOnlyStereo else False), np.uint8)
        skip2 = np.ascontiguousarray((np.asarray(kf2_has_mp) != 0) | ((s2 == 0) if bOnlyStereo else False), np.uint8)
        a1, b1, c1 = (np.ascontiguousarray(a, np.int32) for a in kf1_featvec)
        a2, b2, c2 = (np.ascontiguousarray(a, np.int32) for a in kf2_featvec)
        F = np.ascontiguousarray(F12, np.float32).ravel()
        sc = np.ascontiguousarray(kf2_scale_factors, np.float32); sg = np.ascontiguousarray(kf2_level_sigma2, np.float32)
        m12 = np.full(max(len(k1), 1), -1, np.int32)
        n = C.c_int32()
        check(self._L.fbe_search_for_triangulation(self._h, ptr(k1), ptr(d1), len(k1), ptr(skip1), ptr(s1), ptr(a1), ptr(b1), ptr(c1), len(a1),
                                                   ptr(k2), ptr(d2), len(k2), ptr(skip2), ptr(s2), ptr(a2), ptr(b2), ptr(c2), len(a2), ptr(F),
                                                   C.c_float(float(epipole[0])), C.c_float(float(epipole[1])), ptr(sc), ptr(sg), len(sc),
                                                   ptr(m12), C.byref(n)))
        return n.value, m12[:len(k1)]

    def FuseSearch(self, KF: Frame, kf_uright, inv_level_sigma2, proj, proj_ur, level, radius, mp_desc, check_chi2: bool):
        """The candidate search of both ORBmatcher::Fuse overloads (ORBmatcher.cc:826-976 with the reprojection gate,
        :978-1101 without).  -> (best_idx[n], best_dist[n]); the caller applies `bestDist <= TH_LOW` and replays the map updates."""
        proj = np.ascontiguousarray(proj, np.float32); level = np.ascontiguousarray(level, np.int32)
        radius = np.ascontiguousarray(radius, np.float32); mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
        inv = np.ascontiguousarray(inv_level_sigma2, np.float32)
        ur = None if kf_uright is None else np.ascontiguousarray(kf_uright, np.float32)
        pur = None if proj_ur is None else np.ascontiguousarray(proj_ur, np.float32)
        n = len(level)
        bi = np.full(max(n, 1), -1, np.int32); bd = np.zeros(max(n, 1), np.int32)
        v = KF.view()
        check(self._L.fbe_fuse_search(self._h, C.byref(v), None if ur is None else ptr(ur), ptr(inv), len(inv), ptr(proj),
                                      None if pur is None else ptr(pur), ptr(level), ptr(radius), ptr(mp_desc), n, C.c_int32(bool(check_chi2)),
                                      ptr(bi), ptr(bd)))
        return bi[:n], bd[:n]

    def SearchBySim3(self, KF1: Frame, KF2: Frame, proj12, level12, desc1, proj21, level21, desc2, pre12, th: float):
        """SearchBySim3 (ORBmatcher.cc:1103-1327) after the host-side projections: proj12 / level12 = key-frame-1 map points in
        key frame 2 (NaN u = skipped), proj21 / level21 the other direction; pre12 = vpMatches12 on entry as key-frame-2
        indices (-1 = NULL).  Two device searches (the Fuse candidate search without the reprojection gate, bound TH_HIGH)
        and the agreement check, exactly what host/ORBmatcher_fbe.cc does.  -> (nFound, matches12[n1])."""
        sf = np.ascontiguousarray(KF1.scale_factors, np.float32)
        level12 = np.ascontiguousarray(level12, np.int32); level21 = np.ascontiguousarray(level21, np.int32)
        bi1, bd1 = self.FuseSearch(KF2, None, sf, proj12, None, level12, (np.float32(th) * sf[level12]).astype(np.float32), desc1, False)
        bi2, bd2 = self.FuseSearch(KF1, None, sf, proj21, None, level21, (np.float32(th) * sf[level21]).astype(np.float32), desc2, False)
        m1 = np.where(bd1 <= 100, bi1, -1)
        m2 = np.where(bd2 <= 100, bi2, -1)
        out = np.ascontiguousarray(pre12, np.int32).copy()
        i1 = np.nonzero(m1 >= 0)[0]
        ok = i1[m2[m1[i1]] == i1]
        out[ok] = m1[ok]
        return len(ok), out

    def ComputeDistinctiveDescriptors(self, desc, start):
        """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:242-307) for many map points: `desc` rows start[p]..start[p+1]-1
        are the descriptors observed for point p.  -> (best index inside each list or -1, that row's median distance)."""
        desc = np.ascontiguousarray(desc, np.uint8); start = np.ascontiguousarray(start, np.int32)
        npts = len(start) - 1
        best = np.full(max(npts, 1), -1, np.int32); med = np.zeros(max(npts, 1), np.int32)
        check(self._L.fbe_distinctive_descriptors(self._h, ptr(desc), ptr(start), npts, ptr(best), ptr(med)))
        return best[:npts], med[:npts]

    def BruteForceTop2(self, q_desc, t_desc):
        q_desc = np.ascontiguousarray(q_desc, np.uint8)
        t_desc = np.ascontiguousarray(t_desc, np.uint8)
        nq = len(q_desc)
        bi, bd, sd = (np.zeros(max(nq, 1), np.int32) for _ in range(3))
        check(self._L.fbe_bruteforce_top2(self._h, ptr(q_desc), nq, ptr(t_desc), len(t_desc), ptr(bi), ptr(bd), ptr(sd)))
        return bi[:nq], bd[:nq], sd[:nq]


def bird_map_second_pass(m12, mp_world, Tcw, cur_cam_xyz, n_cur, filter_size):
    """BirdMapPointMatch second pass (src/ORBmatcher.cc:1865-1895): host arithmetic, as in the reference.
    -> (InlierMatches, assigned[n_cur] = map point index written to mvpMapPointsBird[k] or -1)."""
    assigned = np.full(n_cur, -1, np.int32)
    inliers = 0
    idx = np.nonzero(m12 > 0)[0]                      # `> 0`: a match to keypoint 0 is dropped (quirk Q8)
    if len(idx):
        pc = transform_points(Tcw, mp_world[idx])
        d = np.linalg.norm((pc - cur_cam_xyz[m12[idx]].astype(np.float32)).astype(np.float64), axis=1)
        for i1, dis in zip(idx, d):
            if dis < filter_size:
                assigned[m12[i1]] = i1                # last writer in i1 order wins
                inliers += 1
    return inliers, assigned
