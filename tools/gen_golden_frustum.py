#!/usr/bin/env python3
"""tests/golden/frustum.npz: Frame::isInFrustum (src/Frame.cc:435-491) evaluated with REAL OpenCV 4.13 calls (cv2.gemm for
`mRcw*P+mtcw` and for UpdatePoseMatrices' `-mRcw.t()*mtcw`, cv2.subtract, cv2.norm) for every cv::Mat expression of the
reference, numpy float32 scalars for its float expressions and glibc logf (via ctypes) for MapPoint::PredictScale.
Build container only (needs cv2); the oracle restatement is checked against the committed file everywhere."""
import ctypes, os, sys
import numpy as np
import cv2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from frustum_scenes import scene

f32 = np.float32
libm = ctypes.CDLL("libm.so.6")
libm.logf.restype = ctypes.c_float; libm.logf.argtypes = [ctypes.c_float]
libm.ceilf.restype = ctypes.c_float; libm.ceilf.argtypes = [ctypes.c_float]


def cv_is_in_frustum(sc):
    Rcw, tcw = sc["Rcw"], sc["tcw"].reshape(3, 1)
    Ow = sc["Ow"].reshape(3, 1)
    fx, fy, cx, cy = (f32(x) for x in sc["K"])
    mnx, mxx, mny, mxy = (f32(x) for x in sc["bounds"])
    n = len(sc["pos"])
    o = dict(in_view=np.zeros(n, np.uint8), proj=np.zeros((n, 2), f32), proj_xr=np.zeros(n, f32), level=np.zeros(n, np.int32), view_cos=np.zeros(n, f32))
    with np.errstate(all="ignore"):
        for i in range(n):
            P = sc["pos"][i].reshape(3, 1)
            Pc = cv2.gemm(Rcw, P, 1.0, tcw, 1.0)
            PcX, PcY, PcZ = f32(Pc[0, 0]), f32(Pc[1, 0]), f32(Pc[2, 0])
            if PcZ < f32(0):
                continue
            invz = f32(1) / PcZ
            u = f32(f32(fx * PcX) * invz) + cx
            v = f32(f32(fy * PcY) * invz) + cy
            if u < mnx or u > mxx or v < mny or v > mxy:
                continue
            maxD, minD = f32(1.2) * sc["max_dist"][i], f32(0.8) * sc["min_dist"][i]
            PO = cv2.subtract(P, Ow)
            dist = f32(cv2.norm(PO))
            if dist < minD or dist > maxD:
                continue
            Pn = sc["normal"][i]
            dot = np.float64(0)
            for k in range(3):
                dot += np.float64(PO[k, 0]) * np.float64(Pn[k])
            view_cos = f32(dot / np.float64(dist))
            if view_cos < f32(sc["cos_limit"]):
                continue
            ratio = sc["max_dist"][i] / dist
            q = f32(libm.logf(float(ratio))) / f32(sc["log_scale"])
            lv = int(libm.ceilf(float(q)))
            lv = 0 if lv < 0 else min(lv, int(sc["n_levels"]) - 1)
            o["in_view"][i] = 1; o["proj"][i] = (u, v); o["proj_xr"][i] = u - f32(sc["mbf"]) * invz; o["level"][i] = lv; o["view_cos"][i] = view_cos
    return o


if __name__ == "__main__":
    d = {}
    for seed in range(3):
        sc = scene(seed, cv2_pose=True)
        for k, v in cv_is_in_frustum(sc).items():
            d[f"s{seed}_{k}"] = v
        d[f"s{seed}_Ow"] = sc["Ow"]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frustum.npz"), **d)
    print("written", {k: (int(v.sum()) if "in_view" in k else v.shape) for k, v in d.items() if "in_view" in k})
