#!/usr/bin/env python3
"""tests/golden/undistort.npz: cv2.fisheye.undistortPoints(pts, K, D, R=I, P=K) known answers (cv2 4.13.0, build container
only) for Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:638-669, 741-795) with the shipped calibration
(Examples/Monocular/fisheye.yaml:8-16) and two stronger distortions."""
import os, sys
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rng = np.random.default_rng(638)
d = {"cv2_version": np.array(cv2.__version__)}
cases = [((348.5, 347.0, 480.0, 302.0), (-0.0488316, 0.000298406, -0.00591118, 0.00193258), 950, 400),     # fisheye.yaml
         ((348.5, 347.0, 480.0, 302.0), (0.12, -0.03, 0.004, -0.0005), 950, 400),
         ((600.0, 600.0, 640.0, 360.0), (-0.2, 0.05, -0.01, 0.001), 1280, 720)]
for i, (Kp, D, w, h) in enumerate(cases):
    K = np.array([[Kp[0], 0, Kp[2]], [0, Kp[1], Kp[3]], [0, 0, 1]], np.float32)
    Dm = np.array(D, np.float32).reshape(4, 1)
    pts = np.stack([rng.uniform(0, w, 3000), rng.uniform(0, h, 3000)], 1).astype(np.float32)
    pts[:4] = [[0, 0], [w, 0], [0, h], [w, h]]          # the corners ComputeImageBounds uses
    pts[4] = [Kp[2], Kp[3]]                              # principal point (theta_d == 0 branch)
    pts[5:200] = np.rint(pts[5:200])                     # integral level-0 keypoint positions
    out = cv2.fisheye.undistortPoints(pts.reshape(-1, 1, 2), K, Dm, R=np.eye(3), P=K).reshape(-1, 2)
    d[f"K{i}"] = np.array(Kp, np.float32); d[f"D{i}"] = np.array(D, np.float32); d[f"size{i}"] = np.array([w, h], np.int32)
    d[f"pts{i}"] = pts; d[f"out{i}"] = out.astype(np.float32)
d["ncases"] = np.array(len(cases))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "undistort.npz"), **d)
print("written", {k: v.shape for k, v in d.items() if k.startswith("out")})
