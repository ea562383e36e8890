"""Timing of fbe_bird_refine (host buffers, wall clock incl. H2D/D2H and allocation) beside the CPU oracle and cv2, same inputs."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bird_scenes as S
from fishbirdeyevisualslam_b200.matcher import BirdGuideRefine
from oracle import oracle as O

img, contour = S.bird_image(0), S.contour_image(0)
xy = S.corner_points(img, 0, 2000)
k = S.as_kps(xy)
for _ in range(3):
    BirdGuideRefine(contour, img, k)
t = time.perf_counter()
R = 50
for _ in range(R):
    keep, out, it = BirdGuideRefine(contour, img, k)
gpu_ms = (time.perf_counter() - t) / R * 1e3
t = time.perf_counter()
for _ in range(5):
    ok = O.bird_near_edges(contour, xy); o, oit = O.corner_subpix(img, xy[ok > 0])
cpu_ms = (time.perf_counter() - t) / 5 * 1e3
line = {"n": len(xy), "kept": int(keep.sum()), "mean_iters": float(it.mean()), "gpu_host_api_ms": gpu_ms, "cpu_oracle_ms": cpu_ms,
        "equal": bool(np.stack([out["x"], out["y"]], 1).tobytes() == o.tobytes())}
try:
    import cv2
    cv2.setNumThreads(1)
    crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.001)
    t = time.perf_counter()
    for _ in range(5):
        cv2.cornerSubPix(img, xy[ok > 0].copy().reshape(-1, 1, 2), (5, 5), (-1, -1), crit)
    line["cv2_subpix_ms_1thread"] = (time.perf_counter() - t) / 5 * 1e3
except ImportError:
    pass
# a batch of frames in one call (config C4 style): 64 frames x 2 000 candidates
from fishbirdeyevisualslam_b200.matcher import BirdGuideRefineBatch
from fishbirdeyevisualslam_b200._lib import KP_DTYPE
B = 64
imgs = np.stack([S.bird_image(b % 8) for b in range(B)]); conts = np.stack([S.contour_image(b % 8) for b in range(B)])
kb = np.zeros((B, 2000), KP_DTYPE)
for b in range(B):
    kb[b] = S.as_kps(S.corner_points(imgs[b], b % 8, 2000))
nb = np.full(B, 2000, np.int32)
for _ in range(3):
    BirdGuideRefineBatch(conts, imgs, kb, nb)
t = time.perf_counter()
for _ in range(10):
    _, ob, nob, itb = BirdGuideRefineBatch(conts, imgs, kb, nb)
bms = (time.perf_counter() - t) / 10 * 1e3
line["batch"] = {"frames": B, "kept_total": int(nob.sum()), "gpu_host_api_ms": bms, "ms_per_frame": bms / B}
import json
print(json.dumps(line))
