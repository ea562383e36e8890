"""Known answers for the bird-view refinement row from REAL cv2 4.13.0 calls (run in the build container; cv2 does the work):
cv2.cornerSubPix(img, pts, (5,5), (-1,-1), (EPS+MAX_ITER, 40, 0.001)) -- the call of src/Frame.cc:349-350 -- and a few
cv2.getRectSubPix(..., CV_32F) patches.  Writes tests/golden/bird_refine.npz."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bird_scenes as S  # noqa: E402
from fishbirdeyevisualslam_b200 import synth  # noqa: E402

cv2.setNumThreads(1)
out = {"cv2_version": cv2.__version__, "ncases": 3}
crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.001)
for i in range(3):
    img = S.bird_image(i)
    pts = S.corner_points(img, i)
    orb = cv2.ORB_create(2000).detect(img, None)                 # the reference's own detector (Frame.cc:336-338)
    pts = np.ascontiguousarray(np.concatenate([pts, np.float32([k.pt for k in orb])[:800]]))
    ref = cv2.cornerSubPix(img, pts.copy().reshape(-1, 1, 2), (5, 5), (-1, -1), crit).reshape(-1, 2)
    out[f"crc{i}"] = synth.crc(img)
    out[f"pts{i}"] = pts
    out[f"out{i}"] = ref
    rng = np.random.default_rng(100 + i)
    c = np.stack([rng.uniform(10, 370, 8), rng.uniform(10, 370, 8)], 1).astype(np.float32)
    out[f"rc{i}"] = c
    out[f"rp{i}"] = np.stack([cv2.getRectSubPix(img, (13, 13), (float(x), float(y)), patchType=cv2.CV_32F) for x, y in c])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bird_refine.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
