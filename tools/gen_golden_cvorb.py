"""Known answers for the bird-view cv::ORB row (f-3) from REAL cv2 4.13.0 calls (run in the build container; cv2 does the work):
   cv2.ORB_create(2000).detect(img, mask)                         -- src/Frame.cc:336-338
   cv2.ORB_create(2000).compute(img, refined keypoints)           -- src/Frame.cc:355
plus the two OpenCV primitives this row adds (INTER_LINEAR_EXACT resize, the float sepFilter2D Gaussian cv::ORB ends up with).
Writes tests/golden/cvorb.npz."""
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bird_scenes as S  # noqa: E402
from fishbirdeyevisualslam_b200 import synth  # noqa: E402
from fishbirdeyevisualslam_b200._lib import KP_DTYPE  # noqa: E402

cv2.setNumThreads(1)


def cvkps(kps):
    return np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in kps], dtype=KP_DTYPE)


def tocv(a):
    return [cv2.KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]), int(k["octave"]),
                         int(k["class_id"])) for k in a]


out = {"cv2_version": cv2.__version__, "ncases": 6}
crit = (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 40, 0.001)
for i in range(6):
    img = S.bird_image(i) if i < 5 else np.ascontiguousarray(synth.road_frame(384, 384, 77))
    mask = S.bird_mask(i)
    orb = cv2.ORB_create(2000)
    det = cvkps(orb.detect(img, mask))
    # the reference refines the kept keypoints with cornerSubPix before compute(); here every detection is refined
    pts = np.ascontiguousarray(np.stack([det["x"], det["y"]], 1)).reshape(-1, 1, 2)
    ref = cv2.cornerSubPix(img, pts.copy(), (5, 5), (-1, -1), crit).reshape(-1, 2)
    moved = det.copy()
    moved["x"], moved["y"] = ref[:, 0], ref[:, 1]
    kc, dc = orb.compute(img, tocv(moved))
    kc = cvkps(kc)
    out[f"img_crc{i}"] = synth.crc(img)
    out[f"det_n{i}"] = len(det)
    out[f"det_crc{i}"] = zlib.crc32(det.tobytes()) & 0xFFFFFFFF
    out[f"cmp_n{i}"] = len(kc)
    out[f"cmp_kps_crc{i}"] = zlib.crc32(kc.tobytes()) & 0xFFFFFFFF
    out[f"cmp_desc_crc{i}"] = zlib.crc32(np.ascontiguousarray(dc).tobytes()) & 0xFFFFFFFF
    if i < 2:                                  # two cases in full, so that a mismatch can be localised without cv2
        out[f"det{i}"] = det
        out[f"moved{i}"] = moved
        out[f"cmp_kps{i}"] = kc
        out[f"cmp_desc{i}"] = dc
rng = np.random.default_rng(5)
a = rng.integers(0, 256, (107, 129), dtype=np.uint8)
out["rs_src"] = a
out["rs_dst"] = cv2.resize(a, (108, 89), interpolation=cv2.INTER_LINEAR_EXACT)
kern = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
out["blur_src"] = a
out["blur_dst"] = cv2.sepFilter2D(a, cv2.CV_8U, kern, kern, borderType=cv2.BORDER_REFLECT_101)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cvorb.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") and v.shape else v) for k, v in out.items()})
