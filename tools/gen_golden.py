#!/usr/bin/env python3
"""Generates tests/golden/*.npz in the BUILD container (needs cv2 4.13.0 and oracle/_ref/libfbe_ref.so):
  prims.npz    -- OpenCV primitive known-answer vectors produced by cv2 itself (resize, border, blur, FAST, fastAtan2)
  extract.npz  -- keypoints + descriptors produced by the reference's OWN ORBextractor.cc (verbatim build) on seeded
                  synthetic images, stored in full for small cases and as CRC32 for the BASELINE-size cases
The fixtures pin oracle/prim.hpp and oracle/orb_oracle.cpp on machines that have neither cv2 nor /root/reference.
"""
import os, sys, zlib
import numpy as np
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fishbirdeyevisualslam_b200 import synth
from oracle import oracle as O

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
cv2.setNumThreads(1)
rng = np.random.default_rng(2026)
d = {"cv2_version": np.array(cv2.__version__)}
# resize
sizes = [(64, 48, 53, 40), (131, 97, 109, 81), (50, 37, 42, 31), (200, 120, 167, 100)]
for i, (sw, sh, dw, dh) in enumerate(sizes):
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    d[f"resize_src{i}"] = src
    d[f"resize_dst{i}"] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
# border
src = rng.integers(0, 256, (23, 31), dtype=np.uint8)
d["border_src"] = src
d["border_dst"] = cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
# blur
for i, (w, h) in enumerate([(61, 47), (33, 9), (7, 7)]):
    src = rng.integers(0, 256, (h, w), dtype=np.uint8)
    d[f"blur_src{i}"] = src
    d[f"blur_dst{i}"] = cv2.GaussianBlur(src, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
# FAST
for i, (h, w, th) in enumerate([(37, 38, 15), (37, 38, 5), (60, 80, 20), (45, 41, 0)]):
    im = synth.frame(h, w, 900 + i)
    det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    k = det.detect(im)
    d[f"fast_src{i}"] = im
    d[f"fast_th{i}"] = np.array(th)
    d[f"fast_out{i}"] = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k], np.int32).reshape(-1, 3)
# fastAtan2
yx = rng.integers(-70000, 70000, (512, 2)).astype(np.float32)
yx[0] = (0, 0); yx[1] = (0, 5); yx[2] = (5, 0); yx[3] = (-5, 0); yx[4] = (0, -5); yx[5] = (3, 3)
d["atan_yx"] = yx
d["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
# cvRound
v = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 2.4999, 2.5001, 1e6 + 0.5], np.float32)
d["round_in"] = v
d["round_out"] = np.array([int(np.rint(np.float32(t))) for t in v], np.int32)
np.savez_compressed(os.path.join(out, "prims.npz"), **d)

e = {}
small = [(120, 160, 300, 4, 31), (240, 320, 500, 6, 32), (200, 300, 500, 5, 5)]
for i, (h, w, nf, nl, seed) in enumerate(small):
    img = synth.frame(h, w, seed)
    k, dsc = O.RefExtractor(nf, 1.2, nl, 15, 5)(img)
    e[f"small{i}_cfg"] = np.array([h, w, nf, nl, seed], np.int32)
    e[f"small{i}_img_crc"] = np.array(synth.crc(img), np.uint32)
    e[f"small{i}_kps"] = k.view(np.uint8).reshape(-1, 28)
    e[f"small{i}_desc"] = dsc
big = [(480, 640, 1000, 8, 1), (720, 1280, 2000, 8, 2), (384, 384, 1000, 8, 3), (400, 950, 2000, 8, 4),
       (2160, 3840, 8000, 12, 5)]          # BASELINE config C5 (stress): verbatim-reference fixture
rows = []
for (h, w, nf, nl, seed) in big:
    img = synth.frame(h, w, seed)
    k, dsc = O.RefExtractor(nf, 1.2, nl, 15, 5)(img)
    rows.append([h, w, nf, nl, seed, synth.crc(img), len(k), zlib.crc32(k.tobytes()) & 0xFFFFFFFF, zlib.crc32(dsc.tobytes()) & 0xFFFFFFFF])
e["big"] = np.array(rows, np.int64)
np.savez_compressed(os.path.join(out, "extract.npz"), **e)
print({k: os.path.getsize(os.path.join(out, k)) for k in os.listdir(out)})
