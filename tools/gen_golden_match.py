#!/usr/bin/env python3
"""tests/golden/match.npz: outputs of the reference's OWN matcher (src/ORBmatcher.cc compiled verbatim,
oracle/_ref/libfbe_refmatch.so -- build container only) on the seeded scenes of tests/test_oracle_vs_refmatch.py."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
import test_oracle_vs_refmatch as T
assert O.refmatch() is not None, "build oracle/_ref first (make -C oracle refmatch)"
d = {}
for seed in T.SEEDS:
    for k, v in T.scene_outputs(O.RefMatch(), seed, True).items():
        d[f"s{seed}_{k}"] = v
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "match.npz"), **d)
print("written", len(d), "arrays")
