#!/usr/bin/env python3
"""tests/golden/vocabulary.npz: outputs of the reference's OWN DBoW2 TemplatedVocabulary (loadFromTextFile + transform, compiled
verbatim: oracle/_ref/libfbe_refvoc.so -- build container only) on the seeded vocabularies of tests/test_vocabulary.py."""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O
import test_vocabulary as T
from voc_scenes import query_descriptors, random_vocabulary, write_text
assert O.refvoc() is not None, "build oracle/_ref first (make -C oracle refvoc)"
d = {}
for case in T.CASES:
    seed, k, L, ragged, levelsup = case
    voc = random_vocabulary(seed, k, L, ragged)
    q = query_descriptors(seed, voc)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "voc.txt")
        write_text(path, voc)
        ids, vals, (fid, fst, fit), _ = O.ref_voc_transform(path, q, levelsup)
    d.update({f"c{seed}_ids": ids, f"c{seed}_vals": vals, f"c{seed}_fid": fid, f"c{seed}_fst": fst, f"c{seed}_fit": fit})
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vocabulary.npz"), **d)
print("written", len(d), "arrays")
