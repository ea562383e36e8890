"""DBoW2 vocabulary descent: the restated oracle + the host-side folding against the reference's OWN
TemplatedVocabulary (loadFromTextFile + transform, compiled verbatim: oracle/_ref/libfbe_refvoc.so), and against its
committed outputs (tests/golden/vocabulary.npz, tools/gen_golden_vocabulary.py) where that build is absent.  CPU only."""
import os

import numpy as np
import pytest

from fishbirdeyevisualslam_b200.vocabulary import fold, parse_text
from voc_scenes import query_descriptors, random_vocabulary, write_text

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vocabulary.npz")
CASES = [(0, 10, 3, False, 4), (1, 10, 3, False, 2), (2, 6, 4, True, 2), (3, 4, 2, False, 4)]      # (seed, k, L, ragged, levelsup)


def oracle_outputs(oracle, case):
    seed, k, L, ragged, levelsup = case
    voc = random_vocabulary(seed, k, L, ragged)
    q = query_descriptors(seed, voc)
    w, nd, wt = oracle.bow_transform(L, voc[2], voc[3], voc[4], voc[5], q, levelsup)
    ids, vals, (fid, fst, fit) = fold(w, nd, wt)
    return voc, q, dict(ids=ids, vals=vals, fid=fid, fst=fst, fit=fit)


@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_verbatim_dbow2(oracle, case, tmp_path):
    if oracle.refvoc() is None:
        pytest.skip("oracle/_ref/libfbe_refvoc.so not built (no /root/reference); the golden test covers this machine")
    voc, q, got = oracle_outputs(oracle, case)
    path = str(tmp_path / "voc.txt")
    write_text(path, voc)
    ids, vals, (fid, fst, fit), size = oracle.ref_voc_transform(path, q, case[4])
    assert size == int(voc[3].sum())
    assert np.array_equal(got["ids"], ids) and np.array_equal(got["vals"].view(np.int64), vals.view(np.int64))     # doubles bit-equal
    assert np.array_equal(got["fid"], fid) and np.array_equal(got["fst"], fst) and np.array_equal(got["fit"], fit)
    assert len(ids) > 10 and abs(vals.sum() - 1.0) < 1e-9
    # the text parser mirror reads back what the reference's loader reads
    k, L, n1, n2, parent, is_word, desc, weight = parse_text(path)
    assert (k, L, n1, n2) == (voc[0], voc[1], 0, 0) and np.array_equal(parent, voc[2]) and np.array_equal(is_word, voc[3])
    assert np.array_equal(desc, voc[4]) and np.array_equal(weight, voc[5])


@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_committed_dbow2_outputs(oracle, case):
    g = np.load(GOLD)
    _, _, got = oracle_outputs(oracle, case)
    for k_, v in got.items():
        ref = g[f"c{case[0]}_{k_}"]
        assert np.array_equal(v.view(np.int64) if v.dtype == np.float64 else v, ref.view(np.int64) if ref.dtype == np.float64 else ref), k_
