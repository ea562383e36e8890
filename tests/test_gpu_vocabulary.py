"""DBoW2 vocabulary descent on the GPU (fbe_vocabulary_create + fbe_bow_transform through the C-ABI) against the oracle, and the
complete transform(features, BowVector&, FeatureVector&, levelsup) against the verbatim DBoW2 outputs (tests/golden/vocabulary.npz)."""
import os

import numpy as np
import pytest

from voc_scenes import query_descriptors, random_vocabulary, write_text

pytestmark = pytest.mark.gpu
import test_vocabulary as T


@pytest.mark.parametrize("case", T.CASES)
def test_gpu_descent_equals_oracle_and_reference(oracle, case, tmp_path):
    from fishbirdeyevisualslam_b200.vocabulary import Vocabulary
    seed, k, L, ragged, levelsup = case
    voc = random_vocabulary(seed, k, L, ragged)
    q = query_descriptors(seed, voc)
    V = Vocabulary(k, L, voc[2], voc[3], voc[4], voc[5])
    w_g, n_g, wt_g = V.transform_features(q, levelsup)
    w_o, n_o, wt_o = oracle.bow_transform(L, voc[2], voc[3], voc[4], voc[5], q, levelsup)
    assert np.array_equal(w_g, w_o) and np.array_equal(n_g, n_o) and np.array_equal(wt_g.view(np.int64), wt_o.view(np.int64))
    ids, vals, (fid, fst, fit) = V.transform(q, levelsup)
    g = np.load(T.GOLD)
    assert np.array_equal(ids, g[f"c{seed}_ids"]) and np.array_equal(vals.view(np.int64), g[f"c{seed}_vals"].view(np.int64))
    assert np.array_equal(fid, g[f"c{seed}_fid"]) and np.array_equal(fst, g[f"c{seed}_fst"]) and np.array_equal(fit, g[f"c{seed}_fit"])
    # the text-file route gives the same tree
    path = str(tmp_path / "voc.txt")
    write_text(path, voc)
    V2 = Vocabulary.from_text(path)
    assert V2.size() == V.size() == int(voc[3].sum())
    w2, n2, _ = V2.transform_features(q[:100], levelsup)
    assert np.array_equal(w2, w_o[:100]) and np.array_equal(n2, n_o[:100])


def test_orbvoc_sized_tree(oracle):
    """k = 10, L = 5 (111 110 nodes, 3.5 MB of node descriptors): the descent of 4000 descriptors equals the oracle's."""
    from fishbirdeyevisualslam_b200.vocabulary import Vocabulary
    voc = random_vocabulary(9, 10, 5, False)
    q = query_descriptors(9, voc, n=4000)
    V = Vocabulary(10, 5, voc[2], voc[3], voc[4], voc[5])
    w_g, n_g, wt_g = V.transform_features(q, 4)
    w_o, n_o, wt_o = oracle.bow_transform(5, voc[2], voc[3], voc[4], voc[5], q, 4)
    assert np.array_equal(w_g, w_o) and np.array_equal(n_g, n_o) and np.array_equal(wt_g, wt_o)
    assert len(np.unique(n_o)) >= 8 and len(np.unique(w_o)) > 1000        # a duplicated sibling never wins a tie
    e_w, e_n, e_t = V.transform_features(np.zeros((0, 32), np.uint8), 4)
    assert len(e_w) == 0
