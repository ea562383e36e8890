"""Host mirror of the vocabulary calls of the reference (`ORBVocabulary` = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>):
loading the text format and `transform(features, BowVector&, FeatureVector&, levelsup)`.  The per-descriptor tree descent
runs on the GPU (fbe_bow_transform); folding the per-feature results into the BowVector / FeatureVector is the reference's
own few lines of std::map bookkeeping, reproduced here in the same order so that the doubles come out bit-identical."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


def parse_text(path: str):
    """TemplatedVocabulary::loadFromTextFile (TemplatedVocabulary.h:1338-1436): header `k L scoring weighting`, then one row
    per node: parent id, nIsLeaf, 32 descriptor bytes, weight.  -> (k, L, scoring, weighting, parent, is_word, desc, weight)."""
    with open(path) as f:
        lines = f.read().split("\n")
    k, L, n1, n2 = (int(x) for x in lines[0].split()[:4])
    if k < 0 or k > 20 or L < 1 or L > 10 or n1 < 0 or n1 > 5 or n2 < 0 or n2 > 3:
        raise ValueError("Vocabulary loading failure: This is not a correct text file!")
    rows = [ln.split() for ln in lines[1:] if ln.strip()]
    n = len(rows)
    parent = np.array([int(r[0]) for r in rows], np.int32)
    is_word = np.array([int(r[1]) > 0 for r in rows], np.uint8)
    desc = np.array([[int(x) for x in r[2:34]] for r in rows], np.uint8).reshape(n, 32)
    weight = np.array([float(r[34]) for r in rows], np.float64)
    return k, L, n1, n2, parent, is_word, desc, weight


class Vocabulary:
    """Device-resident vocabulary tree.  Scoring L1_NORM / weighting TF_IDF (what ORBvoc.txt declares: `10 6 0 0`) is the
    combination transform() reproduces; other combinations raise."""

    def __init__(self, k, L, parent, is_word, desc, weight, scoring=0, weighting=0, device=0):
        if scoring != 0 or weighting != 0:
            raise NotImplementedError("only L1_NORM scoring with TF_IDF weighting (the reference's vocabulary) is mirrored")
        self._L = _lib.load()
        self.k, self.depth = int(k), int(L)
        parent = np.ascontiguousarray(parent, np.int32); is_word = np.ascontiguousarray(is_word, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8); weight = np.ascontiguousarray(weight, np.float64)
        self.n_words = int(is_word.astype(bool).sum())
        h = C.c_void_p()
        check(self._L.fbe_vocabulary_create(self.k, self.depth, ptr(parent), ptr(is_word), ptr(desc), ptr(weight), len(parent),
                                            C.c_int32(device), C.byref(h)))
        self._h = h

    @classmethod
    def from_text(cls, path, device=0):
        k, L, n1, n2, parent, is_word, desc, weight = parse_text(path)
        return cls(k, L, parent, is_word, desc, weight, n1, n2, device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.fbe_vocabulary_destroy(self._h)
            self._h = None

    __del__ = close

    def size(self):
        return self.n_words

    def transform_features(self, desc, levelsup=4):
        """Per-descriptor (word id, node id at level L - levelsup, word weight)."""
        desc = np.ascontiguousarray(desc, np.uint8)
        n = len(desc)
        w = np.zeros(max(n, 1), np.int32); nd = np.zeros(max(n, 1), np.int32); wt = np.zeros(max(n, 1), np.float64)
        check(self._L.fbe_bow_transform(self._h, ptr(desc), n, C.c_int32(levelsup), ptr(w), ptr(nd), ptr(wt)))
        return w[:n], nd[:n], wt[:n]

    def transform(self, desc, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup) (:1127-1205) -> (bow ids, bow values, (node ids, start, items))."""
        w, nd, wt = self.transform_features(desc, levelsup)
        return fold(w, nd, wt)


def fold(word_id, node_id, weight):
    """:1150-1204 for TF_IDF + L1: v.addWeight(id, w) and fv.addFeature(nid, i) in feature order for w > 0, then v.normalize(L1)
    (BowVector.cpp:63-88: norm accumulated in std::map order)."""
    bow = {}
    fv = {}
    for i in range(len(word_id)):
        wv = float(weight[i])
        if wv > 0:
            wid = int(word_id[i])
            bow[wid] = bow[wid] + wv if wid in bow else wv
            fv.setdefault(int(node_id[i]), []).append(i)
    ids = sorted(bow)
    norm = 0.0
    for k in ids:
        norm += abs(bow[k])
    vals = [bow[k] / norm for k in ids] if norm > 0.0 else [bow[k] for k in ids]
    nodes = sorted(fv)
    start, items = [0], []
    for nid in nodes:
        items.extend(fv[nid])
        start.append(len(items))
    return (np.array(ids, np.int32), np.array(vals, np.float64),
            (np.array(nodes, np.int32), np.array(start, np.int32), np.array(items, np.int32)))
