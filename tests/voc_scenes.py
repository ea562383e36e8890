"""Synthetic DBoW2 vocabularies in the text format of TemplatedVocabulary::loadFromTextFile (no trailing newline: the
reference's `while(!f.eof())` loop would otherwise create a node from the empty last line)."""
import numpy as np


def random_vocabulary(seed, k=10, L=3, ragged=False, stop_frac=0.05):
    """Rows in creation order (breadth first like saveToTextFile writes them): parent, is_word, descriptor, weight."""
    rng = np.random.default_rng(3000 + seed)
    parent, is_word, desc, weight = [], [], [], []
    level_nodes = [0]
    node_desc = {0: rng.integers(0, 256, 32, dtype=np.uint8)}
    for lvl in range(1, L + 1):
        nxt = []
        for p in level_nodes:
            nchild = k if not ragged else int(rng.integers(2, k + 1))
            for c in range(nchild):
                nid = len(parent) + 1
                d = node_desc[p].copy()
                flips = rng.choice(256, int(rng.integers(8, 60 // lvl + 9)), replace=False)
                for b in flips:
                    d[b >> 3] ^= np.uint8(1 << (b & 7))
                if c and rng.random() < 0.1:
                    d = desc[-1].copy()                       # duplicate sibling: a distance tie, the first child must win
                node_desc[nid] = d
                parent.append(p); desc.append(d)
                leaf = lvl == L
                is_word.append(1 if leaf else 0)
                weight.append(0.0 if (leaf and rng.random() < stop_frac) else (float(rng.uniform(0.1, 9.0)) if leaf else 0.0))
                nxt.append(nid)
        level_nodes = nxt
    return (k, L, np.array(parent, np.int32), np.array(is_word, np.uint8), np.array(desc, np.uint8).reshape(-1, 32),
            np.array(weight, np.float64))


def write_text(path, voc):
    k, L, parent, is_word, desc, weight = voc
    rows = [f"{k} {L} 0 0"]
    for i in range(len(parent)):
        rows.append(f"{int(parent[i])} {int(is_word[i])} " + " ".join(str(int(x)) for x in desc[i]) + f" {float(weight[i])!r}")
    with open(path, "w") as f:
        f.write("\n".join(rows))


def query_descriptors(seed, voc, n=1500):
    """Descriptors near random leaves (so that words repeat) plus pure noise."""
    rng = np.random.default_rng(4000 + seed)
    _, _, parent, is_word, desc, _ = voc
    leaves = np.nonzero(is_word)[0]
    pick = rng.choice(leaves, n)
    q = desc[pick].copy()
    for i in range(n):
        for b in rng.choice(256, int(rng.integers(0, 40)), replace=False):
            q[i, b >> 3] ^= np.uint8(1 << (b & 7))
    q[rng.random(n) < 0.1] = rng.integers(0, 256, 32, dtype=np.uint8)
    return q
