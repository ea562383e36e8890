import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import oracle as O
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200.extractor import ORBextractor
h, w, nf, nl, seed = 200, 300, 500, 5, 5
img = synth.frame(h, w, seed)
o = O.OracleExtractor(nf, 1.2, nl, 15, 5); o(img)
g = ORBextractor(nf, 1.2, nl, 15, 5); g(img)
for l in range(nl):
    a = g.debug_candidates(l); b = o.candidates(l)
    sa = set(map(tuple, a.tolist())); sb = set(map(tuple, b.tolist()))
    print('level', l, 'gpu', len(a), 'oracle', len(b), 'common', len(sa & sb), 'only gpu', sorted(sa - sb)[:8], 'only oracle', sorted(sb - sa)[:8])
    if len(a) == len(b) and sa == sb:
        d = np.where((a != b).any(axis=1))[0]
        print('   same set; first order diff at', d[:3], a[d[:3]].tolist(), b[d[:3]].tolist())
L0 = o.level_padded(0)
y, x = 132 + 19, 19 + 19
dx = [0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1]; dy = [3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3]
print('oracle center', L0[y, x], 'ring', [int(L0[y + dy[k], x + dx[k]]) - int(L0[y, x]) for k in range(16)])
