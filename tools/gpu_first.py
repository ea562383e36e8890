import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import oracle as O
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200.extractor import ORBextractor
for (h, w, nf, nl, seed) in [(480, 640, 1000, 8, 1), (720, 1280, 2000, 8, 2), (384, 384, 1000, 8, 3), (400, 950, 2000, 8, 4), (200, 300, 500, 5, 5)]:
    img = synth.frame(h, w, seed)
    o = O.OracleExtractor(nf, 1.2, nl, 15, 5)
    ko, do = o(img)
    g = ORBextractor(nf, 1.2, nl, 15, 5)
    t = time.time(); kg, dg = g(img); t1 = time.time() - t
    t = time.time(); kg, dg = g(img); t2 = time.time() - t
    ok_pyr = all((g.pyramid_level(l) == o.level_padded(l)).all() for l in range(nl))
    ok_cand = all(np.array_equal(g.debug_candidates(l), o.candidates(l)) for l in range(nl))
    ok_blur = all((o.level_blurred(l) is None) or (g.debug_blurred(l) == o.level_blurred(l)).all() for l in range(nl))
    same_n = len(kg) == len(ko)
    print(h, w, 'pyr', ok_pyr, 'cand', ok_cand, 'blur', ok_blur, 'n', len(kg), len(ko), 't first %.3f second %.4f' % (t1, t2))
    if same_n:
        for f in ('x', 'y', 'size', 'response', 'octave', 'class_id'):
            print('   ', f, np.array_equal(kg[f], ko[f]))
        da = np.abs(kg['angle'] - ko['angle']); print('    angle maxdiff', da.max(), 'n bitdiff', int((kg['angle'].view(np.uint32) != ko['angle'].view(np.uint32)).sum()))
        bad = np.where((dg != do).any(axis=1))[0]
        print('    desc mismatching rows', len(bad), 'of which boundary-flagged', int(o.boundary[bad].sum()), 'angle-diff rows', int((da[bad] > 0).sum()))
    else:
        for l in range(nl):
            print('   level', l, 'oracle nkeys', o.level_nkeys(l))
