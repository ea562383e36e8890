import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import oracle as O
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline
from fishbirdeyevisualslam_b200.matcher import Frame
B = 6
fh, fw, bh, bw = 720, 1280, 384, 384
# a sequence of B*2 pairs: scene drifts by (+3,+2) px per frame
fr = np.stack([synth.frame(fh, fw, 100, (min(3 * i, 8) - 4, min(2 * i, 8) - 4), noise_seed=i) for i in range(2 * B)])
bi = np.stack([synth.frame(bh, bw, 200, (min(2 * i, 8) - 4, min(i, 8) - 4), noise_seed=50 + i) for i in range(2 * B)])
pipe = FrontBirdPipeline(B)
dF = torch.from_numpy(fr).cuda(); dB = torch.from_numpy(bi).cuda()
of = O.OracleExtractor(2000, 1.2, 8, 15, 5); ob = O.OracleExtractor(1000, 1.2, 8, 15, 5)
prevF = prevB = None
allok = True
for step in range(2):
    pipe.step_dev(dF[step * B:].data_ptr(), dB[step * B:].data_ptr())
    res, fm, bm = pipe.fetch()
    for p in range(B):
        i = step * B + p
        kf, df = of(fr[i]); kb, db = ob(bi[i])
        gk, gd, gbk, gbd = pipe.fetch_pair(p)
        okx = res['n_front'][p] == len(kf) and res['n_bird'][p] == len(kb) and gk[:len(kf)].tobytes() == kf.tobytes() and \
            (gd[:len(kf)] == df).all() and gbk[:len(kb)].tobytes() == kb.tobytes() and (gbd[:len(kb)] == db).all()
        curF = Frame.front(kf, df, fw, fh); curB = Frame.bird(kb, db, bw, bh)
        okm = True
        if prevF is not None:
            pm = np.ascontiguousarray(np.stack([prevF.kps['x'], prevF.kps['y']], 1), np.float32)
            n_o, m_o = O.search_for_initialization(prevF, curF, pm, 100, 0.9, True)
            nb_o, d_o = O.birdview_match(prevB.kps, prevB.desc, curB, 10, 0.9, True)
            m12b = np.full(len(prevB.kps), -1, np.int32)
            # oracle's internal vnMatches12 incl. index-0 matches is not exported; compare counts + DMatch subset
            gm = bm[p][:len(prevB.kps)]
            okm = res['front_matches'][p] == n_o and np.array_equal(fm[p][:len(prevF.kps)], m_o) and res['bird_matches'][p] == nb_o and \
                np.array_equal(np.stack([np.nonzero(gm > 0)[0], gm[gm > 0]], 1), d_o[:, :2])
        else:
            okm = res['front_matches'][p] == 0 and res['bird_matches'][p] == 0
        print('step', step, 'pair', p, 'extract', okx, 'match', okm, dict(zip(res.dtype.names, res[p].tolist())))
        allok &= bool(okx and okm)
        prevF, prevB = curF, curB
print('ALL OK' if allok else 'MISMATCH')
# quick throughput probe
B2 = 64
pipe2 = FrontBirdPipeline(B2)
F2 = torch.from_numpy(synth.cheap_batch(B2, fh, fw, 1)).cuda(); Bd2 = torch.from_numpy(synth.cheap_batch(B2, bh, bw, 2)).cuda()
for _ in range(3): pipe2.step_dev(F2.data_ptr(), Bd2.data_ptr())
pipe2.sync()
t = time.time()
for _ in range(5): pipe2.step_dev(F2.data_ptr(), Bd2.data_ptr())
pipe2.sync(); dt = (time.time() - t) / 5
print('batch', B2, 'step %.2f ms' % (dt * 1e3), 'pairs/s %.0f' % (B2 / dt), 'last_step_ms', pipe2.last_step_ms())
