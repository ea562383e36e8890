"""The device-resident batch pipeline (config C2/C4 shape): per-pair outputs equal single-image extraction + oracle
matching; results do not depend on the batch size (shard equality); repeated steps are deterministic."""
import numpy as np
import pytest

from fishbirdeyevisualslam_b200 import synth

pytestmark = pytest.mark.gpu

FH, FW, BH, BW = 720, 1280, 384, 384


def sequence(n, seed):
    fr = np.stack([synth.frame(FH, FW, seed, (min(3 * i, 8) - 4, min(2 * i, 8) - 4), noise_seed=i) for i in range(n)])
    bi = np.stack([synth.frame(BH, BW, seed + 1, (min(2 * i, 8) - 4, min(i, 8) - 4), noise_seed=50 + i) for i in range(n)])
    return fr, bi


def run_pipeline(fr, bi, batch):
    import torch
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline
    pipe = FrontBirdPipeline(batch)
    dF, dB = torch.from_numpy(fr).cuda(), torch.from_numpy(bi).cuda()
    out = []
    for s in range(len(fr) // batch):
        pipe.step_dev(dF[s * batch:].data_ptr(), dB[s * batch:].data_ptr())
        res, fm, bm = pipe.fetch()
        for p in range(batch):
            fk, fd, bk, bd = pipe.fetch_pair(p)
            nf, nb = int(res["n_front"][p]), int(res["n_bird"][p])
            out.append(dict(res=res[p].copy(), fk=fk[:nf].copy(), fd=fd[:nf].copy(), bk=bk[:nb].copy(), bd=bd[:nb].copy(),
                            fm=fm[p].copy(), bm=bm[p].copy()))
    pipe.close()
    return out


def test_pipeline_equals_oracle_and_carries_across_steps(oracle):
    from fishbirdeyevisualslam_b200.matcher import Frame
    n, B = 8, 4
    fr, bi = sequence(n, 100)
    out = run_pipeline(fr, bi, B)
    of, ob = oracle.OracleExtractor(2000, 1.2, 8, 15, 5), oracle.OracleExtractor(1000, 1.2, 8, 15, 5)
    prev = None
    for i, r in enumerate(out):
        kf, df = of(fr[i])
        kb, db = ob(bi[i])
        assert r["fk"].tobytes() == kf.tobytes() and np.array_equal(r["fd"], df)
        assert r["bk"].tobytes() == kb.tobytes() and np.array_equal(r["bd"], db)
        F, Bf = Frame.front(kf, df, FW, FH), Frame.bird(kb, db, BW, BH)
        if prev is None:
            assert r["res"]["front_matches"] == 0 and r["res"]["bird_matches"] == 0
        else:
            pm = np.ascontiguousarray(np.stack([prev[0].kps["x"], prev[0].kps["y"]], 1), np.float32)
            n_o, m_o = oracle.search_for_initialization(prev[0], F, pm, 100, 0.9, True)
            nb_o, d_o = oracle.birdview_match(prev[1].kps, prev[1].desc, Bf, 10, 0.9, True)
            assert r["res"]["front_matches"] == n_o and np.array_equal(r["fm"][:prev[0].N], m_o) and n_o > 100
            gm = r["bm"][:prev[1].N]
            assert r["res"]["bird_matches"] == nb_o and nb_o > 50
            assert np.array_equal(np.stack([np.nonzero(gm > 0)[0], gm[gm > 0]], 1), d_o[:, :2])
        prev = (F, Bf)


def test_shard_equality_and_determinism():
    """Processing 12 pairs as 1x12, 2x6 or 4x3 (or twice) gives byte-identical per-pair results."""
    fr, bi = sequence(12, 300)
    ref = run_pipeline(fr, bi, 12)
    for batch in (6, 3, 12):
        got = run_pipeline(fr, bi, batch)
        for i, (a, b) in enumerate(zip(ref, got)):
            assert a["res"].tobytes() == b["res"].tobytes()
            assert a["fk"].tobytes() == b["fk"].tobytes() and np.array_equal(a["fd"], b["fd"])
            assert a["bk"].tobytes() == b["bk"].tobytes() and np.array_equal(a["bd"], b["bd"])
            if i:       # match lists are indexed by the PREVIOUS pair's keypoints
                nq, nb = len(ref[i - 1]["fk"]), len(ref[i - 1]["bk"])
                assert np.array_equal(a["fm"][:nq], b["fm"][:nq]) and np.array_equal(a["bm"][:nb], b["bm"][:nb])


def test_host_step_equals_device_step():
    import torch
    from fishbirdeyevisualslam_b200 import _lib
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline, PinnedBuffer
    B = 4
    fr, bi = sequence(B, 500)
    p1, p2 = FrontBirdPipeline(B), FrontBirdPipeline(B)
    dF, dB = torch.from_numpy(fr).cuda(), torch.from_numpy(bi).cuda()
    p1.step_dev(dF.data_ptr(), dB.data_ptr())
    r1, f1, b1 = p1.fetch()
    hF, hB = PinnedBuffer(fr.shape), PinnedBuffer(bi.shape)
    hF.array[...] = fr
    hB.array[...] = bi
    r2 = np.zeros(B, _lib.PAIR_RESULT_DTYPE)
    f2 = np.zeros((B, p2.front_cap), np.int32)
    b2 = np.zeros((B, p2.bird_cap), np.int32)
    p2.step_host(hF.ptr, hB.ptr, r2, f2, b2)
    assert r1.tobytes() == r2.tobytes()
    for p in range(1, B):
        nq, nb = int(r1["n_front"][p - 1]), int(r1["n_bird"][p - 1])
        assert np.array_equal(f1[p][:nq], f2[p][:nq]) and np.array_equal(b1[p][:nb], b2[p][:nb])


def test_async_submit_equals_sync_steps():
    """Up to three steps in flight through host buffers (H2D of the next steps beside the kernels of step N) give exactly
    the results of the synchronous step-by-step path."""
    from fishbirdeyevisualslam_b200 import _lib
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline, PinnedBuffer
    B, S = 3, 6
    fr, bi = sequence(B * S, 700)
    p1, p2 = FrontBirdPipeline(B), FrontBirdPipeline(B)
    want = []
    hF, hB = [PinnedBuffer((B,) + fr.shape[1:]) for _ in range(S)], [PinnedBuffer((B,) + bi.shape[1:]) for _ in range(S)]
    for s in range(S):
        hF[s].array[...] = fr[s * B:(s + 1) * B]
        hB[s].array[...] = bi[s * B:(s + 1) * B]
        r = np.zeros(B, _lib.PAIR_RESULT_DTYPE)
        f = np.zeros((B, p1.front_cap), np.int32)
        b = np.zeros((B, p1.bird_cap), np.int32)
        p1.step_host(hF[s].ptr, hB[s].ptr, r, f, b)
        want.append((r, f, b))
    res = [PinnedBuffer((B,), _lib.PAIR_RESULT_DTYPE) for _ in range(S)]
    fm = [PinnedBuffer((B, p2.front_cap), np.int32) for _ in range(S)]
    bm = [PinnedBuffer((B, p2.bird_cap), np.int32) for _ in range(S)]
    tickets = []
    for s in range(S):
        tickets.append(p2.submit_host(hF[s].ptr, hB[s].ptr, res[s].array, fm[s].array, bm[s].array))
        if s >= 2:
            p2.wait(tickets[s - 2])
    p2.wait(tickets[-2])
    p2.wait(tickets[-1])
    prev_nf = prev_nb = 0
    for s in range(S):
        r, f, b = want[s]
        assert res[s].array.tobytes() == r.tobytes()
        for p in range(B):
            if s or p:
                assert np.array_equal(fm[s].array[p][:prev_nf], f[p][:prev_nf]) and np.array_equal(bm[s].array[p][:prev_nb], b[p][:prev_nb])
            prev_nf, prev_nb = int(r["n_front"][p]), int(r["n_bird"][p])


def test_bench_size_properties():
    """Size-independent properties at the bench shape (128 pairs per step, consecutive frames): the same 256 pairs as 2 x 128
    and 4 x 64 give identical records and match lists; every match obeys the reference's acceptance rules; a sample of
    pairs equals the oracle extraction bit for bit."""
    import torch
    from fishbirdeyevisualslam_b200 import _lib
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline
    from oracle import oracle as O
    n = 256
    fr, bi = synth.cheap_batch(n, FH, FW, 900), synth.cheap_batch(n, BH, BW, 901)
    dF, dB = torch.from_numpy(fr).cuda(), torch.from_numpy(bi).cuda()

    def run(batch, keep=()):
        pipe = FrontBirdPipeline(batch)
        recs, fms, bms, kept = [], [], [], {}
        for s in range(n // batch):
            pipe.step_dev(dF[s * batch:].data_ptr(), dB[s * batch:].data_ptr())
            res, fm, bm = pipe.fetch()
            recs.append(res.copy()); fms.append(fm.copy()); bms.append(bm.copy())
            for g in keep:
                if s * batch <= g < (s + 1) * batch:
                    kept[g] = pipe.fetch_pair(g - s * batch)
        pipe.close()
        return np.concatenate(recs), np.concatenate(fms), np.concatenate(bms), kept

    sample = (0, 63, 64, 127, 128, 200, 255)
    r1, f1, b1, k1 = run(128, sample + tuple(g - 1 for g in sample if g))
    r2, f2, b2, _ = run(64)
    assert r1.tobytes() == r2.tobytes()
    assert (r1["n_front"] >= 2000).all() and (r1["n_bird"] >= 1000).all()
    for g in range(1, n):
        nq, nb = int(r1["n_front"][g - 1]), int(r1["n_bird"][g - 1])
        assert np.array_equal(f1[g][:nq], f2[g][:nq]) and np.array_equal(b1[g][:nb], b2[g][:nb])
        # front matches: one-to-one (the steal rule leaves at most one query per target), count = record
        m = f1[g][:nq]
        hit = m[m >= 0]
        assert len(hit) == r1["front_matches"][g] and len(np.unique(hit)) == len(hit) and (hit < r1["n_front"][g]).all()
        assert (b1[g][:nb] >= 0).sum() == r1["bird_matches"][g]
    # accepted front matches obey TH_LOW on the actual descriptors, queries are octave-0 keypoints
    for g in sample:
        if g == 0:
            continue
        (qk, qd, _, _), (tk, td, _, _) = k1[g - 1], k1[g]
        m = f1[g][:int(r1["n_front"][g - 1])]
        qi = np.nonzero(m >= 0)[0]
        assert (qk["octave"][qi] == 0).all()
        d = np.unpackbits(qd[qi] ^ td[m[qi]], axis=1).sum(1)
        assert (d <= 50).all()
    # sampled pairs equal the oracle extraction
    of, ob = O.OracleExtractor(2000, 1.2, 8, 15, 5), O.OracleExtractor(1000, 1.2, 8, 15, 5)
    for g in (0, 127, 128, 255):
        fk, fd, bk, bd = k1[g]
        kf, df = of(fr[g]); kb, db = ob(bi[g])
        assert fk[:len(kf)].tobytes() == kf.tobytes() and np.array_equal(fd[:len(kf)], df)
        assert bk[:len(kb)].tobytes() == kb.tobytes() and np.array_equal(bd[:len(kb)], db)


def test_fisheye_front_camera(oracle):
    """The reference's real front camera (fisheye.yaml: k1 != 0): keypoints are undistorted ON THE DEVICE between descriptors
    and grid, the grid spans the undistorted image bounds, and the frame-to-frame search runs on mvKeysUn -- the order of
    the reference's Frame constructor (ExtractORB -> UndistortKeyPoints -> AssignFeaturesToGrid)."""
    import torch
    from fishbirdeyevisualslam_b200.matcher import ComputeImageBounds, Frame
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline
    K, D = (348.5 * 1280 / 960, 347.0 * 720 / 600, 640.0, 362.0), (-0.0488316, 0.000298406, -0.00591118, 0.00193258)
    n, B = 4, 2
    fr, bi = sequence(n, 800)
    pipe = FrontBirdPipeline(B, front_fisheye=(K, D))
    dF, dB = torch.from_numpy(fr).cuda(), torch.from_numpy(bi).cuda()
    of = oracle.OracleExtractor(2000, 1.2, 8, 15, 5)
    bounds_want = ComputeImageBounds(FW, FH, K, D)
    prev = None
    for s in range(n // B):
        pipe.step_dev(dF[s * B:].data_ptr(), dB[s * B:].data_ptr())
        res, fm, bm = pipe.fetch()
        for p in range(B):
            i = s * B + p
            fk, fd, _, _ = pipe.fetch_pair(p)
            fu, bounds = pipe.fetch_front_undistorted(p)
            kf, df = of(fr[i])
            nf = len(kf)
            assert fk[:nf].tobytes() == kf.tobytes() and np.array_equal(fd[:nf], df)            # mvKeys stay distorted
            up = oracle.fisheye_undistort(np.stack([kf["x"], kf["y"]], 1), K, D)
            un = kf.copy()
            un["x"], un["y"] = up[:, 0], up[:, 1]
            assert fu[:nf].tobytes() == un.tobytes() and bounds == bounds_want
            assert np.abs(un["x"] - kf["x"]).max() > 2.0                                        # the model really moves points
            F = Frame(un, df, bounds[0], bounds[2], float(np.float32(64) / (np.float32(bounds[1]) - np.float32(bounds[0]))),
                      float(np.float32(48) / (np.float32(bounds[3]) - np.float32(bounds[2]))), 64, 48)
            if prev is not None:
                pm = np.ascontiguousarray(np.stack([prev.kps["x"], prev.kps["y"]], 1), np.float32)
                n_o, m_o = oracle.search_for_initialization(prev, F, pm, 100, 0.9, True)
                assert res["front_matches"][p] == n_o and np.array_equal(fm[p][:prev.N], m_o) and n_o > 100
            prev = F
    pipe.close()


@pytest.mark.parametrize("mode", ["dev", "host"])
def test_soak_repeated_steps_reproduce(mode):
    """tools/soak.py in short form: four input batches cycled for 40 steps (device-resident or through the asynchronous
    host path with three steps in flight); every step must reproduce the first pass over the same (batch, predecessor)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "soak.py"), "40", "12", mode], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches: 0" in r.stdout



def test_submit_host_returns_features():
    """The host-buffer step that also returns what ORBextractor::operator() returns (keypoints + descriptors of every frame):
    equal to the per-pair fetch of the same step, for three steps in flight."""
    from fishbirdeyevisualslam_b200 import _lib
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline, PinnedBuffer
    B, S = 3, 4
    fr, bi = sequence(B * S, 900)
    ref = run_pipeline(fr, bi, B)
    pipe = FrontBirdPipeline(B)
    hF, hB = [PinnedBuffer((B,) + fr.shape[1:]) for _ in range(S)], [PinnedBuffer((B,) + bi.shape[1:]) for _ in range(S)]
    for s in range(S):
        hF[s].array[...] = fr[s * B:(s + 1) * B]
        hB[s].array[...] = bi[s * B:(s + 1) * B]
    res = [PinnedBuffer((B,), _lib.PAIR_RESULT_DTYPE) for _ in range(S)]
    fm = [PinnedBuffer((B, pipe.front_cap), np.int32) for _ in range(S)]
    bm = [PinnedBuffer((B, pipe.bird_cap), np.int32) for _ in range(S)]
    feat = [(PinnedBuffer((B, pipe.front_cap), _lib.KP_DTYPE), PinnedBuffer((B, pipe.front_cap, 32), np.uint8),
             PinnedBuffer((B, pipe.bird_cap), _lib.KP_DTYPE), PinnedBuffer((B, pipe.bird_cap, 32), np.uint8)) for _ in range(S)]
    q = []
    for s in range(S):
        q.append(pipe.submit_host(hF[s].ptr, hB[s].ptr, res[s].array, fm[s].array, bm[s].array, features=tuple(b.array for b in feat[s])))
        if len(q) >= 3:
            pipe.wait(q.pop(0))
    for t in q:
        pipe.wait(t)
    for s in range(S):
        for p in range(B):
            r = ref[s * B + p]
            nf, nb = int(res[s].array["n_front"][p]), int(res[s].array["n_bird"][p])
            assert res[s].array[p].tobytes() == r["res"].tobytes()
            assert feat[s][0].array[p][:nf].tobytes() == r["fk"].tobytes() and np.array_equal(feat[s][1].array[p][:nf], r["fd"])
            assert feat[s][2].array[p][:nb].tobytes() == r["bk"].tobytes() and np.array_equal(feat[s][3].array[p][:nb], r["bd"])
    pipe.close()


def test_row_capacity_overflow_is_reported_per_step():
    """A search window with more candidates than front_row_cap fails THAT step with FBE_E_CAPACITY (naming the pair); the
    handle stays usable: a later step without such a window succeeds."""
    import torch
    from fishbirdeyevisualslam_b200 import _lib
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline
    B = 2
    fr, bi = sequence(2 * B, 1100)
    pipe = FrontBirdPipeline(B, front_row_cap=4)
    dF, dB = torch.from_numpy(fr).cuda(), torch.from_numpy(bi).cuda()
    pipe.step_dev(dF.data_ptr(), dB.data_ptr())
    with pytest.raises(_lib.FbeError) as ei:
        pipe.fetch()
    assert ei.value.code == _lib.FBE_E_CAPACITY and "pair" in str(ei.value) and "front_row_cap" in str(ei.value)
    flatF, flatB = torch.full_like(dF[:B], 90), torch.full_like(dB[:B], 90)        # no corners -> no candidates -> no overflow
    pipe.step_dev(flatF.data_ptr(), flatB.data_ptr())
    pipe.step_dev(flatF.data_ptr(), flatB.data_ptr())
    res, _, _ = pipe.fetch()
    assert (res["n_front"] == 0).all() and (res["front_matches"] == 0).all()
    pipe.close()
    ok = FrontBirdPipeline(B, front_row_cap=1024)
    ok.step_dev(dF.data_ptr(), dB.data_ptr())
    ok.step_dev(dF[B:].data_ptr(), dB[B:].data_ptr())
    res, _, _ = ok.fetch()
    assert (res["front_matches"] > 100).all()
    ok.close()
