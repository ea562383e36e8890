"""Secondary measurements for the BASELINE configs that are not the bench.py headline (C1, C3, C5), through the host-buffer
C-ABI (H2D + kernels + D2H per call, wall clock), each checked against the oracle and timed beside it on one host core.
   python tools/bench_configs.py > gpurun_out/configs.json
One JSON object per line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200.extractor import ORBextractor
from fishbirdeyevisualslam_b200.matcher import Frame, ORBmatcher, BaseXY2BirdPixel
from oracle import oracle as O


def timeit(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return float(np.median(ts))


def c1():
    a, b = synth.frame_pair_in_time(480, 640, 1001)
    ex = ORBextractor(1000, 1.2, 8, 15, 5)
    m = ORBmatcher(0.9, True)

    def gpu():
        ka, da = ex(a); kb, db = ex(b)
        F1, F2 = Frame.front(ka, da, 640, 480), Frame.front(kb, db, 640, 480)
        pm = np.ascontiguousarray(np.stack([ka["x"], ka["y"]], 1), np.float32)
        return ka, da, kb, db, m.SearchForInitialization(F1, F2, pm, 100)
    oe = O.RefExtractor(1000, 1.2, 8, 15, 5) if O.ref() is not None else O.OracleExtractor(1000, 1.2, 8, 15, 5)

    def cpu():
        ka, da = oe(a); kb, db = oe(b)
        F1, F2 = Frame.front(ka, da, 640, 480), Frame.front(kb, db, 640, 480)
        pm = np.ascontiguousarray(np.stack([ka["x"], ka["y"]], 1), np.float32)
        return ka, da, kb, db, O.search_for_initialization(F1, F2, pm, 100, 0.9, True)
    g, c = gpu(), cpu()
    ok = g[0].tobytes() == c[0].tobytes() and (g[1] == c[1]).all() and g[2].tobytes() == c[2].tobytes() and (g[3] == c[3]).all() and \
        g[4][0] == c[4][0] and np.array_equal(g[4][1], c[4][1])
    tg, tc = timeit(gpu, 20), timeit(cpu, 3, 1)
    return {"config": "C1 640x480 pair @1000, extract x2 + SearchForInitialization(window 100)", "parity": bool(ok), "keypoints": [len(g[0]), len(g[2])],
            "matches": int(g[4][0]), "gpu_ms_per_pair_host_api": tg * 1e3, "cpu_1core_ms_per_pair": tc * 1e3}


def c3():
    rng = np.random.default_rng(3003)
    img = synth.frame(384, 384, 3003)
    kb, db = ORBextractor(2000, 1.2, 8, 15, 5)(img)
    FB = Frame.bird(kb, db, 384, 384)
    nmp = 20000
    src = rng.integers(0, FB.N, nmp)
    mp_desc = FB.desc[src].copy()
    nflip = rng.integers(0, 41, nmp)
    bitpos = rng.integers(0, 256, (nmp, 40))
    for j in range(40):
        sel = nflip > j
        mp_desc[sel, bitpos[sel, j] >> 3] ^= (1 << (bitpos[sel, j] & 7)).astype(np.uint8)
    pix = np.stack([FB.kps["x"][src], FB.kps["y"][src]], 1).astype(np.float32) + rng.uniform(-2.5, 2.5, (nmp, 2)).astype(np.float32)
    pix[rng.random(nmp) < 0.03, 0] = np.nan
    pix = np.ascontiguousarray(pix)
    import ctypes as C
    from fishbirdeyevisualslam_b200._lib import check, ptr
    m = ORBmatcher(0.9, True)

    def gpu():
        m12 = np.full(nmp, -1, np.int32); n = C.c_int32(); v = FB.view()
        check(m._L.fbe_bird_map_point_match(m._h, ptr(pix), ptr(mp_desc), nmp, C.byref(v), 10, ptr(m12), C.byref(n)))
        return n.value, m12
    cpu = lambda: O.bird_map_point_match(pix, mp_desc, FB, 10, 0.9)
    g, c = gpu(), cpu()
    ok = g[0] == c[0] and np.array_equal(g[1], c[1])
    tg, tc = timeit(gpu, 20), timeit(cpu, 3, 1)
    return {"config": "C3 BirdMapPointMatch 20000 map points vs %d bird keypoints, window 10" % FB.N, "parity": bool(ok), "matches": int(g[0]),
            "gpu_ms_host_api": tg * 1e3, "cpu_1core_ms": tc * 1e3}


def c5():
    img = synth.frame(2160, 3840, 5005)
    ex = ORBextractor(8000, 1.2, 12, 15, 5)
    oe = O.RefExtractor(8000, 1.2, 12, 15, 5) if O.ref() is not None else O.OracleExtractor(8000, 1.2, 12, 15, 5)
    kg, dg = ex(img)
    ko, do = oe(img)
    ok = kg.tobytes() == ko.tobytes() and (dg == do).all()
    t_ext, t_ext_cpu = timeit(lambda: ex(img), 10), timeit(lambda: oe(img), 2, 0)
    m = ORBmatcher(0.9, True)
    rng = np.random.default_rng(5)
    q = np.ascontiguousarray(dg[:8000]); t = np.ascontiguousarray(np.roll(dg[:8000], 17, 0))
    bi, bd, sd = m.BruteForceTop2(q, t)
    obi, obd, osd = O.bruteforce_top2(q[:500], t)
    ok_bf = np.array_equal(bi[:500], obi) and np.array_equal(bd[:500], obd) and np.array_equal(sd[:500], osd)
    t_bf = timeit(lambda: m.BruteForceTop2(q, t), 10)
    t_bf_cpu = timeit(lambda: O.bruteforce_top2(q[:500], t), 1, 0) * len(q) / 500
    return {"config": "C5 3840x2160 @8000, 12 levels: extract + brute-force %dx%d Hamming top-2" % (len(q), len(t)), "parity_extract": bool(ok),
            "parity_bruteforce(500 queries)": bool(ok_bf), "keypoints": int(len(kg)), "gpu_extract_ms_host_api": t_ext * 1e3,
            "cpu_1core_extract_ms": t_ext_cpu * 1e3, "gpu_bruteforce_ms_host_api": t_bf * 1e3, "cpu_1core_bruteforce_ms(extrapolated)": t_bf_cpu * 1e3,
            "hamming_pairs_per_s": len(q) * len(t) / t_bf}


if __name__ == "__main__":
    for f in (c1, c3, c5):
        print(json.dumps(f()), flush=True)
