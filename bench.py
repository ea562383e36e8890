#!/usr/bin/env python3
"""bench.py -- front+bird frame-pairs/sec, ORB extract + grid + match (BASELINE.json metric, config C4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--pairs P] [--impl ours|reference]

Workload (both arms, BASELINE.json configs[3] on the pair shape of configs[1]): an offline job of P = 4096 synthetic frame pairs
per GPU -- front 1280x720 @2000 features + bird 384x384 @1000 features -- sharded by frame (rank r owns pairs
[r*P, (r+1)*P); no data-path collective), processed in steps of B = 128 pairs: extract both views, build both grids, match
every pair against its predecessor (front: SearchForInitialization, window 100; bird: BirdviewMatch, window 10).  The P
pairs are P/B DISTINCT batches (distinct seeds per rank and batch); steps cycle through them.

ours      : the sm_100a pipeline through the C-ABI.  `value` = device-timed pairs/s with all P pairs resident in HBM;
            `e2e` = the same through HOST (pinned) buffers: every step's input H2D and the D2H of what the reference's calls
            return (keypoints + descriptors of every frame, match lists, counts) are inside the timed region.
reference : the reference's own CPU code on the host cores: ORBextractor.cc and ORBmatcher.cc compiled verbatim
            (oracle/_ref, -O3 timing variants) on all host threads, a bounded sample of the same pairs per step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRONT = (720, 1280)
BIRD = (384, 384)
FRONT_FEATURES, BIRD_FEATURES = 2000, 1000
# B_img = P0 + 2*P>=1 + 64*n_kp ; front 4 912 576 B + bird 829 464 B
BYTES_PER_PAIR = 5_742_040
# pixels of all pyramid levels (what the FAST kernel reads once): front 2 853 088, bird 456 460
PYR_PIXELS_FRONT, PYR_PIXELS_BIRD = 2_853_088, 456_460
# dram__bytes_read.sum + dram__bytes_write.sum of one front k_fast_cells launch over 128 images (ncu --set full,
# the k_fast_cells (406, 128, 1) row of profiles/r2_29_stage_ncu.md): 384.71 MB + 36.98 MB
FAST_DRAM_BYTES_PER_FRONT_IMAGE = int((384.710400e6 + 36.979200e6) / 128)
# warp-instructions executed by the same launch (smsp__inst_executed.sum of the same capture)
FAST_WARP_INST_PER_FRONT_IMAGE = 1_024_480_494 / 128
FAST_PROFILE = "profiles/r2_29_stage_ncu.md (source-level detail: profiles/r2_01_fast_blur_ncu.md)"
# warp-instructions of ALL kernels of one 128-pair step (sum of smsp__inst_executed.sum over profiles/r2_29_stage_ncu.md)
STEP_WARP_INST_PER_PAIR = 2_250.3e6 / 128
METRIC = "front+bird frame-pairs/sec ORB extract+match at 1/2/4/8 B200 vs host CPU ref"
WORKLOAD = ("C4: offline job of 4096 synthetic front(1280x720 @2000)+bird(384x384 @1000) frame pairs per GPU, sharded by frame; "
            "extract + grid + frame-to-frame match (pair shape of C2)")


def config_dict(world: int, batch: int, pairs: int) -> dict:
    """The SAME dict in both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "pairs_per_gpu": pairs, "pairs_per_step_per_gpu": batch, "distinct_batches_per_gpu": pairs // batch,
            "seeds": "front 10000 + 100*rank + batch, bird 20000 + 100*rank + batch (synth.cheap_batch)",
            "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
            "l2": f"{batch * (FRONT[0] * FRONT[1] + BIRD[0] * BIRD[1]) / 1e6:.0f} MB of distinct input per step per GPU (> 126 MB L2), "
                  f"{pairs // batch} distinct batches cycled"}


def batch_seeds(rank: int, k: int):
    return 10000 + 100 * rank + k, 20000 + 100 * rank + k


def synth_batches(rank: int, nbatch: int, batch: int, out_front: np.ndarray, out_bird: np.ndarray):
    """Fill out_front [nbatch, batch, 720, 1280] / out_bird [nbatch, batch, 384, 384] with this rank's distinct batches
    (host threads in parallel: numpy releases the GIL inside the big array operations)."""
    from concurrent.futures import ThreadPoolExecutor
    from fishbirdeyevisualslam_b200 import synth

    def one(k):
        sf, sb = batch_seeds(rank, k)
        out_front[k] = synth.cheap_batch(batch, FRONT[0], FRONT[1], sf)
        out_bird[k] = synth.cheap_batch(batch, BIRD[0], BIRD[1], sb)

    with ThreadPoolExecutor(min(nbatch, max(2, (os.cpu_count() or 4)))) as ex:
        list(ex.map(one, range(nbatch)))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed regions: NVML (every ~2 ms) when pynvml is importable,
    else `nvidia-smi --query-gpu` (every ~100 ms)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.sm, self.mx, self.seen, self.stop, self.index, self.src = [], [], set(), False, index, "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.src = "nvml"
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        if self.nvml is not None:
            n = self.nvml
            bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)), "hw_slowdown"),
                    (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)), "hw_thermal_slowdown"),
                    (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)), "sw_thermal_slowdown"),
                    (getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)), "sw_power_cap")]
            get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
            try:
                self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
            except Exception:
                pass
            while not self.stop:
                try:
                    self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                    r = int(get_reasons(self.h))
                    for b, name in bits:
                        if r & b:
                            self.seen.add(name)
                except Exception:
                    pass
                time.sleep(0.002)
            return
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    r = [c.strip() for c in out.splitlines()[0].split(",")]
                    if r[0].replace(".", "").isdigit():
                        self.sm.append(float(r[0]))
                    if r[1].replace(".", "").isdigit():
                        self.mx.append(float(r[1]))
                    for i, name in enumerate(self.NAMES):
                        if len(r) > 2 + i and r[2 + i].lower().startswith("active"):
                            self.seen.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"], "source": self.src}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": [n for n in self.NAMES if n in self.seen], "samples": len(self.sm), "source": self.src}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_libs():
    """(extractor factory, matcher, kind, build) -- the verbatim reference builds (timing variants when the host runs them),
    else the oracle port."""
    from oracle import oracle as O
    O.lib()
    if O.ref(True) is not None and O.refmatch(True) is not None:
        return (lambda nf: O.RefExtractor(nf, 1.2, 8, 15, 5, timing=True)), O.RefMatch(O.refmatch(True)), "reference", \
            "verbatim ORBextractor.cc + ORBmatcher.cc, g++ -O3 -march=x86-64-v3 -ffp-contract=off, scalar OpenCV shim"
    if O.ref() is not None and O.refmatch() is not None:
        return (lambda nf: O.RefExtractor(nf, 1.2, 8, 15, 5)), O.RefMatch(O.refmatch()), "reference", \
            "verbatim ORBextractor.cc + ORBmatcher.cc, g++ -O2 -ffp-contract=off, scalar OpenCV shim"

    class _Port:
        search_for_initialization = staticmethod(O.search_for_initialization)
        birdview_match = staticmethod(O.birdview_match)
        grid_assign = staticmethod(O.grid_assign)
    return (lambda nf: O.OracleExtractor(nf, 1.2, 8, 15, 5)), _Port, "port", "oracle restatement, g++ -O2"


def _cpu_worker(args):
    """extract front+bird for a run of consecutive pairs and match each against its predecessor, on one host thread"""
    from fishbirdeyevisualslam_b200.matcher import Frame
    fronts, birds = args
    mk, M, _, _ = _cpu_libs()
    ef, eb = mk(FRONT_FEATURES), mk(BIRD_FEATURES)
    prev = None
    nm = 0
    for f, b in zip(fronts, birds):
        kf, df = ef(f)
        kb, db = eb(b)
        F, Bf = Frame.front(kf, df, FRONT[1], FRONT[0]), Frame.bird(kb, db, BIRD[1], BIRD[0])
        M.grid_assign(kf, F.min_x, F.min_y, F.inv_w, F.inv_h, 64, 48)          # Frame::AssignFeaturesToGrid (also runs inside the
        M.grid_assign(kb, 0.0, 0.0, Bf.inv_w, Bf.inv_h, 32, 32)                # matcher wrapper's Frame construction)
        if prev is not None:
            pm = np.ascontiguousarray(np.stack([prev[0].kps["x"], prev[0].kps["y"]], 1), np.float32)
            nm += M.search_for_initialization(prev[0], F, pm, 100, 0.9, True)[0]
            nm += M.birdview_match(prev[1].kps, prev[1].desc, Bf, 10, 0.9, True)[0]
        prev = (F, Bf)
    return nm


def cpu_reference_rate(fr, bi, pairs_per_thread: int, threads: int):
    """-> pairs/s of `threads` host threads each working through `pairs_per_thread` consecutive pairs of (fr, bi).
    ctypes releases the GIL, so plain threads use all cores."""
    from concurrent.futures import ThreadPoolExecutor
    n = len(fr)
    jobs = []
    for t in range(threads):
        idx = [(t * pairs_per_thread + j) % n for j in range(pairs_per_thread)]
        jobs.append((fr[idx], bi[idx]))
    _cpu_libs()
    t0 = time.perf_counter()
    if threads == 1:
        _cpu_worker(jobs[0])
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(_cpu_worker, jobs))
    return pairs_per_thread * threads / (time.perf_counter() - t0)


def cv2_primitives_ms(front: np.ndarray, bird: np.ndarray):
    """Second data point: the OpenCV primitives the reference calls, through SIMD-tuned cv2 on ONE thread --
    ComputePyramid (resize + copyMakeBorder), the per-cell FAST calls with threshold fallback, GaussianBlur -- for one
    front+bird pair.  The shim primitives of the verbatim build are scalar; this bounds what a tuned OpenCV build saves."""
    try:
        import cv2
    except Exception:
        return None
    cv2.setNumThreads(1)
    ini, mn = cv2.FastFeatureDetector_create(15, True), cv2.FastFeatureDetector_create(5, True)
    out = {"pyramid": 0.0, "fast_per_cell": 0.0, "blur": 0.0}
    for img in (front, bird):
        t0 = time.perf_counter()
        levels = []
        h0, w0 = img.shape
        sc = 1.0
        for l in range(8):
            w, h = int(round(w0 / sc)), int(round(h0 / sc))
            src = img if l == 0 else cv2.resize(levels[-1][19:-19, 19:-19], (w, h), interpolation=cv2.INTER_LINEAR)
            levels.append(cv2.copyMakeBorder(src, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
            sc *= 1.2
        t1 = time.perf_counter()
        for lv in levels:                                        # the reference's cell loop (src/ORBextractor.cc:765-829)
            H, W = lv.shape[0] - 38, lv.shape[1] - 38
            minb, maxbx, maxby = 16, W - 16, H - 16
            ncols, nrows = (maxbx - minb) // 30, (maxby - minb) // 30
            wc, hc = -(-(maxbx - minb) // ncols), -(-(maxby - minb) // nrows)
            for i in range(nrows):
                iy = minb + i * hc
                if iy >= maxby - 3:
                    continue
                my = min(iy + hc + 6, maxby)
                for j in range(ncols):
                    ix = minb + j * wc
                    if ix >= maxbx - 6:
                        continue
                    mx = min(ix + wc + 6, maxbx)
                    cell = lv[19 + iy:19 + my, 19 + ix:19 + mx]
                    if not ini.detect(cell):
                        mn.detect(cell)
        t2 = time.perf_counter()
        for lv in levels:
            cv2.GaussianBlur(lv[19:-19, 19:-19], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        t3 = time.perf_counter()
        out["pyramid"] += 1e3 * (t1 - t0); out["fast_per_cell"] += 1e3 * (t2 - t1); out["blur"] += 1e3 * (t3 - t2)
    out["total_ms_per_pair"] = sum(out.values())
    out["note"] = "cv2 %s, setNumThreads(1); per-cell FAST includes the Python call overhead of ~3000 cv2 calls per pair" % cv2.__version__
    return out


def cpu_baseline_block(fr, bi):
    """cpu_baseline of the 1-GPU line: all host threads (the headline `value`), one core, and the cv2-primitive data point."""
    threads = os.cpu_count() or 1
    _, _, kind, build = _cpu_libs()
    one = cpu_reference_rate(fr, bi, 6, 1)                       # ~1.6 core-seconds
    per_thread = 6
    allc = cpu_reference_rate(fr, bi, per_thread, threads)       # ~1.6 s of wall time, ~1.6 x threads core-seconds
    prim = cv2_primitives_ms(fr[0], bi[0])
    return {"value": allc, "unit": "pairs/s", "cores": threads, "kind": kind, "build": build,
            "sample": f"{per_thread * threads} pairs of rank 0's batch 0 ({per_thread} consecutive pairs on each of {threads} host threads)",
            "one_core": {"value": one, "unit": "pairs/s", "cores": 1, "sample": "6 consecutive pairs on one thread"},
            "all_cores": {"value": allc, "unit": "pairs/s", "cores": threads},
            "cv2_primitives_one_core": prim}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from fishbirdeyevisualslam_b200 import synth
    threads = os.cpu_count() or 1
    per_thread = 4
    sf, sb = batch_seeds(0, 0)
    fr, bi = synth.cheap_batch(args.batch, FRONT[0], FRONT[1], sf), synth.cheap_batch(args.batch, BIRD[0], BIRD[1], sb)
    _, _, kind, build = _cpu_libs()
    rates = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_rate(fr, bi, per_thread, threads)
        if i >= args.warmup:
            rates.append(r)
    v = statistics.median(rates)
    sample = f"{per_thread * threads} pairs of rank 0's batch 0 per step ({per_thread} consecutive pairs on each of {threads} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * per_thread * threads / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args.gpus, args.batch, args.pairs),
            "note": "reference CPU code path on the host cores (not a GPU run); each step is a bounded sample of the workload",
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind, "build": build, "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def bind_to_gpu_numa(index):
    """Pin this rank's host threads to the CPU cores of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is
    allocated (first touch places the pages there), so that with 8 ranks the H2D copies of a rank do not cross the socket
    interconnect.  Returns the node number or None when the topology is not exposed (then nothing is changed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------- our arm
def parity_check(pipe_cls, hF_last, hB_last, feat, res, fm, bm, pairs):
    """In-run parity: the e2e loop's LAST step (what the public host-buffer call actually returned) against the CPU oracle for a
    few sampled pairs -- keypoints, descriptors, match lists, counts, byte for byte."""
    from fishbirdeyevisualslam_b200.matcher import Frame
    from oracle import oracle as O
    of, ob = O.OracleExtractor(FRONT_FEATURES, 1.2, 8, 15, 5), O.OracleExtractor(BIRD_FEATURES, 1.2, 8, 15, 5)
    fk, fd, bk, bd = feat
    ok = True
    detail = []
    crc = 0
    for p in pairs:
        frames = []
        for q in (p - 1, p):
            kf, df = of(hF_last[q])
            kb, db = ob(hB_last[q])
            nf, nb = int(res["n_front"][q]), int(res["n_bird"][q])
            same = (nf == len(kf) and nb == len(kb) and fk[q][:nf].tobytes() == kf.tobytes() and np.array_equal(fd[q][:nf], df)
                    and bk[q][:nb].tobytes() == kb.tobytes() and np.array_equal(bd[q][:nb], db))
            ok &= bool(same)
            crc = zlib.crc32(fk[q][:nf].tobytes() + fd[q][:nf].tobytes() + bk[q][:nb].tobytes() + bd[q][:nb].tobytes(), crc)
            frames.append((Frame.front(kf, df, FRONT[1], FRONT[0]), Frame.bird(kb, db, BIRD[1], BIRD[0])))
        (F0, B0), (F1, B1) = frames
        pm = np.ascontiguousarray(np.stack([F0.kps["x"], F0.kps["y"]], 1), np.float32)
        n_o, m_o = O.search_for_initialization(F0, F1, pm, 100, 0.9, True)
        nb_o, d_o = O.birdview_match(B0.kps, B0.desc, B1, 10, 0.9, True)
        gm = bm[p][:B0.N]
        same = (int(res["front_matches"][p]) == n_o and np.array_equal(fm[p][:F0.N], m_o) and int(res["bird_matches"][p]) == nb_o
                and np.array_equal(np.stack([np.nonzero(gm > 0)[0], gm[gm > 0]], 1), d_o[:, :2]))
        ok &= bool(same)
        crc = zlib.crc32(fm[p][:F0.N].tobytes() + gm.tobytes(), crc)
        detail.append({"pair": int(p), "front_kps": int(F1.N), "bird_kps": int(B1.N), "front_matches": int(n_o), "bird_matches": int(nb_o)})
    return {"parity": bool(ok), "pairs_checked": detail, "crc32_of_checked_outputs": crc & 0xFFFFFFFF,
            "against": "oracle (CPU restatement pinned to the verbatim reference build), on the outputs the last e2e step returned"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from fishbirdeyevisualslam_b200 import _lib, shard
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline, PinnedBuffer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    numa_node = bind_to_gpu_numa(local) if world > 1 else None
    torch.cuda.set_device(local)
    # NCCL prints its version banner on stdout; keep stdout for the single JSON line (everything else goes to stderr)
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.load()
    B = args.batch
    nbatch = max(1, args.pairs // B)
    pairs = nbatch * B
    t_setup = time.perf_counter()
    pipe = FrontBirdPipeline(B, FRONT, BIRD, FRONT_FEATURES, BIRD_FEATURES, device=local)
    # this rank's P pairs: pinned host copy (e2e inputs) + device-resident copy (`value` inputs)
    hF, hB = PinnedBuffer((nbatch, B) + FRONT), PinnedBuffer((nbatch, B) + BIRD)
    synth_batches(rank, nbatch, B, hF.array, hB.array)
    dF = torch.empty((nbatch, B) + FRONT, dtype=torch.uint8, device="cuda")
    dB = torch.empty((nbatch, B) + BIRD, dtype=torch.uint8, device="cuda")
    dF.copy_(torch.from_numpy(hF.array), non_blocking=True)
    dB.copy_(torch.from_numpy(hB.array), non_blocking=True)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    stream = torch.cuda.ExternalStream(pipe.stream_ptr, device=local)
    fstride, bstride = B * FRONT[0] * FRONT[1], B * BIRD[0] * BIRD[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput (`value`) ------------------------------------------------------------------
    W = max(args.warmup, 3)
    for i in range(W):
        pipe.step_dev(dF.data_ptr() + (i % nbatch) * fstride, dB.data_ptr() + (i % nbatch) * bstride)
    pipe.sync()
    pipe._L.fbe_pipeline_stage_timing(pipe._h, 1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.fbe_kernel_launch_count()
    clk = ClockSampler(local)
    clk.__enter__()                      # samples through ALL timed regions (device-resident loop and the e2e loops)
    e0.record(stream)
    for i in range(args.steps):
        k = (W + i) % nbatch
        pipe.step_dev(dF.data_ptr() + k * fstride, dB.data_ptr() + k * bstride)
    pipe.join()               # a step ends on the pipeline's matching stream: order the public stream after it
    e1.record(stream)
    pipe.sync()
    barrier()
    launches = L.fbe_kernel_launch_count() - launches0
    import ctypes as C
    stage = (C.c_double * 12)()
    nst = C.c_int32()
    pipe._L.fbe_pipeline_stage_ms(pipe._h, stage, C.byref(nst))
    pipe._L.fbe_pipeline_stage_timing(pipe._h, 0)
    res_dev, _, _ = pipe.fetch(with_matches=False)
    ms = allmax(e0.elapsed_time(e1))
    value = world * B * args.steps / (ms * 1e-3)

    # ---- the north_star collective: NCCL all-gather of the last step's match records (outside `value`) -----------
    rec_bytes = shard.match_record_bytes(pipe.front_cap, pipe.bird_cap)
    shard.gather_matches(pipe, world)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    g0.record()
    NG = 10
    for _ in range(NG):
        gres, gfm, gbm = shard.gather_matches(pipe, world)
    g1.record()
    torch.cuda.synchronize()
    gather_ms = allmax(g0.elapsed_time(g1) / NG)

    # ---- end to end through host buffers (`e2e`) ----------------------------------------------------------------
    # The public host-buffer API: submit(step i) enqueues H2D of the inputs, the step and D2H of the results; wait(i-2)
    # collects an earlier step.  Inputs and results live in pinned host memory; every step's H2D + D2H is inside the
    # timed region (three steps in flight, so the copies of steps i+1, i+2 travel while step i computes).
    NBUF = 3
    hres = [PinnedBuffer((B,), _lib.PAIR_RESULT_DTYPE) for _ in range(NBUF)]
    hfm = [PinnedBuffer((B, pipe.front_cap), np.int32) for _ in range(NBUF)]
    hbm = [PinnedBuffer((B, pipe.bird_cap), np.int32) for _ in range(NBUF)]
    hfeat = [(PinnedBuffer((B, pipe.front_cap), _lib.KP_DTYPE), PinnedBuffer((B, pipe.front_cap, 32), np.uint8),
              PinnedBuffer((B, pipe.bird_cap), _lib.KP_DTYPE), PinnedBuffer((B, pipe.bird_cap, 32), np.uint8)) for _ in range(NBUF)]

    def e2e_loop(n, first, with_features):
        q = []
        for i in range(n):
            j = i % NBUF
            k = (first + i) % nbatch
            feats = tuple(b.array for b in hfeat[j]) if with_features else None
            q.append(pipe.submit_host(hF.ptr + k * fstride, hB.ptr + k * bstride, hres[j].array, hfm[j].array, hbm[j].array, features=feats))
            if len(q) >= NBUF:
                pipe.wait(q.pop(0))           # the results of step i - 2 are now in the host buffers
        for t in q:
            pipe.wait(t)

    def timed_e2e(with_features):
        e2e_loop(4, 0, with_features)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(args.steps, 4, with_features)
        torch.cuda.synchronize()
        return world * B * args.steps / allmax(time.perf_counter() - t0)

    e2e_matches_only = timed_e2e(False)
    e2e_value = timed_e2e(True)
    clk.__exit__(None, None, None)
    h2d = int(fstride + bstride)
    d2h_matches = int(hres[0].nbytes + hfm[0].nbytes + hbm[0].nbytes)
    d2h_feat = int(sum(b.nbytes for b in hfeat[0]))
    # what the host link gives a plain pinned->device copy of the same bytes (the ceiling of any host-buffer path), every rank
    # copying at the same time
    hp = torch.empty(h2d, dtype=torch.uint8, pin_memory=True)
    dp = torch.empty(h2d, dtype=torch.uint8, device="cuda")
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dp.copy_(hp, non_blocking=True)
    barrier()
    c0.record()
    for _ in range(5):
        dp.copy_(hp, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    link_gbs = 5 * h2d / (c0.elapsed_time(c1) * 1e-3) / 1e9
    link_all = [link_gbs]
    if world > 1:
        lt = torch.tensor([link_gbs], dtype=torch.float64, device="cuda")
        lall = [torch.zeros_like(lt) for _ in range(world)]
        dist.all_gather(lall, lt)
        link_all = [float(x.item()) for x in lall]
    del hp, dp

    # ---- checks (outside every timed region) -----------------------------------------------------------------------
    # (1) parity of what the last e2e step returned, against the CPU oracle (rank 0)
    last_i = args.steps - 1
    last_j, last_k = last_i % NBUF, (4 + last_i) % nbatch
    check = {}
    if rank == 0:
        lastres = hres[last_j].array
        check = {"mean_front_kps": float(lastres["n_front"].mean()), "mean_bird_kps": float(lastres["n_bird"].mean()),
                 "mean_front_matches": float(lastres["front_matches"][1:].mean()), "mean_bird_matches": float(lastres["bird_matches"][1:].mean())}
        sample = sorted({1, B // 3, (2 * B) // 3, B - 1} - {0})
        check.update(parity_check(FrontBirdPipeline, hF.array[last_k], hB.array[last_k], tuple(b.array for b in hfeat[last_j]),
                                  lastres, hfm[last_j].array, hbm[last_j].array, sample))
    # (2) sharded == single: every rank runs its batches 0 and 1 on a FRESH pipeline and the records are all-gathered on the
    #     device; rank 0 then re-computes the LAST rank's two batches itself (regenerated from their seeds) and compares bytes
    vp = FrontBirdPipeline(B, FRONT, BIRD, FRONT_FEATURES, BIRD_FEATURES, device=local)
    gathered = []
    for k in range(min(2, nbatch)):
        vp.step_dev(dF.data_ptr() + k * fstride, dB.data_ptr() + k * bstride)
        gathered.append(tuple(t.cpu().numpy() for t in shard.gather_matches(vp, world)))
    vp.close()
    if rank == 0:
        src = world - 1
        if src == 0:
            rf, rb = hF.array[:2], hB.array[:2]
        else:
            rf = np.empty((min(2, nbatch), B) + FRONT, np.uint8)
            rb = np.empty((min(2, nbatch), B) + BIRD, np.uint8)
            synth_batches(src, min(2, nbatch), B, rf, rb)
        sp = FrontBirdPipeline(B, FRONT, BIRD, FRONT_FEATURES, BIRD_FEATURES, device=local)
        same = True
        for k in range(min(2, nbatch)):
            tF, tB = torch.from_numpy(rf[k]).cuda(), torch.from_numpy(rb[k]).cuda()
            sp.step_dev(tF.data_ptr(), tB.data_ptr())
            r1, f1, b1 = sp.fetch()
            g = gathered[k]
            same &= g[0][src].tobytes() == r1.tobytes()
            for p in range(B):
                if k == 0 and p == 0:
                    continue                                    # the first pair of a run has no predecessor: its lists are unspecified
                nq = int(r1["n_front"][p - 1]) if p else int(gathered[k - 1][0][src][B - 1][0])
                nb = int(r1["n_bird"][p - 1]) if p else int(gathered[k - 1][0][src][B - 1][1])
                same &= bool(np.array_equal(g[1][src][p][:nq], f1[p][:nq]) and np.array_equal(g[2][src][p][:nb], b1[p][:nb]))
        sp.close()
        check["sharded_equals_single"] = bool(same)
        check["sharded_equals_single_what"] = (f"rank {src}'s first {min(2, nbatch) * B} pairs: NCCL-gathered device records vs rank 0 recomputing them "
                                               "alone on a fresh pipeline (counts + front / bird match lists, bytes)")
    barrier()

    if rank == 0:
        peak, peak_src = peaks()
        n = max(nst.value, 1)
        fast_ms = stage[1] / n
        # dominant kernel = FAST cells of the front extractor: one launch reads every pyramid pixel of B images once
        alg_bytes = B * PYR_PIXELS_FRONT
        achieved = alg_bytes / (fast_ms * 1e-3) / 1e9 if fast_ms > 0 else 0.0
        if world == 1:
            cpu = cpu_baseline_block(hF.array[0], hB.array[0])
        else:                               # the CPU baseline is a property of the host, measured once: on the 1-GPU line
            cpu = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference", "sample": "measured at N=1 only (see the 1-GPU line / --impl reference)"}
        sm_mhz = clk.summary().get("sm_mhz") or 1965.0
        issue_peak = 148 * 4 * sm_mhz * 1e6
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": config_dict(world, B, pairs),
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_matches + d2h_feat,
                        "returns": "per frame keypoints + descriptors (what ORBextractor::operator() returns), per pair match index lists + counts",
                        "with_features": {"value": e2e_value, "d2h_bytes_per_step": d2h_matches + d2h_feat},
                        "matches_only": {"value": e2e_matches_only, "d2h_bytes_per_step": d2h_matches},
                        "h2d_link_gbs_measured": link_gbs, "h2d_link_gbs_measured_per_rank": link_all,
                        "h2d_gbs_used": e2e_value / world * (h2d / B) / 1e9, "host_numa_node_rank0": numa_node,
                        "note": "three steps in flight: the input copies of the next steps overlap the kernels of step i; bound = max(copy, compute)"},
                "gpu_launches": int(launches),
                "clocks": clk.summary(),
                "roofline": {"bound": "hbm", "kernel": "k_fast_cells (front)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": FAST_DRAM_BYTES_PER_FRONT_IMAGE * B, "peak_source": peak_src,
                             "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of k_fast_cells (front), per image x images per launch; " + FAST_PROFILE,
                             "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": fast_ms,
                             "note": "k_fast_cells is bound by the integer ALU pipe, not by memory: ncu (" + FAST_PROFILE + ") shows 81 % issue-slot "
                                     "utilisation, 74 % alu pipe (VIMNMX3 / PRMT run at half rate) and 4 % of DRAM throughput; ~90 integer "
                                     "instructions per pixel bound it, so frac of the HBM peak stays small by construction",
                             "issue": {"what": "the roof this kernel actually sits under: warp-instruction issue (148 SMs x 4 schedulers x SM clock)",
                                       "achieved_gwarp_inst_s": FAST_WARP_INST_PER_FRONT_IMAGE * B / (fast_ms * 1e-3) / 1e9 if fast_ms > 0 else 0.0,
                                       "peak_gwarp_inst_s": issue_peak / 1e9,
                                       "frac": FAST_WARP_INST_PER_FRONT_IMAGE * B / (fast_ms * 1e-3) / issue_peak if fast_ms > 0 else 0.0,
                                       "source": "smsp__inst_executed.sum of one launch (ncu --set full, " + FAST_PROFILE + ") / live launch time"},
                             "whole_step": {"bytes_per_pair": BYTES_PER_PAIR, "achieved_gbs": value / world * BYTES_PER_PAIR / 1e9,
                                            "frac": value / world * BYTES_PER_PAIR / 1e9 / peak,
                                            "issue_frac": value / world * STEP_WARP_INST_PER_PAIR / issue_peak,
                                            "issue_note": "warp-instructions of every kernel of a step (profiles/r2_29_stage_ncu.md) x pairs/s over "
                                                          "the warp-instruction issue peak: the step as a whole is instruction-bound"}},
                "stage_ms": {"front_fast_cells": fast_ms,
                             "note": "CUDA events around the front FAST launch on its own (highest-priority) stream: agrees with the ncu launch "
                                     "time. The other stages overlap across four streams, so their event spans are neither additive nor kernel "
                                     "times and are not reported (see the ncu launch list under profiles/)"},
                "gather": {"what": "NCCL all_gather_into_tensor of the last step's fixed-stride match records, device-resident (no host staging); outside `value`",
                           "ms": gather_ms, "record_bytes_per_pair": rec_bytes, "bytes_per_rank": rec_bytes * B,
                           "algbw_gbs": rec_bytes * B * world / (gather_ms * 1e-3) / 1e9 if gather_ms > 0 else None, "ranks": world},
                "cpu_baseline": cpu,
                "setup_s": setup_s,
                "check": check}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=32)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="frame pairs per step per GPU")
    ap.add_argument("--pairs", type=int, default=4096, help="distinct frame pairs per GPU (the offline job of config C4)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
