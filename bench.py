#!/usr/bin/env python3
"""bench.py -- front+bird frame-pairs/sec, ORB extract + grid + match (BASELINE.json metric, config C2/C4 shape).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

ours      : the sm_100a pipeline through the C-ABI.  One step = one batch of B synthetic front(1280x720 @2000) +
            bird(384x384 @1000) pairs: extract both, build both grids, match every pair against the previous one
            (front: SearchForInitialization, window 100; bird: BirdviewMatch, window 10).  `value` = device-timed
            pairs/s with inputs resident in HBM; `e2e` = the same through HOST (pinned) buffers, H2D + D2H inside the
            timed region.  Under torchrun each rank owns B pairs per step (weak scaling, no data-path collective).
reference : the reference's own CPU code path on the host cores -- its ORBextractor.cc compiled verbatim
            (oracle/_ref/libfbe_ref.so) when that was built, else the oracle port -- plus the restated matchers,
            all host threads, a bounded sample per step.
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
# profiles/r1_10_fast_ncu.md): 384.76 MB + 35.90 MB
FAST_DRAM_BYTES_PER_FRONT_IMAGE = int((384.759040e6 + 35.900416e6) / 128)
# warp-instructions executed by the same launch (sm__inst_executed.sum of the same capture)
FAST_WARP_INST_PER_FRONT_IMAGE = 1_063_465_360 / 128
METRIC = "front+bird frame-pairs/sec ORB extract+match at 1/2/4/8 B200 vs host CPU ref"


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
def _cpu_worker(args):
    """extract front+bird for a run of consecutive pairs and match each against its predecessor, on one host thread"""
    from fishbirdeyevisualslam_b200.matcher import Frame
    from oracle import oracle as O
    fronts, birds, use_ref = args
    mk = (lambda nf: O.RefExtractor(nf, 1.2, 8, 15, 5)) if use_ref else (lambda nf: O.OracleExtractor(nf, 1.2, 8, 15, 5))
    ef, eb = mk(FRONT_FEATURES), mk(BIRD_FEATURES)
    prev = None
    nm = 0
    for f, b in zip(fronts, birds):
        kf, df = ef(f)
        kb, db = eb(b)
        F, Bf = Frame.front(kf, df, FRONT[1], FRONT[0]), Frame.bird(kb, db, BIRD[1], BIRD[0])
        O.grid_assign(kf, F.min_x, F.min_y, F.inv_w, F.inv_h, 64, 48)
        O.grid_assign(kb, 0.0, 0.0, Bf.inv_w, Bf.inv_h, 32, 32)
        if prev is not None:
            pm = np.ascontiguousarray(np.stack([prev[0].kps["x"], prev[0].kps["y"]], 1), np.float32)
            nm += O.search_for_initialization(prev[0], F, pm, 100, 0.9, True)[0]
            nm += O.birdview_match(prev[1].kps, prev[1].desc, Bf, 10, 0.9, True)[0]
        prev = (F, Bf)
    return nm


def cpu_reference_rate(pairs_per_thread: int, threads: int, repeats: int = 1):
    """-> (pairs/s, kind, cores).  ctypes releases the GIL, so plain threads use all cores."""
    from concurrent.futures import ThreadPoolExecutor
    from fishbirdeyevisualslam_b200 import synth
    from oracle import oracle as O
    O.lib()
    use_ref = O.ref() is not None
    fr = synth.cheap_batch(pairs_per_thread * threads, FRONT[0], FRONT[1], 11)
    bi = synth.cheap_batch(pairs_per_thread * threads, BIRD[0], BIRD[1], 12)
    jobs = [(fr[t * pairs_per_thread:(t + 1) * pairs_per_thread], bi[t * pairs_per_thread:(t + 1) * pairs_per_thread], use_ref)
            for t in range(threads)]
    best = None
    with ThreadPoolExecutor(threads) as ex:
        for _ in range(repeats):
            t0 = time.perf_counter()
            list(ex.map(_cpu_worker, jobs))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return pairs_per_thread * threads / best, ("reference" if use_ref else "port"), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    per_thread = 4
    rates = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        r, kind, _ = cpu_reference_rate(per_thread, threads)
        if i >= args.warmup:
            rates.append(r)
    v = statistics.median(rates)
    sample = f"{per_thread * threads} pairs per step ({per_thread} consecutive pairs on each of {threads} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * per_thread * threads / v, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2/C4: 1280x720@2000 front + 384x384@1000 bird, extract + grid + frame-to-frame match",
                       "note": "reference CPU code path on host cores (not a GPU run)"},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample},
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    from fishbirdeyevisualslam_b200 import _lib, synth
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
    pipe = FrontBirdPipeline(B, FRONT, BIRD, FRONT_FEATURES, BIRD_FEATURES, device=local)
    fr = synth.cheap_batch(B, FRONT[0], FRONT[1], 100 + rank)
    bi = synth.cheap_batch(B, BIRD[0], BIRD[1], 200 + rank)
    dF, dB = torch.from_numpy(fr).cuda(), torch.from_numpy(bi).cuda()
    stream = torch.cuda.ExternalStream(pipe.stream_ptr, device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) ------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        pipe.step_dev(dF.data_ptr(), dB.data_ptr())
    pipe.sync()
    pipe._L.fbe_pipeline_stage_timing(pipe._h, 1)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = L.fbe_kernel_launch_count()
    clk = ClockSampler(local)
    clk.__enter__()                      # samples through BOTH timed regions (device-resident loop and e2e loop)
    if True:
        e0.record(stream)
        for _ in range(args.steps):
            pipe.step_dev(dF.data_ptr(), dB.data_ptr())
        pipe.join()           # a step ends on the pipeline's matching stream: order the public stream after it
        e1.record(stream)
        pipe.sync()
        barrier()
    launches = L.fbe_kernel_launch_count() - launches0
    ms = e0.elapsed_time(e1)
    import ctypes as C
    stage = (C.c_double * 12)()
    nst = C.c_int32()
    pipe._L.fbe_pipeline_stage_ms(pipe._h, stage, C.byref(nst))
    pipe._L.fbe_pipeline_stage_timing(pipe._h, 0)
    res, _, _ = pipe.fetch(with_matches=False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end through host buffers (`e2e`) ----------------------------------------------------------------
    # The public host-buffer API: submit(step i) enqueues H2D of the inputs, the step and D2H of the results; wait(i-1)
    # collects the previous step.  Inputs and results live in pinned host memory; every step's H2D + D2H is inside the
    # timed region (three steps in flight, so the copies of steps i+1, i+2 travel while step i computes).
    NBUF = 3
    hF, hB = PinnedBuffer(fr.shape), PinnedBuffer(bi.shape)
    hF.array[...] = fr
    hB.array[...] = bi
    hres = [PinnedBuffer((B,), _lib.PAIR_RESULT_DTYPE) for _ in range(NBUF)]
    hfm = [PinnedBuffer((B, pipe.front_cap), np.int32) for _ in range(NBUF)]
    hbm = [PinnedBuffer((B, pipe.bird_cap), np.int32) for _ in range(NBUF)]

    def e2e_loop(n):
        q = []
        for i in range(n):
            k = i % NBUF
            q.append(pipe.submit_host(hF.ptr, hB.ptr, hres[k].array, hfm[k].array, hbm[k].array))
            if len(q) >= NBUF:
                pipe.wait(q.pop(0))           # the results of step i - 2 are now in hres / hfm / hbm [k']
        for t in q:
            pipe.wait(t)

    e2e_loop(4)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clk.__exit__(None, None, None)
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    h2d = int(fr.nbytes + bi.nbytes)
    d2h = int(hres[0].nbytes + hfm[0].nbytes + hbm[0].nbytes)
    # what the host link gives a plain pinned->device copy of the same bytes (the ceiling of any host-buffer path)
    hp = torch.empty(h2d, dtype=torch.uint8, pin_memory=True)
    dp = torch.empty(h2d, dtype=torch.uint8, device="cuda")
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dp.copy_(hp, non_blocking=True)
    torch.cuda.synchronize()
    c0.record()
    for _ in range(5):
        dp.copy_(hp, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    link_gbs = 5 * h2d / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del hp, dp

    if rank == 0:
        peak, peak_src = peaks()
        n = max(nst.value, 1)
        names = ["pyramid", "fast_cells", "octree", "describe", "grid", "blur"]
        stage_ms = {("front_" + names[i]): stage[i] / n for i in range(6)}
        stage_ms.update({("bird_" + names[i]): stage[6 + i] / n for i in range(6)})
        # dominant kernel = FAST cells of the front extractor: one launch reads every pyramid pixel of B images once
        fast_ms = stage_ms["front_fast_cells"]
        alg_bytes = B * PYR_PIXELS_FRONT
        achieved = alg_bytes / (fast_ms * 1e-3) / 1e9 if fast_ms > 0 else 0.0
        # CPU baseline beside it (bounded sample: ~2 pairs per host thread)
        threads = os.cpu_count() or 1
        CPU_PAIRS_PER_THREAD = 8            # ~20 core-seconds of the reference CPU path, ~1.3 s of wall time on 16 threads
        if world == 1:
            cpu_rate, kind, cores = cpu_reference_rate(CPU_PAIRS_PER_THREAD, threads)
            cpu_sample = f"{CPU_PAIRS_PER_THREAD * threads} pairs ({CPU_PAIRS_PER_THREAD} consecutive pairs on each of {threads} host threads)"
        else:                               # the CPU baseline is a property of the host, measured once: on the 1-GPU line
            cpu_rate, kind, cores, cpu_sample = None, "reference", 0, "measured at N=1 only (see the 1-GPU line / --impl reference)"
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": "C2/C4: 1280x720@2000 front + 384x384@1000 bird, extract + grid + frame-to-frame match",
                           "pairs_per_step_per_gpu": B, "parallelism": f"frames sharded over {world} GPU(s), no collective",
                           "l2": f"inputs {h2d / 1e6:.0f} MB per step per GPU (> 126 MB L2)" if h2d > 126e6 else "inputs smaller than L2: raise --batch"},
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "h2d_link_gbs_measured": link_gbs, "h2d_gbs_used": e2e_value / world * (h2d / B) / 1e9,
                        "host_numa_node_rank0": numa_node,
                        "note": "three steps in flight: the input copies of the next steps overlap the kernels of step i; bound = max(copy, compute)"},
                "gpu_launches": int(launches),
                "clocks": clk.summary(),
                "roofline": {"bound": "hbm", "kernel": "k_fast_cells (front)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": FAST_DRAM_BYTES_PER_FRONT_IMAGE * B, "peak_source": peak_src,
                             "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of k_fast_cells (front), per image x images per launch; profiles/r1_10_fast_ncu.md",
                             "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": fast_ms,
                             "note": "k_fast_cells is instruction-bound, not memory-bound: ncu (profiles/r1_10_fast_ncu.md) shows 77 % issue-slot "
                                     "utilisation, 72 % alu pipe, 3.0 warp-instructions per SM per clock and 4 % of DRAM throughput; ~90 integer "
                                     "instructions per pixel bound it, so frac of the HBM peak stays small by construction",
                             "issue": {"what": "the roof this kernel actually sits under: warp-instruction issue (148 SMs x 4 schedulers x SM clock)",
                                       "achieved_gwarp_inst_s": FAST_WARP_INST_PER_FRONT_IMAGE * B / (fast_ms * 1e-3) / 1e9,
                                       "peak_gwarp_inst_s": 148 * 4 * (clk.summary().get("sm_mhz") or 1965.0) * 1e6 / 1e9,
                                       "frac": FAST_WARP_INST_PER_FRONT_IMAGE * B / (fast_ms * 1e-3) / (148 * 4 * (clk.summary().get("sm_mhz") or 1965.0) * 1e6),
                                       "source": "sm__inst_executed.sum of one launch (ncu --set full, profiles/r1_10_fast_ncu.md) / live launch time"},
                             "whole_step": {"bytes_per_pair": BYTES_PER_PAIR, "achieved_gbs": value / world * BYTES_PER_PAIR / 1e9,
                                            "frac": value / world * BYTES_PER_PAIR / 1e9 / peak}},
                "stage_ms": stage_ms,
                "cpu_baseline": {"value": cpu_rate, "unit": "pairs/s", "cores": cores, "kind": kind,
                                 "sample": cpu_sample},
                "check": {"mean_front_kps": float(res["n_front"].mean()), "mean_bird_kps": float(res["n_bird"].mean()),
                          "mean_front_matches": float(res["front_matches"].mean()), "mean_bird_matches": float(res["bird_matches"].mean())}}
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128, help="frame pairs per step per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
