"""Host link microbenchmark behind the 4-8 GPU end-to-end numbers: what a pinned host->device copy of one
bench step's input (137 MB) sustains per rank when 1 .. N ranks copy at once, with one stream or two streams of 32 MB chunks,
and with the step's device->host result copy (25.6 MB) running in the other direction.  No kernels, no collective in the timed
region (barriers only line the ranks up).  Launch like the bench:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_microbench.py
Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

H2D_BYTES = 136_839_168          # bench.py e2e.h2d_bytes_per_step
D2H_BYTES = 25_626_624           # bench.py e2e.d2h_bytes_per_step (with features)
CHUNK = 32 << 20
REPS = 20


def timed(fn, streams):
    for s in streams:
        s.synchronize()
    dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]:
        s.wait_event(e0)
    for _ in range(REPS):
        fn()
    for s in streams[1:]:
        ev = torch.cuda.Event()
        ev.record(s)
        streams[0].wait_event(ev)
    e1.record(streams[0])
    e1.synchronize()
    return e0.elapsed_time(e1) / REPS


def main():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h_in = torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory()
    h_in.fill_(rank + 1)                                  # first touch by the rank that copies
    h_out = torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory()
    h_out.zero_()
    d_in = torch.empty(H2D_BYTES, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(D2H_BYTES, dtype=torch.uint8, device="cuda")
    s0, s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

    def one_stream():
        with torch.cuda.stream(s0):
            d_in.copy_(h_in, non_blocking=True)

    def two_streams_chunked():
        for k, off in enumerate(range(0, H2D_BYTES, CHUNK)):
            with torch.cuda.stream(s0 if k % 2 == 0 else s1):
                d_in[off:off + CHUNK].copy_(h_in[off:off + CHUNK], non_blocking=True)

    def both_directions():
        with torch.cuda.stream(s0):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    res = {}
    for _ in range(3):
        one_stream()
    torch.cuda.synchronize()
    # all ranks at once
    for name, fn, streams in (("h2d_one_stream", one_stream, [s0]), ("h2d_two_streams_32MB_chunks", two_streams_chunked, [s0, s1]),
                              ("h2d_with_d2h_opposite", both_directions, [s0, s2])):
        ms = timed(fn, streams)
        res[name] = H2D_BYTES / ms / 1e6                   # GB/s of the H2D direction
    # one rank at a time (the others idle at the barrier inside timed())
    alone = 0.0
    for r in range(world):
        if r == rank:
            ms = timed(one_stream, [s0])
            alone = H2D_BYTES / ms / 1e6
        else:
            dist.barrier()
    res["h2d_alone"] = alone
    # half of the ranks at a time (which ranks share what: the two halves see different rates when all copy)
    if world >= 4:
        half = world // 2
        for name, active in (("h2d_low_half_only", rank < half), ("h2d_high_half_only", rank >= half)):
            if active:
                ms = timed(one_stream, [s0])
                res[name] = H2D_BYTES / ms / 1e6
            else:
                dist.barrier()
                res[name] = 0.0
        dist.barrier()
    keys = sorted(res)
    t = torch.tensor([res[k] for k in keys], dtype=torch.float64, device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        per = {k: [round(float(a[i]), 2) for a in allt] for i, k in enumerate(keys)}
        out = {"what": "pinned host<->device copy rate per rank, GB/s of the H2D direction; all ranks copying at once unless 'alone'",
               "ranks": world, "h2d_bytes": H2D_BYTES, "d2h_bytes": D2H_BYTES, "reps": REPS, "per_rank_gbs": per,
               "aggregate_gbs": {k: round(sum(v), 1) for k, v in per.items() if k != "h2d_alone"},
               "pairs_per_s_ceiling_slowest_rank": {k: round(min(v) * 1e9 / (H2D_BYTES / 128) * world) for k, v in per.items()
                                                    if k != "h2d_alone" and "half" not in k},
               "pairs_per_s_ceiling_if_balanced": {k: round(sum(v) * 1e9 / (H2D_BYTES / 128)) for k, v in per.items()
                                                   if k != "h2d_alone" and "half" not in k}}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
