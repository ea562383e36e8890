"""Host-side sharding logic (config C4) incl. a world_size-2 gloo run on CPU: block ranges, halo, record gather."""
import os
import socket

import numpy as np
import pytest

from fishbirdeyevisualslam_b200.shard import gather_records, halo_index, shard_range, step_plan


def test_ranges_cover_and_balance():
    for n in (1, 7, 8, 4096, 4097, 13):
        for world in (1, 2, 4, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(4096, 3, 8) == (1536, 2048)


def test_halo_and_step_plan():
    assert halo_index(0) is None and halo_index(512) == 511
    assert step_plan(0, 10, 4) == [(0, 4), (4, 4), (8, 2)]
    plan = step_plan(512, 1024, 128)
    assert plan[0] == (511, 128) and sum(n for _, n in plan) == 513 and plan[-1][0] + plan[-1][1] == 1024


def _worker(rank, world, port, n_pairs, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(n_pairs, rank, world)
    idx = np.arange(a, b)
    local = np.stack([idx, idx * 7 + 1, idx % 5, -idx], 1).astype(np.int32)     # stand-in for per-pair match records
    full = gather_records(local, n_pairs, rank, world)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [9, 16])
def test_gather_two_ranks_gloo(n_pairs):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(n_pairs)
    exp = np.stack([idx, idx * 7 + 1, idx % 5, -idx], 1).astype(np.int32)
    assert np.array_equal(got[0], exp) and np.array_equal(got[1], exp)


@pytest.mark.gpu
def test_sharded_run_equals_single_run():
    """The C4 contract at a small size: 2-way and 3-way sharded processing on one GPU == unsharded processing."""
    from fishbirdeyevisualslam_b200 import synth
    from fishbirdeyevisualslam_b200.shard import run_offline_batch
    n = 10
    fr = np.stack([synth.frame(720, 1280, 700, (min(3 * i, 8) - 4, min(2 * i, 8) - 4), noise_seed=i) for i in range(n)])
    bi = np.stack([synth.frame(384, 384, 701, (min(2 * i, 8) - 4, min(i, 8) - 4), noise_seed=50 + i) for i in range(n)])
    ref = run_offline_batch(fr, bi, 0, 1, 4)
    assert (ref[1:, 2] > 50).all()
    for world in (2, 3):
        parts = [run_offline_batch(fr, bi, r, world, 4) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), ref)
