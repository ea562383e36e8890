"""Host-side sharding logic (config C4) incl. a world_size-2 gloo run on CPU: block ranges, halo, record gather."""
import os
import socket

import numpy as np
import pytest

from fishbirdeyevisualslam_b200.shard import gather_records, halo_index, match_record_bytes, shard_range, step_plan


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


def _worker_fixed(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from fishbirdeyevisualslam_b200.shard import gather_fixed_records
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, fcap, bcap = 5, 11, 7
    g = torch.Generator().manual_seed(100 + rank)
    parts = (torch.randint(0, 3000, (B, 4), generator=g, dtype=torch.int32), torch.randint(-1, 2000, (B, fcap), generator=g, dtype=torch.int32),
             torch.randint(-1, 1000, (B, bcap), generator=g, dtype=torch.int32))
    outs = gather_fixed_records(parts, world)
    q.put((rank, [p.numpy() for p in parts], [o.numpy() for o in outs]))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_match_lists_two_ranks_gloo():
    """The fixed-stride {counts, front idx[], bird idx[]} records of two shards, gathered: every rank holds both, in rank order."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_fixed, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {r: (parts, outs) for r, parts, outs in (q.get(timeout=120) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in range(2):
        for k in range(3):
            assert got[r][1][k].shape[0] == 2
            for src in range(2):
                assert np.array_equal(got[r][1][k][src], got[src][0][k])
    assert match_record_bytes(2064, 1064) == 16 + 4 * 2064 + 4 * 1064 <= 16 * 1024


@pytest.mark.gpu
def test_gather_matches_single_gpu_aliases_device_records():
    """gather_matches with world 1: the device tensors alias the pipeline's records (no host staging) and equal fetch()."""
    import torch
    from fishbirdeyevisualslam_b200 import synth
    from fishbirdeyevisualslam_b200.pipeline import FrontBirdPipeline
    from fishbirdeyevisualslam_b200.shard import gather_matches
    B = 4
    pipe = FrontBirdPipeline(B, (240, 320), (256, 256), 500, 300)
    fr = np.stack([synth.frame(240, 320, 5, (i - 4, i // 2 - 2), noise_seed=i) for i in range(2 * B)])       # one scene, drifting window
    bi = np.stack([synth.frame(256, 256, 6, (i - 4, i // 2 - 2), noise_seed=40 + i) for i in range(2 * B)])
    for k in range(2):
        dF, dB = torch.from_numpy(fr[k * B:(k + 1) * B]).cuda(), torch.from_numpy(bi[k * B:(k + 1) * B]).cuda()
        pipe.step_dev(dF.data_ptr(), dB.data_ptr())
        res, fm, bm = gather_matches(pipe, 1)
        hres, hfm, hbm = pipe.fetch()
        assert res.is_cuda and res.shape == (1, B, 4) and fm.shape == (1, B, pipe.front_cap)
        got = res[0].cpu().numpy()
        assert np.array_equal(got[:, 0], hres["n_front"]) and np.array_equal(got[:, 2], hres["front_matches"])
        assert np.array_equal(fm[0].cpu().numpy(), hfm) and np.array_equal(bm[0].cpu().numpy(), hbm)
    assert (hres["front_matches"] > 10).all()
    pipe.close()


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
