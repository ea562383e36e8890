"""Frame sharding for offline batches (config C4): frames are independent through extraction and grid assignment, and
frame-to-frame matching needs only the previous pair, so an offline batch of N pairs is cut into contiguous blocks, one
per GPU, with NO data-path collective.  The first pair of a block is matched against the last pair of the previous
block, which the block's owner re-extracts locally as a one-pair "halo" (cheaper than shipping 190 KB of keypoints).
The only collective is an optional all_gather of the fixed-stride match records at the end (north_star: "NCCL over
NVLink only to gather match results"), NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import numpy as np


def shard_range(n_pairs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of pair indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_pairs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def halo_index(start: int):
    """Index of the pair a block must additionally extract to serve as the match reference of its first pair."""
    return start - 1 if start > 0 else None


def step_plan(start: int, stop: int, batch: int):
    """[(first_pair, n_pairs)] steps of at most `batch` pairs covering the halo + the block, in order."""
    h = halo_index(start)
    first = start if h is None else h
    return [(s, min(batch, stop - s)) for s in range(first, stop, batch)]


def gather_records(local: np.ndarray, n_pairs: int, rank: int, world: int, backend_device: str = "cpu") -> np.ndarray:
    """all_gather of per-pair fixed-stride int32 records -> every rank holds the [n_pairs, stride] table."""
    import torch
    import torch.distributed as dist
    stride = local.shape[1]
    sizes = [shard_range(n_pairs, r, world)[1] - shard_range(n_pairs, r, world)[0] for r in range(world)]
    mx = max(sizes)
    pad = np.zeros((mx, stride), np.int32)
    pad[:len(local)] = local
    t = torch.from_numpy(pad).to(backend_device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:sizes[r]] for r, o in enumerate(outs)], axis=0)


def run_offline_batch(front: np.ndarray, bird: np.ndarray, rank: int, world: int, batch: int, device: int = 0, **pipe_kw) -> np.ndarray:
    """Process this rank's block of an offline sequence on one GPU; returns [n_local, 4] records
    (n_front, n_bird, front_matches, bird_matches), identical to what a single-GPU run yields for the same pairs."""
    import torch
    from .pipeline import FrontBirdPipeline
    n_pairs = len(front)
    start, stop = shard_range(n_pairs, rank, world)
    pipe = FrontBirdPipeline(batch, front.shape[1:], bird.shape[1:], device=device, **pipe_kw)
    rec = np.zeros((stop - start, 4), np.int32)
    for first, n in step_plan(start, stop, batch):
        f = np.zeros((batch,) + front.shape[1:], np.uint8)
        b = np.zeros((batch,) + bird.shape[1:], np.uint8)
        f[:n], b[:n] = front[first:first + n], bird[first:first + n]
        dF, dB = torch.from_numpy(f).cuda(device), torch.from_numpy(b).cuda(device)
        pipe.step_dev(dF.data_ptr(), dB.data_ptr())
        res, _, _ = pipe.fetch(with_matches=False)
        for p in range(n):
            g = first + p
            if g >= start:
                rec[g - start] = [res["n_front"][p], res["n_bird"][p], res["front_matches"][p], res["bird_matches"][p]]
    pipe.close()
    return rec           # the halo pair's own record belongs to the previous block: its result was skipped above (g < start)


class _DeviceView:
    """A raw device allocation of the library seen through __cuda_array_interface__, so that torch can alias it (zero copy)."""

    def __init__(self, ptr: int, shape, typestr: str = "<i4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def match_record_bytes(front_cap: int, bird_cap: int) -> int:
    """Bytes of one pair's fixed-stride match record: {n_front, n_bird, front_matches, bird_matches} + front idx[] + bird idx[]."""
    return 16 + 4 * front_cap + 4 * bird_cap


def gather_matches(pipe, world: int, group=None):
    """north_star's one collective: all-gather of the last step's match results of every shard, DEVICE-RESIDENT.

    The three record arrays of the pipeline (fbe_pair_result[B], front matches12 [B, front_cap], bird matches12
    [B, bird_cap]; fixed stride, -1 = unmatched) are aliased as CUDA tensors -- no host staging, no copy on the
    send side -- and gathered with NCCL (`all_gather_into_tensor`) on torch's current stream, ordered after the pipeline's
    streams; the pipeline's next step is ordered after the gather.  Returns (res [world, B, 4], front [world, B, front_cap],
    bird [world, B, bird_cap]) int32 CUDA tensors, identical on every rank."""
    import torch
    import torch.distributed as dist
    B = pipe.batch
    d_res, d_fm, d_bm = pipe.device_results()
    dev = torch.device("cuda", torch.cuda.current_device())
    res = torch.as_tensor(_DeviceView(d_res, (B, 4)), device=dev)
    fm = torch.as_tensor(_DeviceView(d_fm, (B, pipe.front_cap)), device=dev)
    bm = torch.as_tensor(_DeviceView(d_bm, (B, pipe.bird_cap)), device=dev)
    pipe.join()                                                  # the public stream now follows the step's last kernel
    ps = torch.cuda.ExternalStream(pipe.stream_ptr, device=dev)
    cur = torch.cuda.current_stream(dev)
    cur.wait_stream(ps)
    outs = []
    for t in (res, fm, bm):
        o = torch.empty((world,) + tuple(t.shape), dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_gather_into_tensor(o, t, group=group)
        else:
            o[0].copy_(t)
        outs.append(o)
    ps.wait_stream(cur)                                          # the next step overwrites the records only after they were read
    return tuple(outs)


def gather_fixed_records(parts, world: int, group=None):
    """The same gather for tensors on any device (gloo on CPU in the tests): [B, k] -> [world, B, k] for every tensor."""
    import torch
    import torch.distributed as dist
    outs = []
    for t in parts:
        if world > 1:
            lst = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(lst, t.contiguous(), group=group)
            outs.append(torch.stack(lst))
        else:
            outs.append(t.unsqueeze(0).clone())
    return tuple(outs)
