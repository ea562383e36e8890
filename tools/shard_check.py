"""Multi-GPU check of the frame sharding (run under torchrun, one rank per GPU): every rank processes its contiguous block
of an offline sequence (plus its one-pair halo) on its own GPU, the fixed-stride records are all_gather-ed over NCCL, and
rank 0 compares the gathered table with a single-GPU run over the whole sequence.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/shard_check.py [pairs] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from fishbirdeyevisualslam_b200 import shard, synth

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 96
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fr = synth.cheap_batch(n_pairs, 720, 1280, 77)
bi = synth.cheap_batch(n_pairs, 384, 384, 78)
rec = shard.run_offline_batch(fr, bi, rank, world, batch, device=local)
table = shard.gather_records(rec, n_pairs, rank, world, backend_device=f"cuda:{local}") if world > 1 else rec
if rank == 0:
    ref = shard.run_offline_batch(fr, bi, 0, 1, batch, device=local)
    ok = np.array_equal(table, ref)
    print(f"shard_check: {world} GPU(s), {n_pairs} pairs, batch {batch}: gathered table {'==' if ok else '!='} single-GPU run; "
          f"matches front {int(ref[:, 2].sum())} bird {int(ref[:, 3].sum())}")
    if not ok:
        bad = np.nonzero((table != ref).any(axis=1))[0]
        print("first differing pairs:", bad[:10], table[bad[:3]], ref[bad[:3]])
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0)
