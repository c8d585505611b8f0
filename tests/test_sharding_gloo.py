"""
CPU, world_size 2 over gloo: the multi-GPU layout of the path.  Images shard with no
data-path collective; the only collectives are the timing barrier / max-reduction and the
final gather of per-rank counts, exactly what bench.py does over NCCL.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gcn_grabcut_b200.pipeline import shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_items, rank, world)
    # stand-in for the per-rank work: each rank "processes" its own images, no exchange
    ids = torch.arange(lo, hi, dtype=torch.int64)
    local_checksum = ids.sum()
    elapsed = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)            # time = max over ranks
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([hi - lo], dtype=torch.int64))
    total = local_checksum.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    out_q.put((rank, lo, hi, float(elapsed), [int(c) for c in counts], int(total)))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    world, n_items = 2, 8191
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, c0, s0), (r1, lo1, hi1, t1, c1, s1) = res
    assert (lo0, hi1) == (0, n_items) and hi0 == lo1                     # exact partition
    assert c0 == c1 == [hi0 - lo0, hi1 - lo1] and sum(c0) == n_items
    assert t0 == t1 == 0.020                                             # max over ranks
    assert s0 == s1 == n_items * (n_items - 1) // 2                      # every image exactly once
