"""
CPU, world_size 2 over gloo: the multi-GPU layout of the path.  Images shard with no data-path
collective; the only collectives are the timing barrier / max-reduction and the final gather of
per-rank counts and result digests, exactly what bench.py (config D) does over NCCL.

The per-rank work is the trimap path itself -- on the CPU that is the oracle restatement
(graph build -> ResGCNNet posterior -> guided-filter trimap), the CUDA path has no CPU fallback --
so the test shows what the layout relies on: every reduction of the path is per image, hence the
trimaps of a batch sharded over two ranks are exactly the trimaps of the unsharded batch.
"""
import hashlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gcn_grabcut_b200.pipeline import shard_range

N_IMAGES, H, W, NSEG = 5, 48, 80, 12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch():
    from gcn_grabcut_b200.synthetic import slic_like_labels
    imgs = np.stack([np.random.RandomState(100 + i).randint(0, 256, (H, W, 3), dtype=np.uint8) for i in range(N_IMAGES)])
    labs = np.stack([slic_like_labels(H, W, NSEG, 100 + i) for i in range(N_IMAGES)]).astype(np.int32)
    return imgs, labs


def _trimaps(imgs, labs):
    """The path on the CPU (oracle), one image at a time."""
    from oracle import graph_port, model_port, trimap_port
    torch.set_num_threads(1)
    state = model_port.random_state_dict(32, 2, seed=4)
    out = []
    for img, lab in zip(imgs, labs):
        g = graph_port.build_graph(img, lab)
        probs = model_port.predict_probs(state, torch.from_numpy(g.node_input()), torch.from_numpy(g.edge_index),
                                         torch.from_numpy(g.edge_attr))
        out.append(trimap_port.refine_trimap(probs, lab, img))
    return out


def _digest(tri):
    return int.from_bytes(hashlib.sha1(np.ascontiguousarray(tri).tobytes()).digest()[:7], "little")


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    imgs, labs = _batch()
    lo, hi = shard_range(N_IMAGES, rank, world)
    mine = _trimaps(imgs[lo:hi], labs[lo:hi])                   # no exchange: a rank only touches its images
    elapsed = torch.tensor([0.010 * (rank + 1)], dtype=torch.float64)
    dist.barrier()
    dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)              # time = max over ranks
    # final gather (after the "timed region"): per-rank count and per-image digests, padded to the largest shard
    cap = -(-N_IMAGES // world)
    rec = torch.full((1 + cap,), -1, dtype=torch.int64)
    rec[0] = hi - lo
    rec[1:1 + len(mine)] = torch.tensor([_digest(t) for t in mine], dtype=torch.int64)
    parts = [torch.zeros_like(rec) for _ in range(world)]
    dist.all_gather(parts, rec)
    out_q.put((rank, lo, hi, float(elapsed), [p.tolist() for p in parts]))
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_the_unsharded_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, g0), (r1, lo1, hi1, t1, g1) = res
    assert (lo0, hi1) == (0, N_IMAGES) and hi0 == lo1 and abs((hi0 - lo0) - (hi1 - lo1)) <= 1   # exact, balanced partition
    assert t0 == t1 == 0.020                                                                       # max over ranks
    assert g0 == g1                                                                                # every rank holds the same gather
    counts = [part[0] for part in g0]
    assert counts == [hi0 - lo0, hi1 - lo1] and sum(counts) == N_IMAGES
    gathered = [d for part in g0 for d in part[1:1 + part[0]]]
    imgs, labs = _batch()
    whole = [_digest(t) for t in _trimaps(imgs, labs)]
    assert gathered == whole, "sharded trimaps differ from the unsharded batch"
    assert len(set(whole)) == N_IMAGES                                                             # the digests do discriminate


def test_shard_range_properties():
    for n in (0, 1, 7, 256, 8191, 8192):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    import pytest
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
