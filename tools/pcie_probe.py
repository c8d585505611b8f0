#!/usr/bin/env python
"""
Copy-only ceiling of the end-to-end number: plain pinned host<->device copies, nothing else.

    python tools/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29511 tools/pcie_probe.py                  # N = 1, 2, 4, 8 ranks copying at once

Per step a rank moves what one bench step moves at config B: 256 images x 320x480 x (3 B BGR +
4 B int32 labels) host->device as two plain cudaMemcpyAsync calls (one per buffer) and
256 x 320x480 x 1 B device->host, on two streams, `depth` steps in flight.  With several ranks
the phases n = 1, 2, 4, 8 run back to back inside one launch: ranks >= n idle at the barrier, so
that "what the host can deliver to n GPUs at once" is measured on the same box in the same
minute.  Prints one JSON line per phase (rank 0): images/s equivalent and GB/s, per rank and
aggregate -- the denominator for the e2e scaling efficiency in SCALE_*.json.
"""
import json
import os
import sys
import time

import torch

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)

B, H, W = 256, 320, 480
STEPS = int(os.environ.get("GG_PROBE_STEPS", "30"))
img_h = torch.empty((B, H, W, 3), dtype=torch.uint8).pin_memory()
lab_h = torch.empty((B, H, W), dtype=torch.int32).pin_memory()
tri_h = torch.empty((B, H, W), dtype=torch.uint8).pin_memory()
img_h.random_(0, 255); lab_h.random_(0, 300)
img_d = [torch.empty_like(img_h, device=dev) for _ in range(2)]
lab_d = [torch.empty_like(lab_h, device=dev) for _ in range(2)]
tri_d = torch.zeros((B, H, W), dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def steps(n, h2d=True, d2h=True):
    for i in range(n):
        if h2d:
            with torch.cuda.stream(s_in):
                img_d[i & 1].copy_(img_h, non_blocking=True)
                lab_d[i & 1].copy_(lab_h, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                tri_h.copy_(tri_d, non_blocking=True)
    torch.cuda.synchronize()


in_bytes = img_h.numel() + 4 * lab_h.numel()
out_bytes = tri_h.numel()
phases = [n for n in (1, 2, 4, 8) if n <= world]
for n in phases:
    for mode, (a, b) in (("h2d_only", (True, False)), ("h2d_and_d2h", (True, True))):
        active = rank < n
        if active:
            steps(3, a, b)
        barrier()
        t0 = time.perf_counter()
        if active:
            steps(STEPS, a, b)
        dt_mine = time.perf_counter() - t0
        barrier()
        t = torch.tensor([dt_mine if active else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        if rank == 0:
            per_rank_imgs = B * STEPS / dt
            print(json.dumps({"probe": "copy_only", "ranks_copying": n, "mode": mode,
                              "images_per_s_per_rank": per_rank_imgs, "images_per_s_aggregate": n * per_rank_imgs,
                              "h2d_GBps_per_rank": in_bytes * STEPS / dt / 1e9,
                              "h2d_GBps_aggregate": n * in_bytes * STEPS / dt / 1e9,
                              "d2h_GBps_per_rank": (out_bytes * STEPS / dt / 1e9) if b else 0.0,
                              "bytes_in_per_step": in_bytes, "bytes_out_per_step": out_bytes, "steps": STEPS}),
                  flush=True)
if rank == 0:
    import subprocess
    info = subprocess.run("nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader;"
                          " nproc; numactl -H 2>/dev/null | head -4; nvidia-smi topo -m 2>/dev/null | head -14",
                          shell=True, capture_output=True, text=True).stdout
    sys.stderr.write(info)
if world > 1:
    dist.destroy_process_group()
