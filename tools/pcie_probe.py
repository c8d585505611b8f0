#!/usr/bin/env python
"""Measure pinned host<->device copy bandwidth on the GPU box (the ceiling of the e2e number):
H2D alone, D2H alone, both directions at once, for a few transfer sizes."""
import time
import torch

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
for mb in (8, 32, 68, 275):
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n // 7, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.empty(n // 7, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(8):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        d_in.copy_(h_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 8
        res[mode] = dt
    print(f"{mb:4d} MB: h2d {n / res['h2d'] / 1e9:6.1f} GB/s   d2h(1/7 size) {n / 7 / res['d2h'] / 1e9:6.1f} GB/s   "
          f"both: h2d-equivalent {n / res['both'] / 1e9:6.1f} GB/s", flush=True)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv; nproc; numactl -H 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -12",
                     shell=True, capture_output=True, text=True).stdout)
