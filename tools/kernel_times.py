#!/usr/bin/env python
"""Per-kernel CUDA-event times of the device-resident trimap path at config B (or --shape), one
line per kernel -- the quick A/B harness used while tuning kernels (variants are selected with the
GG_* environment variables the library reads).  Also prints the un-profiled step time."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gcn_grabcut_b200 as gg                                    # noqa: E402
from gcn_grabcut_b200.synthetic import make_batch, random_state_dict   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--height", type=int, default=320)
ap.add_argument("--width", type=int, default=480)
ap.add_argument("--segments", type=int, default=300)
ap.add_argument("--k", type=int, default=4)
ap.add_argument("--hidden", type=int, default=128)
ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--top", type=int, default=12)
ap.add_argument("--tag", default="")
ap.add_argument("--slic", action="store_true", help="label maps from the device SLIC (images only in)")
a = ap.parse_args()

n_unique = min(a.batch, 32)                                     # the cost does not depend on the content
imgs, labs = make_batch(n_unique, a.height, a.width, a.segments, seed0=0,
                        scale=min(a.height, a.width) / 320.0 if min(a.height, a.width) > 320 else 1.0)
reps = (a.batch + n_unique - 1) // n_unique
imgs = np.concatenate([imgs] * reps)[:a.batch]
labs = np.concatenate([labs] * reps)[:a.batch]
path = gg.TrimapPath(random_state_dict(a.hidden, a.layers, seed=0),
                     gg.SuperpixelGraphConfig(n_segments=a.segments, n_nonlocal=a.k),
                     node_cap=int(labs.max()) + 1 + (a.segments // 4 if a.slic else 0), device_slic=a.slic)
img_d, lab_d = torch.from_numpy(imgs).cuda(), (None if a.slic else torch.from_numpy(labs).cuda())
tri_d = torch.empty(imgs.shape[:3], dtype=torch.uint8, device="cuda")
for _ in range(3):
    path.run_device(img_d, lab_d, tri_d)
path.check_status()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    path.run_device(img_d, lab_d, tri_d)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
h = path.h
h.set_option("n_sub", 1)
h.profile(True)
for _ in range(a.steps):
    path.run_device(img_d, lab_d, tri_d)
rows = h.profile_report()
h.profile(False)
tot = sum(r[2] for r in rows) / a.steps
digest = int(torch.bincount(tri_d.flatten().to(torch.int64), minlength=4).mul(torch.tensor([1, 7, 49, 343], device="cuda")).sum())
print(f"[{a.tag}] step {ms:.3f} ms ({a.batch / ms:.1f} k img/s), profiled (serial) {tot:.3f} ms, trimap digest {digest}")
for name, n, t in rows[:a.top]:
    print(f"    {name.rstrip(')'):44s} {n // a.steps:3d} launches/step  {t / a.steps:8.4f} ms/step")
