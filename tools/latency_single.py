#!/usr/bin/env python
"""Config A (BASELINE.json configs[0]) through the drop-in per-image API: one 320x480 image,
~300 regions, the three calls pipeline.segment() makes (GraphBuilder.build, model.predict_probs,
refine_trimap) with numpy in / numpy out, and the batched TrimapPath call with B = 1."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gcn_grabcut_b200 as gg                                   # noqa: E402
from gcn_grabcut_b200.synthetic import make_batch               # noqa: E402
from gcn_grabcut_b200.synthetic import random_state_dict          # noqa: E402

H, W, nseg = 320, 480, 300
imgs, labs = make_batch(4, H, W, nseg, seed0=0)
state = random_state_dict(128, 6, seed=0)
net = gg.ResGCNNet(hidden_channels=128, n_layers=6)
net.load_state_dict(state)
net = net.to("cuda").eval()
cfg = gg.SuperpixelGraphConfig(n_segments=nseg)


def per_image(i):
    t0 = time.perf_counter()
    graph = gg.GraphBuilder(imgs[i], cfg, segments=labs[i]).build()
    t1 = time.perf_counter()
    data = gg.Data(x=torch.tensor(graph.node_input()), edge_index=torch.tensor(graph.edge_index),
                   edge_attr=torch.tensor(graph.edge_attr)).to("cuda")
    probs = net.predict_probs(data)
    t2 = time.perf_counter()
    tri = gg.refine_trimap(probs, graph.segments, imgs[i], 0.55, 0.55, radius=8)
    tri = gg.seed_from_prior(tri, graph)
    t3 = time.perf_counter()
    return tri, (t1 - t0, t2 - t1, t3 - t2)


for i in range(4):
    per_image(i)                                                 # warm-up
ts = np.array([per_image(i % 4)[1] for i in range(40)]) * 1e3
print(f"drop-in per-image API (numpy in/out): graph_build {np.median(ts[:, 0]):.2f} ms, gcn_inference "
      f"{np.median(ts[:, 1]):.2f} ms, refine+seed {np.median(ts[:, 2]):.2f} ms, total {np.median(ts.sum(1)):.2f} ms")
tri_a = per_image(0)[0]
# one device handle holds one set of weights: the batched path is created after the per-image model is done
path = gg.TrimapPath(state, cfg, node_cap=int(labs.max()) + 1, seed_frac=0.1)
for _ in range(4):
    path(imgs[:1], labs[:1])
t = []
for _ in range(40):
    t0 = time.perf_counter()
    tri_b = path(imgs[:1], labs[:1])
    t.append(time.perf_counter() - t0)
print(f"TrimapPath(B=1) host call: {np.median(t) * 1e3:.2f} ms; same trimap as the three calls: "
      f"{bool(np.array_equal(tri_a, tri_b[0]))}")
