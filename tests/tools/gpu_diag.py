#!/usr/bin/env python
"""
Stage-by-stage GPU-vs-oracle diagnostics (TEST TOOLING like tests/: a development aid that uses
the oracle as the checker; run on the B200 box):

    python tools/gpu_diag.py > gpurun_out/diag.log 2>&1

Prints, for each stage of the trimap path, the maximum deviation from the oracle and the
fraction of bit-identical values, and keeps going after a failing stage so that one GPU
call localises as many problems as possible.
"""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import gcn_grabcut_b200 as gg                                   # noqa: E402
from gcn_grabcut_b200 import _native as nat                     # noqa: E402
from gcn_grabcut_b200.synthetic import make_batch               # noqa: E402
from oracle import graph_port, model_port, trimap_port          # noqa: E402


def stat(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.shape != b.shape:
        print(f"  {name:28s} SHAPE MISMATCH {a.shape} vs {b.shape}")
        return
    if a.size == 0:
        print(f"  {name:28s} empty")
        return
    if a.dtype.kind in "iu":
        print(f"  {name:28s} equal={np.array_equal(a, b)}  n_diff={int((a != b).sum())}/{a.size}")
    else:
        d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        rel = d / (np.abs(b.astype(np.float64)) + 1e-12)
        i = np.unravel_index(np.argmax(d), d.shape)
        print(f"  {name:28s} max|d|={d.max():.3e} at {i} (got {a[i]:.7g} ref {b[i]:.7g}) "
              f"max rel={rel.max():.2e} bit-equal={np.mean(a == b):.5f}")


def section(title):
    print(f"\n==== {title}", flush=True)


def run(fn, title):
    section(title)
    t = time.time()
    try:
        fn()
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()
    print(f"  [{time.time() - t:.2f}s]", flush=True)


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    h = nat.handle(0)
    H, W, nseg, B = 320, 480, 300, 4
    imgs, labs = make_batch(B, H, W, nseg, seed0=0)
    refs = [graph_port.build_graph(imgs[b], labs[b]) for b in range(B)]
    state = model_port.random_state_dict(128, 6, seed=0)
    holder = {}

    def graph_stage():
        batch = gg.build_graph_batch(imgs, labs, gg.SuperpixelGraphConfig())
        holder["batch"] = batch
        graphs = batch.to_graphs(labs)
        holder["graphs"] = graphs
        print("  n_nodes", batch.n_nodes.cpu().tolist(), "ref", [r.n_nodes for r in refs])
        print("  n_edges", batch.n_edges.cpu().tolist(), "ref", [r.n_edges for r in refs])
        print("  n_adj", batch.n_adj_pairs.cpu().tolist(), "ref", [len(r.stages["adj_pairs"]) for r in refs])
        print("  n_nl", batch.n_nl_pairs.cpu().tolist(), "ref", [len(r.stages["nl_pairs"]) for r in refs])
        sh = batch.shared_cnt.cpu().numpy().reshape(B, -1)
        for b in range(B):
            g, r = graphs[b], refs[b]
            print(f" image {b}: knn ties {int(r.stages['knn_ties'])}")
            na = len(r.stages["adj_pairs"])
            stat("shared counts", sh[b, :na], r.stages["adj_counts"])
            if g.edge_index.shape == r.edge_index.shape:
                stat("edge_index", g.edge_index, r.edge_index)
                stat("edge_attr", g.edge_attr, r.edge_attr)
            else:
                print("  edge_index shape", g.edge_index.shape, r.edge_index.shape)
                m = min(g.edge_index.shape[1], r.edge_index.shape[1])
                print("  first differing column:",
                      int(np.argmax((g.edge_index[:, :m] != r.edge_index[:, :m]).any(0))))
            for c in range(16):
                stat(f"feat[{c}]", g.node_features[:, c], r.node_features[:, c])
            for c in range(3):
                stat(f"prior[{c}]", g.prior_features[:, c], r.prior_features[:, c])
            stat("centroids", g.node_centroids, r.node_centroids)
            stat("areas", g.node_areas, r.node_areas)

    def net_stage(impl):
        def f():
            h.set_option("gemm_impl", impl)
            net = gg.ResGCNNet(hidden_channels=128, n_layers=6)
            net.load_state_dict(state)
            net = net.to("cuda")
            for b in range(2):
                r = refs[b]
                data = gg.Data(x=torch.tensor(r.node_input()), edge_index=torch.tensor(r.edge_index),
                               edge_attr=torch.tensor(r.edge_attr)).to("cuda")
                logits = net(data).cpu().numpy()
                ref_logits, st = model_port.resgcn_forward(state, torch.tensor(r.node_input()),
                                                           torch.tensor(r.edge_index), torch.tensor(r.edge_attr),
                                                           return_stages=True)
                stat(f"logits[{b}]", logits, ref_logits.numpy())
                probs = net.predict_probs(data)
                ref_probs = torch.softmax(ref_logits, -1).numpy()
                stat(f"probs[{b}]", probs, ref_probs)
                holder[f"probs{b}"] = ref_probs
            h.set_option("gemm_impl", 1)
        return f

    def trimap_stage():
        for b in range(2):
            probs = holder.get(f"probs{b}")
            if probs is None:
                probs = np.random.RandomState(b).dirichlet([1, 1, 1], refs[b].n_nodes).astype(np.float32)
            tri, pbg, pfg = gg.refine_trimap(probs, labs[b], imgs[b], return_planes=True)
            rt, rbg, rfg = trimap_port.refine_trimap(probs, labs[b], imgs[b], return_planes=True)
            stat(f"p_bg[{b}]", pbg, rbg)
            stat(f"p_fg[{b}]", pfg, rfg)
            stat(f"trimap[{b}]", tri, rt)
        rng = np.random.RandomState(0)
        guide = rng.rand(97, 131).astype(np.float32)
        src = rng.rand(97, 131).astype(np.float32)
        stat("guided_filter r=8", gg.guided_filter(guide, src, 8, 1e-3), trimap_port.guided_filter(guide, src, 8, 1e-3))
        stat("guided_filter r=3", gg.guided_filter(guide, src, 3, 1e-2), trimap_port.guided_filter(guide, src, 3, 1e-2))

    def path_stage():
        path = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=int(labs.max()) + 1, chunk=2)
        tri, nn, ne = path(imgs, labs, return_counts=True)
        print("  n_nodes", nn.tolist(), "n_edges", ne.tolist())
        for b in range(B):
            r = refs[b]
            probs = model_port.predict_probs(state, torch.tensor(r.node_input()), torch.tensor(r.edge_index),
                                             torch.tensor(r.edge_attr))
            rt = trimap_port.refine_trimap(probs, labs[b], imgs[b])
            stat(f"path trimap[{b}]", tri[b], rt)
        print("  launches so far:", h.launches())

    run(graph_stage, "graph construction (4 x 320x480, N~300)")
    run(net_stage(0), "ResGCNNet forward, SIMT fp32 transforms")
    run(net_stage(1), "ResGCNNet forward, tcgen05 transforms")
    run(trimap_stage, "guided-filter trimap")
    run(path_stage, "whole path, host buffers, 2 chunks")


if __name__ == "__main__":
    main()
