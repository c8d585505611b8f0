#!/usr/bin/env python
"""
Generate the golden fixtures of tests/golden/ by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it builds the inputs from seeds (gcn_grabcut_b200.synthetic), imports the
reference's own graph_builder.py / model.py / pipeline.py over the oracle shims
(oracle/ref_loader.py -- scikit-image and PyG are not installable here), runs
``GraphBuilder(image, cfg).build()``, ``ResGCNNet.predict_probs`` / ``forward`` and
``refine_trimap`` / ``predict_trimap``, and stores inputs digests + outputs as one
compressed .npz per case.  The reference ships no golden vectors of its own for this path
(tests/test.py only asserts shapes and ranges), so these are the pin for the oracle port.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_loader                      # noqa: E402
from oracle.model_port import random_state_dict    # noqa: E402
from gcn_grabcut_b200.synthetic import geometric_sample, slic_like_labels  # noqa: E402

CASES = [
    # name, H, W, image kind, seed, n_segments, connectivity, n_nonlocal, (D, n_layers)
    ("geom_160x192_n48", 160, 192, "geom", 3, 48, 4, 4, (32, 2)),
    ("rand_64x64_n50_c8", 64, 64, "rand", 42, 50, 8, 4, (32, 2)),
    ("rand_100x100_n30_k0", 100, 100, "rand", 7, 30, 4, 0, (32, 2)),
    ("rand_48x80_n4_small", 48, 80, "rand", 11, 4, 4, 4, (32, 2)),
    ("geom_320x480_n300", 320, 480, "geom", 0, 300, 4, 4, (128, 6)),
]


def make_image(kind: str, H: int, W: int, seed: int) -> np.ndarray:
    if kind == "geom":
        return geometric_sample(H, W, seed)[0]
    rng = np.random.RandomState(seed)              # the reference tests' _img helper (tests/test.py:17-19)
    return rng.randint(20, 220, (H, W, 3), dtype=np.uint8)


def digest(a: np.ndarray) -> str:
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = ref_loader.load()
    import skimage.segmentation as shim_seg        # the shim: lets us hand slic() our label map
    gb, model_mod, pipe = ref.graph_builder, ref.model, ref.pipeline
    from torch_geometric.data import Data, Batch

    for name, H, W, kind, seed, nseg, conn, knl, (D, nl) in CASES:
        img = make_image(kind, H, W, seed)
        seg = slic_like_labels(H, W, nseg, seed)
        cfg = gb.SuperpixelGraphConfig(n_segments=nseg, connectivity=conn, n_nonlocal=knl)
        shim_seg.set_next_labels(seg)
        builder = gb.GraphBuilder(img, cfg)
        graph = builder.build()
        assert np.array_equal(graph.segments, seg)

        state = random_state_dict(D, nl, seed=seed)
        net = model_mod.ResGCNNet(hidden_channels=D, n_layers=nl)
        net.load_state_dict(state)
        net.eval()
        data = Data(x=torch.tensor(graph.node_input(), dtype=torch.float32),
                    edge_index=torch.tensor(graph.edge_index, dtype=torch.long),
                    edge_attr=torch.tensor(graph.edge_attr, dtype=torch.float32))
        with torch.no_grad():
            logits = net(data).numpy()
        probs = net.predict_probs(data)
        tri_refined = pipe.refine_trimap(probs, graph.segments, img, 0.55, 0.55, radius=8)
        tri_r4 = pipe.refine_trimap(probs, graph.segments, img, 0.4, 0.45, radius=4, eps=1e-2)
        tri_direct = net.predict_trimap(data, graph.segments, 0.55, 0.55)
        guide = (builder._gray / np.float32(255.0)).astype(np.float32)
        p_fg = model_mod.project_to_pixels(probs[:, 2].astype(np.float32), seg)
        gf_fg = pipe.guided_filter(guide, p_fg, 8, 1e-3)

        # batched forward: this graph twice + a permuted copy, to pin the per-graph readout
        perm = torch.randperm(graph.n_nodes, generator=torch.Generator().manual_seed(seed))
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(graph.n_nodes)
        data_p = Data(x=data.x[perm], edge_index=inv[data.edge_index], edge_attr=data.edge_attr)
        with torch.no_grad():
            batched = net(Batch.from_data_list([data, data_p, data])).numpy()

        out = dict(
            H=H, W=W, seed=seed, n_segments=nseg, connectivity=conn, n_nonlocal=knl,
            hidden=D, n_layers=nl, image_sha1=digest(img), seg_sha1=digest(seg),
            node_features=graph.node_features, edge_index=graph.edge_index,
            edge_attr=graph.edge_attr, prior_features=graph.prior_features,
            node_centroids=graph.node_centroids, node_areas=graph.node_areas,
            n_nodes=graph.n_nodes, n_edges=graph.n_edges,
            logits=logits, probs=probs, logits_batched=batched, perm=perm.numpy(),
            trimap_refined=tri_refined, trimap_r4=tri_r4, trimap_direct=tri_direct,
            guided_fg=gf_fg.astype(np.float32),
        )
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **out)
        print(f"{name}: N={graph.n_nodes} E={graph.n_edges} -> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
