"""CPU: host-side logic of the drop-in API (no GPU work): inputs, sharding, API surface."""
import numpy as np
import pytest
import torch

import gcn_grabcut_b200 as gg
from gcn_grabcut_b200.pipeline import shard_range
from gcn_grabcut_b200.synthetic import geometric_sample, grid_shape, make_batch, slic_like_labels


def test_geometric_sample_is_seeded_and_shaped():
    a, m = geometric_sample(320, 480, 5)
    b, _ = geometric_sample(320, 480, 5)
    c, _ = geometric_sample(320, 480, 6)
    assert a.shape == (320, 480, 3) and a.dtype == np.uint8 and m.shape == (320, 480)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert set(np.unique(m)) <= {0, 1} and m.sum() > 100


@pytest.mark.parametrize("H,W,n", [(320, 480, 300), (100, 100, 30), (64, 64, 50), (48, 80, 4), (1080, 1920, 2000)])
def test_label_maps_are_dense_and_contiguous(H, W, n):
    """Labels cover 0..N-1 with no gaps (the reference's own check: tests/test.py:112-117)."""
    seg = slic_like_labels(H, W, n, seed=1)
    ny, nx = grid_shape(H, W, n)
    assert seg.dtype == np.int32 and seg.shape == (H, W)
    assert np.array_equal(np.unique(seg), np.arange(ny * nx))
    assert abs(ny * nx - n) <= 0.15 * n + 2
    assert np.array_equal(seg, slic_like_labels(H, W, n, seed=1))


def test_make_batch():
    imgs, labs = make_batch(3, 128, 160, 40, seed0=9)
    assert imgs.shape == (3, 128, 160, 3) and labs.shape == (3, 128, 160)
    assert not np.array_equal(labs[0], labs[1])


@pytest.mark.parametrize("n,w", [(8192, 8), (256, 3), (5, 8), (0, 2), (1000, 1)])
def test_shard_range_partitions_exactly(n, w):
    spans = [shard_range(n, r, w) for r in range(w)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(n, w, w)


def test_api_surface_matches_reference_names():
    for name in ("GraphBuilder", "SuperpixelGraph", "SuperpixelGraphConfig", "compute_auto_prior",
                 "encode_user_hints", "N_NODE_FEATS", "N_EDGE_FEATS", "N_PRIOR_FEATS", "ResGCNNet",
                 "build_model", "_probs_to_trimap", "probs_to_node_trimap", "project_to_pixels",
                 "TRIMAP_BG", "TRIMAP_FG", "TRIMAP_PROB_BG", "TRIMAP_PROB_FG", "CLASS_BG", "CLASS_UNK",
                 "CLASS_FG", "guided_filter", "refine_trimap"):
        assert hasattr(gg, name), name
    assert (gg.N_NODE_FEATS, gg.N_EDGE_FEATS, gg.N_PRIOR_FEATS) == (19, 5, 3)
    assert (gg.TRIMAP_BG, gg.TRIMAP_FG, gg.TRIMAP_PROB_BG, gg.TRIMAP_PROB_FG) == (0, 1, 2, 3)
    cfg = gg.SuperpixelGraphConfig()
    assert (cfg.n_segments, cfg.compactness, cfg.sigma, cfg.use_lab, cfg.connectivity, cfg.n_nonlocal) == \
        (300, 10.0, 1.0, True, 4, 4)


def test_resgcn_container_matches_reference_checkpoints():
    from oracle.model_port import random_state_dict
    for D, n in ((128, 6), (96, 6), (32, 2)):
        net = gg.ResGCNNet(hidden_channels=D, n_layers=n)
        state = random_state_dict(D, n)
        assert set(net.state_dict()) == set(state)
        for k, v in net.state_dict().items():
            assert tuple(v.shape) == tuple(state[k].shape), k
        net.load_state_dict(state)                           # strict
        np.testing.assert_allclose(net.layer_weights().sum(), 1.0, rtol=1e-6)
    # README.md:564-566, :579 -- parameter counts of the reference architecture
    assert sum(p.numel() for p in gg.ResGCNNet(hidden_channels=128, n_layers=6).parameters()) == 187826
    assert sum(p.numel() for p in gg.ResGCNNet(hidden_channels=96, n_layers=6).parameters()) == 107090
    groups = gg.ResGCNNet(hidden_channels=32, n_layers=3).param_groups(1e-3)
    assert len(groups) > 1 and any(g["lr"] < 1e-3 for g in groups)
    with pytest.raises(ValueError):
        gg.build_model("nope")
    # the variants keep the reference's state-dict keys and shapes (model.py:239-414) and its README
    # parameter counts; without a GPU a forward fails loudly (no CPU fallback)
    from oracle.model_port import random_gcn_trimap_state, random_gat_trimap_state
    for net, state in ((gg.build_model("gcn", hidden_channels=64, n_layers=3), random_gcn_trimap_state(64, 3)),
                       (gg.GATTrimapNet(hidden_channels=64, n_heads=8, n_layers=2), random_gat_trimap_state(64, 8, 2))):
        assert set(net.state_dict()) == set(state)
        for k, v in net.state_dict().items():
            assert tuple(v.shape) == tuple(state[k].shape), k
        net.load_state_dict(state)
        assert sorted(net._tensor_keys()) == sorted(k for k in state if not k.endswith("num_batches_tracked"))
    assert isinstance(gg.build_model("gat"), gg.GATTrimapNet) and gg.build_model("gat").n_layers == 6
    with pytest.raises(ValueError):
        gg.GATTrimapNet(n_heads=3)
    if not torch.cuda.is_available():
        from gcn_grabcut_b200 import _native as nat
        with pytest.raises(nat.NativeError):
            gg.build_model("gcn", hidden_channels=32, n_layers=1)(gg.Data(x=torch.zeros(2, 19), edge_index=torch.zeros(2, 0, dtype=torch.long)))


def test_host_side_label_helpers():
    probs = np.array([[0.7, 0.2, 0.1], [0.1, 0.2, 0.7], [0.4, 0.3, 0.3], [0.3, 0.3, 0.4], [0.6, 0.0, 0.6]], np.float32)
    assert gg.probs_to_node_trimap(probs).tolist() == [0, 1, 2, 3, 1]         # FG overrides BG
    seg = np.array([[0, 1], [2, 5]], np.int32)
    vals = np.arange(3, dtype=np.float32)
    assert np.array_equal(gg.project_to_pixels(vals, seg), np.array([[0, 1], [2, 0]], np.float32))
    hints = gg.encode_user_hints(seg, [(0, 0)], [(1, 1), (9, 9)])
    assert hints.shape == (6, 3) and hints[0, 0] == 1 and hints[5, 1] == 1 and hints[1, 2] == 1


def test_graph_container_node_input():
    g = gg.SuperpixelGraph(segments=np.zeros((2, 2), np.int32), node_features=np.ones((3, 16), np.float32),
                           edge_index=np.zeros((2, 0), np.int64), edge_attr=np.zeros((0, 5), np.float32),
                           n_nodes=3, n_edges=0)
    x = g.node_input()
    assert x.shape == (3, 19) and x.dtype == np.float32 and np.all(x[:, 16:] == 0)
    d = g.to_pyg()
    assert d.x.shape == (3, 19) and d.edge_index.dtype == torch.long and d.node_area.shape == (3,)
