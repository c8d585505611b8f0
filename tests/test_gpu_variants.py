"""GPU parity of the network variants GCNTrimapNet / GATTrimapNet (reference
model.py:142-162, 216-414) against logits of the unmodified reference classes (golden fixture) and the
oracle restatement.  Tolerance: posterior <= 1e-4 (north_star), logits 5e-4."""
import numpy as np
import pytest
import torch

from test_oracle_golden import variant_cases, variant_state

pytestmark = pytest.mark.gpu

POST_ATOL = 1e-4


@pytest.fixture(scope="module")
def gg():
    import gcn_grabcut_b200 as gg
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return gg


def _net(gg, tag, z):
    D, n, H, _ = (int(v) for v in z[f"{tag}/meta"])
    net = gg.GCNTrimapNet(hidden_channels=D, n_layers=n) if tag.startswith("gcn") else \
        gg.GATTrimapNet(hidden_channels=D, n_heads=H, n_layers=n)
    missing, unexpected = net.load_state_dict(variant_state(z, tag), strict=True)
    assert not missing and not unexpected
    return net.to("cuda").eval()


@pytest.mark.parametrize("tag", variant_cases()[1])
def test_variant_vs_reference_golden(gg, tag):
    z, _ = variant_cases()
    net = _net(gg, tag, z)
    batch = torch.tensor(z[f"{tag}/batch"]) if f"{tag}/batch" in z else None
    data = gg.Data(x=torch.tensor(z[f"{tag}/x"]), edge_index=torch.tensor(z[f"{tag}/edge_index"]),
                   edge_attr=torch.tensor(z[f"{tag}/edge_attr"]), batch=batch).to("cuda")
    logits = net(data).cpu().numpy()
    probs = net.predict_probs(data)
    want = z[f"{tag}/logits"]
    want_p = torch.softmax(torch.tensor(want), -1).numpy()
    print(f"{tag}: max|dlogit| {np.abs(logits - want).max():.3g}  max|dprob| {np.abs(probs - want_p).max():.3g}")
    assert probs.dtype == np.float32 and probs.shape == want.shape
    np.testing.assert_allclose(logits, want, atol=5e-4, rtol=1e-4)
    np.testing.assert_allclose(probs, want_p, atol=POST_ATOL)


@pytest.mark.parametrize("variant", ["gcn", "gat"])
def test_variant_on_built_graph(gg, variant):
    """build_model(variant) on a graph the builder produced (the call pipeline.segment makes), against the
    oracle restatement; predict_trimap through the label map."""
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port
    imgs, labs = make_batch(1, 160, 192, 48, seed0=11)
    graph = gg.GraphBuilder(imgs[0], gg.SuperpixelGraphConfig(n_segments=48), segments=labs[0]).build()
    net = gg.build_model(variant, hidden_channels=64, n_layers=3)
    state = model_port.random_gcn_trimap_state(64, 3, seed=9) if variant == "gcn" else \
        model_port.random_gat_trimap_state(64, 8, 3, seed=9)
    net.load_state_dict(state)
    net = net.to("cuda")
    x, ei, ea = torch.tensor(graph.node_input()), torch.tensor(graph.edge_index), torch.tensor(graph.edge_attr)
    data = gg.Data(x=x, edge_index=ei, edge_attr=ea).to("cuda")
    probs = net.predict_probs(data)
    fwd = model_port.gcn_trimap_forward if variant == "gcn" else model_port.gat_trimap_forward
    want = torch.softmax(fwd(state, x, ei, ea), -1).numpy()
    np.testing.assert_allclose(probs, want, atol=POST_ATOL)
    tri = net.predict_trimap(data, graph.segments)
    assert np.array_equal(tri, model_port.probs_to_trimap(probs, graph.segments, 0.55, 0.55))


@pytest.mark.parametrize("variant", ["gcn", "gat"])
def test_variant_degenerate_graphs(gg, variant):
    """Graphs the reference accepts at the edges of its domain: no edges at all (every gate is 0 ->
    GCNTrimapNet still has its input row in the concat head, GATTrimapNet reduces to skip + context), a
    single node, self loops in the edge list (GCNConv / GATv2Conv drop and re-add them), missing edge_attr."""
    from oracle import model_port
    state = model_port.random_gcn_trimap_state(32, 2, seed=3) if variant == "gcn" else \
        model_port.random_gat_trimap_state(32, 4, 2, seed=3)
    net = gg.build_model(variant, hidden_channels=32, n_layers=2) if variant == "gcn" else \
        gg.GATTrimapNet(hidden_channels=32, n_heads=4, n_layers=2)
    net.load_state_dict(state)
    net = net.to("cuda")
    fwd = model_port.gcn_trimap_forward if variant == "gcn" else model_port.gat_trimap_forward
    gen = torch.Generator().manual_seed(0)
    cases = {
        "no_edges": (torch.randn(6, 19, generator=gen), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 5)),
        "one_node": (torch.randn(1, 19, generator=gen), torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, 5)),
        "self_loops": (torch.randn(5, 19, generator=gen), torch.tensor([[0, 1, 2, 2, 3, 4, 1], [1, 0, 2, 3, 2, 4, 3]]),
                       torch.rand(7, 5, generator=gen)),
    }
    for name, (x, ei, ea) in cases.items():
        got = net(gg.Data(x=x, edge_index=ei, edge_attr=ea).to("cuda")).cpu().numpy()
        want = fwd(state, x, ei, ea).numpy()
        assert np.isfinite(got).all(), name
        np.testing.assert_allclose(got, want, atol=5e-4, rtol=1e-4, err_msg=name)
    x, ei, _ = cases["self_loops"]
    got = net(gg.Data(x=x, edge_index=ei).to("cuda")).cpu().numpy()               # edge_attr None -> zeros (model.py:290)
    np.testing.assert_allclose(got, fwd(state, x, ei, torch.zeros(7, 5)).numpy(), atol=5e-4, rtol=1e-4)


def test_variant_state_errors(gg):
    """A forward for a variant whose weights are not loaded fails with GG_ERR_STATE; a state-dict with a
    missing key is rejected before anything reaches the device."""
    import ctypes as C
    from gcn_grabcut_b200 import _native as nat
    from oracle import model_port
    h = nat.handle(0)
    net = gg.GCNTrimapNet(hidden_channels=32, n_layers=1)
    net.load_state_dict(model_port.random_gcn_trimap_state(32, 1, seed=0))
    net = net.to("cuda")
    x = torch.randn(5, 19)
    ei = torch.tensor([[0, 1, 2, 3], [1, 0, 3, 2]])
    net(gg.Data(x=x, edge_index=ei, edge_attr=torch.rand(4, 5)).to("cuda"))
    xd = x.cuda()
    rp = torch.zeros(6, dtype=torch.int32, device="cuda")
    goff = torch.tensor([0, 5], dtype=torch.int64, device="cuda")
    out = torch.empty(5, 3, device="cuda")
    rc = nat.lib().gg_variant_forward(h.ptr, 2, nat.ptr(xd), nat.ptr(rp), None, None, None, nat.ptr(goff), 1, 5, 0,
                                      nat.ptr(out), None, C.c_void_p(0))
    assert rc == nat.GG_ERR_STATE
    bad = model_port.random_gcn_trimap_state(32, 1, seed=0)
    del bad["blocks.0.bn.running_var"]
    with pytest.raises(KeyError):
        nat.load_variant_state_dict(h, 1, 32, 1, 0, bad, net._tensor_keys())
