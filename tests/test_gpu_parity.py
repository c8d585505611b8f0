"""
GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Every call goes through the
C ABI of libgcn_grabcut_b200.so (via the ctypes mirror of the reference API); the oracle
(oracle/*.py, a CPU restatement pinned against the reference's own outputs) is the checker.

Stated tolerances
  structure (n_nodes, edge_index, shared-boundary counts, degrees) ........ bit-exact
  node / edge attributes, prior ........................................... rtol 1e-5, atol 2e-6
  logits / posteriors ..................................................... atol 1e-4
  trimap .................................................................. pixel-exact except
        pixels whose filtered posterior lies within 1e-5 of a decision boundary
"""
import numpy as np
import pytest
import torch

from helpers import golden_cases, golden_inputs, load_golden

pytestmark = pytest.mark.gpu

FEAT_RTOL, FEAT_ATOL = 1e-5, 2e-6
POST_ATOL = 1e-4
TRI_TOL = 1e-5


@pytest.fixture(scope="module")
def gg():
    import gcn_grabcut_b200 as g
    from gcn_grabcut_b200 import _native
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    _native.handle(0)                        # fails loudly if the .so is missing / not sm_100
    return g


def _oracle_graph(img, seg, conn=4, k=4):
    """The oracle's graph.  Where the k-th and (k+1)-th colour distances of a region are EQUAL the
    reference's np.argpartition may pick either; the CUDA path documents "lower region index
    first", so for such images the oracle is asked for the same rule -- the comparison stays
    bit-exact either way (there is no tie escape)."""
    from oracle import graph_port
    ref = graph_port.build_graph(img, seg, conn, k)
    if int(ref.stages["knn_ties"]) > 0:
        ref = graph_port.build_graph(img, seg, conn, k, tie_break="lower_index")
    return ref


def _assert_graph_matches(got, ref, ties=0):
    assert got.n_nodes == ref.n_nodes
    assert got.n_edges == ref.n_edges
    assert np.array_equal(got.edge_index, ref.edge_index), "edge list differs (must be bit-exact)"
    for name in ("node_features", "prior_features", "node_centroids", "node_areas"):
        a, b = getattr(got, name), getattr(ref, name)
        assert a.dtype == np.float32 and a.shape == b.shape, name
        np.testing.assert_allclose(a, b, rtol=FEAT_RTOL, atol=FEAT_ATOL, err_msg=name)
    np.testing.assert_allclose(got.edge_attr, ref.edge_attr, rtol=FEAT_RTOL, atol=FEAT_ATOL)


# ----------------------------------------------------------------------------- pixel planes
def test_pixel_planes_all_colours(gg):
    """Lab / HSV of ALL 2^24 colours against the float64 restatement (float32-rounded)."""
    import ctypes as C
    from gcn_grabcut_b200 import _native as nat
    from oracle.thirdparty import rgb2hsv, rgb2lab
    h = nat.handle(0)
    idx = np.arange(1 << 24, dtype=np.uint32)
    bgr = np.stack([idx & 255, (idx >> 8) & 255, idx >> 16], -1).astype(np.uint8).reshape(1, 4096, 4096, 3)
    t = torch.from_numpy(bgr).cuda()
    lab = torch.empty(1, 4096, 4096, 3, dtype=torch.float32, device="cuda")
    hsv = torch.empty_like(lab)
    nat.check(nat.lib().gg_pixel_planes(h.ptr, nat.ptr(t), 1, 4096, 4096, nat.ptr(lab), nat.ptr(hsv),
                                        C.c_void_p(0), C.c_void_p(0), C.c_void_p(nat.current_stream(0))))
    lab, hsv = lab.cpu().numpy()[0], hsv.cpu().numpy()[0]
    bad_lab = bad_hsv = 0
    worst = 0.0
    for r0 in range(0, 4096, 256):
        rgb = bgr[0, r0:r0 + 256, :, ::-1]
        rl = rgb2lab(rgb).astype(np.float32)
        rh = rgb2hsv(rgb).astype(np.float32)
        bad_lab += int(np.sum(rl != lab[r0:r0 + 256]))
        bad_hsv += int(np.sum(rh != hsv[r0:r0 + 256]))
        worst = max(worst, float(np.abs(rl - lab[r0:r0 + 256]).max()))
    print(f"Lab float32 mismatches: {bad_lab} of {3 << 24} (max |d| {worst:.3g}); HSV mismatches: {bad_hsv}")
    assert bad_hsv == 0                       # exact-rational argument, must hold bit for bit
    assert bad_lab <= 64 and worst < 1e-4     # 1-ulp float32 events of the fp64 cbrt only


def test_gray_and_gradient_planes(gg):
    import ctypes as C
    import cv2
    from gcn_grabcut_b200 import _native as nat
    h = nat.handle(0)
    rng = np.random.RandomState(5)
    img = rng.randint(0, 256, (2, 37, 53, 3), dtype=np.uint8)
    t = torch.from_numpy(img).cuda()
    gray = torch.empty(2, 37, 53, dtype=torch.float32, device="cuda")
    grad = torch.empty_like(gray)
    nat.check(nat.lib().gg_pixel_planes(h.ptr, nat.ptr(t), 2, 37, 53, C.c_void_p(0), C.c_void_p(0),
                                        nat.ptr(gray), nat.ptr(grad), C.c_void_p(nat.current_stream(0))))
    for b in range(2):
        g = cv2.cvtColor(img[b], cv2.COLOR_BGR2GRAY).astype(np.float32)
        gx = cv2.Sobel(g, cv2.CV_32F, 1, 0, ksize=3)
        gy = cv2.Sobel(g, cv2.CV_32F, 0, 1, ksize=3)
        assert np.array_equal(gray[b].cpu().numpy(), g)
        assert np.array_equal(grad[b].cpu().numpy(), np.sqrt(gx ** 2 + gy ** 2))


# ----------------------------------------------------------------------------- graph construction
@pytest.mark.parametrize("name", golden_cases())
def test_graph_builder_vs_golden(gg, name):
    """GraphBuilder(image, cfg, segments).build() against the reference's own output."""
    g = load_golden(name)
    img, seg = golden_inputs(g, name)
    cfg = gg.SuperpixelGraphConfig(n_segments=int(g["n_segments"]), connectivity=int(g["connectivity"]),
                                   n_nonlocal=int(g["n_nonlocal"]))
    got = gg.GraphBuilder(img, cfg, segments=seg).build()
    assert got.n_nodes == int(g["n_nodes"]) and got.n_edges == int(g["n_edges"])
    assert got.edge_index.dtype == np.int64 and np.array_equal(got.edge_index, g["edge_index"])
    assert got.node_input().shape == (got.n_nodes, 19)
    for key, mine in (("node_features", got.node_features), ("edge_attr", got.edge_attr),
                      ("prior_features", got.prior_features), ("node_centroids", got.node_centroids),
                      ("node_areas", got.node_areas)):
        np.testing.assert_allclose(mine, g[key], rtol=FEAT_RTOL, atol=FEAT_ATOL, err_msg=key)
        print(f"{name}:{key}: bit-equal {np.mean(mine == g[key]):.5f}, max|d| {np.abs(mine - g[key]).max():.3g}")


@pytest.mark.parametrize("conn,k,H,W,nseg", [(4, 4, 320, 480, 300), (8, 4, 200, 264, 120),
                                             (4, 8, 240, 320, 200), (4, 0, 128, 160, 60),
                                             (4, 16, 161, 203, 90), (4, 8, 300, 420, 700)])
def test_graph_batch_vs_oracle(gg, conn, k, H, W, nseg):
    from gcn_grabcut_b200.synthetic import make_batch
    B = 6
    imgs, labs = make_batch(B, H, W, nseg, seed0=100)
    cfg = gg.SuperpixelGraphConfig(n_segments=nseg, connectivity=conn, n_nonlocal=k)
    batch = gg.build_graph_batch(imgs, labs, cfg)
    graphs = batch.to_graphs(labs)
    shared = batch.shared_cnt.cpu().numpy().reshape(B, batch.pair_cap)
    n_adj = batch.n_adj_pairs.cpu().numpy()
    total_ties = 0
    for b in range(B):
        ref = _oracle_graph(imgs[b], labs[b], conn, k)
        ties = int(ref.stages["knn_ties"])
        total_ties += ties
        _assert_graph_matches(graphs[b], ref, ties)
        # adjacency pairs and shared boundary lengths: integer, bit-exact
        assert n_adj[b] == len(ref.stages["adj_pairs"])
        assert np.array_equal(shared[b, :n_adj[b]], ref.stages["adj_counts"])
        assert np.array_equal(graphs[b].edge_index[:, :n_adj[b]].T, ref.stages["adj_pairs"])
    print(f"conn={conn} k={k}: kNN tie rows over the batch: {total_ties}")


def test_graph_edge_cases(gg):
    rng = np.random.RandomState(3)
    # (a) ragged label maps: absent labels (area-0 regions), label 0 absent, tiny non-multiple-of-32 image
    img = rng.randint(0, 256, (2, 19, 45, 3), dtype=np.uint8)
    seg = np.zeros((2, 19, 45), np.int32)
    seg[0, :, 20:] = 3
    seg[0, 10:, :10] = 5
    seg[1, :9, :] = 1
    seg[1, 9:, :22] = 2
    seg[1, 9:, 22:] = 7
    for k in (0, 2, 4):
        cfg = gg.SuperpixelGraphConfig(connectivity=8, n_nonlocal=k)
        graphs = gg.build_graph_batch(img, seg, cfg).to_graphs(seg)
        for b in range(2):
            ref = _oracle_graph(img[b], seg[b], 8, k)
            _assert_graph_matches(graphs[b], ref, int(ref.stages["knn_ties"]))
    # (b) N <= k+1: no non-local edges (graph_builder.py:291)
    seg2 = np.zeros((1, 32, 32), np.int32)
    seg2[0, :, 16:] = 1
    seg2[0, 16:, :16] = 2
    img2 = rng.randint(0, 256, (1, 32, 32, 3), dtype=np.uint8)
    g2 = gg.build_graph_batch(img2, seg2, gg.SuperpixelGraphConfig(n_nonlocal=4)).to_graphs(seg2)[0]
    ref2 = _oracle_graph(img2[0], seg2[0], 4, 4)
    _assert_graph_matches(g2, ref2)
    assert g2.n_edges == 2 * 3 and np.all(g2.edge_attr[:, 4] == 0)
    # (c) label >= node_cap is reported, not silently dropped
    from gcn_grabcut_b200._native import NativeError
    with pytest.raises(NativeError):
        gg.build_graph_batch(img2, seg2, node_cap=2)


def test_graph_invariants_full_size(gg):
    """Config B (256 x 320x480, ~300 regions): size-independent properties of the result."""
    from gcn_grabcut_b200.synthetic import make_batch
    B, H, W = 256, 320, 480
    imgs, labs = make_batch(8, H, W, 300, seed0=7)
    reps = B // 8
    imgs_b, labs_b = np.tile(imgs, (reps, 1, 1, 1)), np.tile(labs, (reps, 1, 1))
    batch = gg.build_graph_batch(imgs_b, labs_b, gg.SuperpixelGraphConfig())
    no, eo = batch.node_off.cpu().numpy(), batch.edge_off.cpu().numpy()
    x = batch.x.cpu().numpy()
    ei = batch.edge_index.cpu().numpy()
    for b in range(B):
        n0, n1, e0, e1 = no[b], no[b + 1], eo[b], eo[b + 1]
        assert n1 - n0 == labs_b[b].max() + 1
        np.testing.assert_allclose(x[n0:n1, 11].sum(), 1.0, rtol=1e-5)        # area ratios sum to 1
        src, dst = ei[0, e0:e1], ei[1, e0:e1]
        half = (e1 - e0) // 2
        assert np.array_equal(src[:half], dst[half:]) and np.array_equal(dst[:half], src[half:])
        assert np.all(src[:half] < dst[:half])
        assert np.array_equal(np.bincount(src, minlength=n1 - n0), np.bincount(dst, minlength=n1 - n0))
    # identical images in the batch give identical graphs (replicas 0 and 8, 16, ...)
    for b in range(8, B, 8):
        assert np.array_equal(ei[:, eo[b]:eo[b + 1]], ei[:, eo[0]:eo[1]])
        np.testing.assert_allclose(x[no[b]:no[b + 1]], x[no[0]:no[1]], rtol=1e-6, atol=1e-7)
    # against the oracle on the 8 distinct images
    graphs = batch.to_graphs(labs_b)
    for b in range(8):
        ref = _oracle_graph(imgs[b], labs[b])
        _assert_graph_matches(graphs[b], ref, int(ref.stages["knn_ties"]))


# ----------------------------------------------------------------------------- network
def _golden_data(gg, g):
    x = torch.tensor(np.concatenate([g["node_features"], g["prior_features"]], 1))
    return gg.Data(x=x, edge_index=torch.tensor(g["edge_index"]), edge_attr=torch.tensor(g["edge_attr"]))


@pytest.mark.parametrize("impl", ["simt", "tc"])
@pytest.mark.parametrize("name", golden_cases())
def test_resgcn_vs_golden(gg, name, impl):
    from gcn_grabcut_b200 import _native as nat
    from oracle.model_port import random_state_dict
    g = load_golden(name)
    D, n = int(g["hidden"]), int(g["n_layers"])
    net = gg.ResGCNNet(hidden_channels=D, n_layers=n)
    net.load_state_dict(random_state_dict(D, n, seed=int(g["seed"])))
    net = net.to("cuda").eval()
    nat.handle(0).set_option("gemm_impl", 0 if impl == "simt" else 1)
    try:
        data = _golden_data(gg, g).to("cuda")
        logits = net(data).cpu().numpy()
        probs = net.predict_probs(data)
    finally:
        nat.handle(0).set_option("gemm_impl", 1)
    print(f"{name}[{impl}]: max|dlogit| {np.abs(logits - g['logits']).max():.3g}  "
          f"max|dprob| {np.abs(probs - g['probs']).max():.3g}")
    assert probs.dtype == np.float32 and probs.shape == (int(g["n_nodes"]), 3)
    np.testing.assert_allclose(probs.sum(1), 1.0, atol=1e-5)
    np.testing.assert_allclose(logits, g["logits"], atol=5e-4, rtol=1e-4)
    np.testing.assert_allclose(probs, g["probs"], atol=POST_ATOL)


def test_resgcn_batched_equals_single(gg):
    """The reference's invariant (tests/test.py:294-306): a batch == one graph at a time @1e-4,
    on its own test graphs (path graph, randn features, seeded)."""
    from oracle import model_port
    state = model_port.random_state_dict(32, 2, seed=5)
    net = gg.ResGCNNet(hidden_channels=32, n_layers=2)
    net.load_state_dict(state)
    net = net.to("cuda")

    def make(N, seed):
        gen = torch.Generator().manual_seed(seed)
        x = torch.randn(N, 19, generator=gen)
        s, d = torch.arange(N - 1), torch.arange(1, N)
        ei = torch.stack([torch.cat([s, d]), torch.cat([d, s])])
        ea = torch.rand(ei.size(1), 5, generator=gen)
        return x, ei, ea
    graphs = [make(40, s) for s in (1, 2, 3)]
    one = torch.cat([net(gg.Data(x=x, edge_index=ei, edge_attr=ea).to("cuda")).cpu() for x, ei, ea in graphs])
    xb = torch.cat([g[0] for g in graphs])
    eib = torch.cat([g[1] + 40 * i for i, g in enumerate(graphs)], 1)
    eab = torch.cat([g[2] for g in graphs])
    batch = torch.arange(3).repeat_interleave(40)
    both = net(gg.Data(x=xb, edge_index=eib, edge_attr=eab, batch=batch).to("cuda")).cpu()
    assert both.shape == (120, 3)
    assert torch.allclose(one, both, atol=1e-4)
    ref = model_port.resgcn_forward(state, xb, eib, eab, batch)
    assert torch.allclose(both, ref, atol=5e-4, rtol=1e-4)
    # output depends on input (tests/test.py:282-292)
    assert not torch.allclose(one[:40], one[40:80])
    # shuffled COO order gives the same result (CSR is rebuilt deterministically)
    perm = torch.randperm(eib.size(1), generator=torch.Generator().manual_seed(0))
    shuf = net(gg.Data(x=xb, edge_index=eib[:, perm], edge_attr=eab[perm], batch=batch).to("cuda")).cpu()
    assert torch.allclose(shuf, both, atol=1e-5)


def test_predict_trimap_and_node_labels(gg):
    from oracle import model_port
    g = load_golden("geom_160x192_n48")
    img, seg = golden_inputs(g, "geom_160x192_n48")
    net = gg.ResGCNNet(hidden_channels=32, n_layers=2)
    net.load_state_dict(model_port.random_state_dict(32, 2, seed=int(g["seed"])))
    net = net.to("cuda")
    tri = net.predict_trimap(_golden_data(gg, g).to("cuda"), seg, 0.55, 0.55)
    assert tri.shape == seg.shape and tri.dtype == np.uint8
    assert set(np.unique(tri)).issubset({0, 1, 2, 3})
    probs = net.predict_probs(_golden_data(gg, g).to("cuda"))
    assert np.array_equal(tri, model_port.probs_to_trimap(probs, seg, 0.55, 0.55))
    # padding rule: labels without a probability row become GC_PR_BGD (model.py:672-677)
    assert np.array_equal(gg._probs_to_trimap(g["probs"][:10], seg, 0.55, 0.55),
                          model_port.probs_to_trimap(g["probs"][:10], seg, 0.55, 0.55))
    assert np.array_equal(gg.probs_to_node_trimap(g["probs"]), model_port.probs_to_node_trimap(g["probs"]))


# ----------------------------------------------------------------------------- projection
@pytest.mark.parametrize("name", golden_cases())
def test_refine_trimap_vs_golden(gg, name):
    from oracle import trimap_port
    g = load_golden(name)
    img, seg = golden_inputs(g, name)
    for (tf, tb, r, eps, key) in ((0.55, 0.55, 8, 1e-3, "trimap_refined"), (0.4, 0.45, 4, 1e-2, "trimap_r4")):
        tri, pbg, pfg = gg.refine_trimap(g["probs"], seg, img, tf, tb, radius=r, eps=eps, return_planes=True)
        _, rbg, rfg = trimap_port.refine_trimap(g["probs"], seg, img, tf, tb, r, eps, return_planes=True)
        np.testing.assert_allclose(pbg, rbg, atol=2e-6)
        np.testing.assert_allclose(pfg, rfg, atol=2e-6)
        near = trimap_port.near_threshold_mask(rbg, rfg, tf, tb, TRI_TOL)
        bad = (tri != g[key]) & ~near
        print(f"{name}:{key}: mismatches {int((tri != g[key]).sum())} (near a threshold: {int(near.sum())}), "
              f"planes bit-equal {np.mean(pfg == rfg):.5f}")
        assert tri.dtype == np.uint8 and not bad.any()


def test_guided_filter_vs_cv2(gg):
    from oracle import trimap_port
    rng = np.random.RandomState(0)
    for (H, W, r) in ((97, 131, 8), (64, 64, 4), (33, 200, 1), (20, 24, 12)):
        guide = rng.rand(H, W).astype(np.float32)
        src = (rng.rand(H, W) > 0.5).astype(np.float32) * 0.7 + 0.1
        out = gg.guided_filter(guide, src, r, 1e-3)
        ref = trimap_port.guided_filter(guide, src, r, 1e-3)
        np.testing.assert_allclose(out, ref, atol=3e-6, rtol=1e-5)


# ----------------------------------------------------------------------------- whole path
@pytest.mark.parametrize("edge_aware", [True, False])
def test_trimap_path_host_vs_oracle(gg, edge_aware):
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port, trimap_port
    B, H, W = 10, 200, 264
    imgs, labs = make_batch(B, H, W, 120, seed0=40)
    state = model_port.random_state_dict(64, 3, seed=1)
    path = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=int(labs.max()) + 1,
                         edge_aware=edge_aware, chunk=4)            # 3 chunks: exercises the pipelining
    tri, nn, ne = path(imgs, labs, return_counts=True)
    assert tri.shape == (B, H, W) and tri.dtype == np.uint8
    n_bad = 0
    for b in range(B):
        ref = _oracle_graph(imgs[b], labs[b])
        assert nn[b] == ref.n_nodes and ne[b] == ref.n_edges
        probs = model_port.predict_probs(state, torch.tensor(ref.node_input()), torch.tensor(ref.edge_index),
                                         torch.tensor(ref.edge_attr))
        if edge_aware:
            rt, rbg, rfg = trimap_port.refine_trimap(probs, labs[b], imgs[b], return_planes=True)
            near = trimap_port.near_threshold_mask(rbg, rfg, 0.55, 0.55, 2e-4)
        else:
            rt = model_port.probs_to_trimap(probs, labs[b], 0.55, 0.55)
            pn = probs[labs[b]]
            near = trimap_port.near_threshold_mask(pn[..., 0], pn[..., 2], 0.55, 0.55, 2e-4)
        n_bad += int(((tri[b] != rt) & ~near).sum())
    assert n_bad == 0
    # device-resident entry point gives the same trimaps
    td = path.run_device(torch.from_numpy(imgs).cuda(), torch.from_numpy(labs).cuda())
    assert np.array_equal(td.cpu().numpy(), tri)


def test_knn_ties_lower_index_rule(gg):
    """Exact kNN ties, constructed: a regular grid of equal-sized flat-coloured cells whose
    colours come from a 5-colour palette gives many regions with bit-identical mean Lab, i.e.
    many equal colour distances (0 and otherwise) at the k-th position.  The CUDA path must follow
    its documented rule (equal distances -> lower region index), which the oracle restates with a
    stable sort; edge lists bit-exact, k in {2, 4, 8}, both selection kernels (N <= 320 registers,
    N > 512 shared memory) and the per-lane top-k kernel (N > 2048)."""
    from oracle import graph_port
    palette = np.array([[20, 200, 90], [20, 200, 90], [250, 10, 10], [128, 128, 128], [0, 0, 0], [37, 99, 181]],
                       dtype=np.uint8)
    for (gy, gx, cell, ks) in ((12, 16, 8, (2, 4, 8)), (30, 36, 6, (4,)), (50, 48, 4, (4,))):
        H, W = gy * cell, gx * cell
        rng = np.random.RandomState(gy)
        col = palette[rng.randint(0, len(palette), gy * gx)]
        seg = (np.arange(H)[:, None] // cell * gx + np.arange(W)[None, :] // cell).astype(np.int32)
        img = col[seg]
        for k in ks:
            ref = graph_port.build_graph(img, seg, 4, k, tie_break="lower_index")
            assert int(ref.stages["knn_ties"]) > 0, "the constructed image must contain ties"
            cfg = gg.SuperpixelGraphConfig(n_segments=gy * gx, n_nonlocal=k)
            got = gg.build_graph_batch(img[None], seg[None], cfg).to_graphs(seg[None])[0]
            assert got.n_edges == ref.n_edges and np.array_equal(got.edge_index, ref.edge_index), (gy * gx, k)
            np.testing.assert_allclose(got.edge_attr, ref.edge_attr, rtol=FEAT_RTOL, atol=FEAT_ATOL)
            print(f"N={gy * gx} k={k}: {int(ref.stages['knn_ties'])} tie rows, edge list bit-exact")


# ----------------------------------------------------------------------------- larger configurations
def test_knn_topk_kernel_vs_oracle(gg):
    """The per-lane top-k kernel k_knn<K> (graphs of more than 2048 regions in production, also covered at N ~ 10^4 by the config E test) forced on a
    700-region graph with the "knn_legacy" option: same edge list as the oracle, bit for bit, for k = 4 and 16."""
    from gcn_grabcut_b200 import _native as nat
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import graph_port
    imgs, labs = make_batch(1, 333, 1001, 700, seed0=21)
    h = nat.handle(0)
    h.set_option("knn_legacy", 1)
    try:
        for k in (4, 16):
            g = gg.build_graph_batch(imgs, labs, gg.SuperpixelGraphConfig(n_segments=700, n_nonlocal=k)).to_graphs(labs)[0]
            o = graph_port.build_graph(imgs[0], labs[0], n_nonlocal=k)
            if int(o.stages["knn_ties"]) > 0:
                o = graph_port.build_graph(imgs[0], labs[0], n_nonlocal=k, tie_break="lower_index")
            assert g.n_edges == o.n_edges and np.array_equal(g.edge_index, o.edge_index), f"k={k}"
    finally:
        h.set_option("knn_legacy", 0)


def test_config_c_full_hd_dense_nonlocal(gg):
    """BASELINE config C shape: 1080x1920, ~2000 superpixels, dense non-local edges (k=16),
    against the oracle on one image (the CPU side takes a few seconds per image)."""
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port, trimap_port
    H, W, nseg, k = 1080, 1920, 2000, 16
    imgs, labs = make_batch(2, H, W, nseg, seed0=11, scale=min(H, W) / 320)
    cfg = gg.SuperpixelGraphConfig(n_segments=nseg, n_nonlocal=k)
    graphs = gg.build_graph_batch(imgs, labs, cfg).to_graphs(labs)
    ref = _oracle_graph(imgs[0], labs[0], 4, k)
    _assert_graph_matches(graphs[0], ref)
    print(f"config C: N={ref.n_nodes} E={ref.n_edges} knn ties={int(ref.stages['knn_ties'])}")
    state = model_port.random_state_dict(128, 6, seed=0)
    path = gg.TrimapPath(state, cfg, node_cap=int(labs.max()) + 1)
    tri = path(imgs, labs)
    probs = model_port.predict_probs(state, torch.tensor(ref.node_input()), torch.tensor(ref.edge_index),
                                     torch.tensor(ref.edge_attr))
    rt, rbg, rfg = trimap_port.refine_trimap(probs, labs[0], imgs[0], return_planes=True)
    near = trimap_port.near_threshold_mask(rbg, rfg, 0.55, 0.55, 2e-4)
    assert int(((tri[0] != rt) & ~near).sum()) == 0


def test_config_e_4k_vs_oracle(gg):
    """BASELINE config E shape: 2160x3840, ~10k regions, wider/deeper network (D=256, n=8),
    against the oracle in full: the reference's N x N matrices (kNN distances, contrast) are
    evaluated by the oracle in row blocks (oracle/graph_port.py), so the per-lane top-k kernel
    k_knn<K> that serves N > 2048 is compared bit for bit like every other size."""
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port, trimap_port
    H, W, nseg = 2160, 3840, 10000
    imgs, labs = make_batch(1, H, W, nseg, seed0=5, scale=min(H, W) / 320)
    cfg = gg.SuperpixelGraphConfig(n_segments=nseg, n_nonlocal=4)
    batch = gg.build_graph_batch(imgs, labs, cfg)
    g = batch.to_graphs(labs)[0]
    n = int(labs.max()) + 1
    ref = _oracle_graph(imgs[0], labs[0], 4, 4)
    assert ref.n_nodes == n > 2048
    _assert_graph_matches(g, ref)
    na = int(batch.n_adj_pairs[0])
    assert na == len(ref.stages["adj_pairs"])
    assert np.array_equal(batch.shared_cnt.cpu().numpy()[:na], ref.stages["adj_counts"])
    print(f"config E: N={ref.n_nodes} E={ref.n_edges} knn tie rows={int(ref.stages['knn_ties'])}; "
          f"prior max|d| {np.abs(g.prior_features - ref.prior_features).max():.3g}")
    state = model_port.random_state_dict(256, 8, seed=0)
    path = gg.TrimapPath(state, cfg, node_cap=n)
    tri = path(imgs, labs)
    probs = model_port.predict_probs(state, torch.tensor(ref.node_input()), torch.tensor(ref.edge_index),
                                     torch.tensor(ref.edge_attr))
    net = gg.ResGCNNet(hidden_channels=256, n_layers=8)
    net.load_state_dict(state)
    got_probs = net.to("cuda").predict_probs(gg.Data(x=torch.tensor(g.node_input()), edge_index=torch.tensor(g.edge_index),
                                                     edge_attr=torch.tensor(g.edge_attr)).to("cuda"))
    assert np.abs(got_probs - probs).max() < POST_ATOL
    rt, rbg, rfg = trimap_port.refine_trimap(probs, labs[0], imgs[0], return_planes=True)
    near = trimap_port.near_threshold_mask(rbg, rfg, 0.55, 0.55, 2e-4)
    assert tri.shape == (1, H, W) and int(((tri[0] != rt) & ~near).sum()) == 0


def test_trimap_path_streaming_submit_result(gg):
    """submit/result with several batches in flight (chunk slots rotate across calls) gives the
    same trimaps and counts as one synchronous call per batch; a capacity overflow in any batch in
    flight is reported by a wait and the handle stays usable."""
    import gcn_grabcut_b200._native as nat
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port
    H, W = 120, 168
    batches = [make_batch(b, H, W, 60, seed0=70 + 10 * i) for i, b in enumerate((7, 3, 9, 1, 6))]
    node_cap = max(int(l.max()) for _, l in batches) + 1
    state = model_port.random_state_dict(64, 2, seed=2)
    path = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=node_cap, chunk=2)
    want = [path(i, l, return_counts=True) for i, l in batches]
    pins = [(torch.from_numpy(i).pin_memory(), torch.from_numpy(l).pin_memory()) for i, l in batches]
    for rep in range(2):
        pending = [path.submit(i, l, return_counts=True) for i, l in pins]
        got = [p.result() for p in reversed(pending)][::-1]        # waits may come in any order
        for (t0, n0, e0), (t1, n1, e1) in zip(want, got):
            assert np.array_equal(t0, t1) and np.array_equal(n0, n1) and np.array_equal(e0, e1)
    # more than 8 calls in flight are refused, not queued
    pending = [path.submit(*pins[3]) for _ in range(8)]
    with pytest.raises(nat.NativeError):
        path.submit(*pins[3])
    for p in pending:
        assert np.array_equal(p.result(), want[3][0])
    # overflow: a label >= node_cap
    bad = batches[1][1].copy()
    bad[0, 0, 0] = node_cap + 5
    p_ok, p_bad = path.submit(*pins[0]), path.submit(batches[1][0], bad)
    with pytest.raises(nat.NativeError):
        p_ok.result()
        p_bad.result()
    for p in (p_ok, p_bad):
        try:
            p.result()
        except nat.NativeError:
            pass
    assert np.array_equal(path(*batches[2]), want[2][0])


def test_fast_float32_paths_selftest(gg):
    """Device self test: the short division / square-root sequences of the pixel kernels return
    the IEEE round-to-nearest results on their whole domains (csrc/pixel_math.cuh)."""
    import ctypes as C
    from gcn_grabcut_b200 import _native as nat
    h = nat.handle(0)
    out = (C.c_int64 * 4)()
    nat.check(nat.lib().gg_selftest_math(h.ptr, out))
    print("selftest mismatches [sqrt, sat, hue, gradn]:", list(out))
    assert list(out) == [0, 0, 0, 0]


# ----------------------------------------------------------------------------- trimap hand-off
def test_seed_from_prior_vs_reference_golden(gg):
    """gg_seed_from_prior against the reference's own _seed_from_prior outputs
    (tests/golden/handoff/seed_from_prior.npz): bit-exact, for graphs built by the CUDA builder."""
    import os
    from gcn_grabcut_b200.synthetic import geometric_sample, slic_like_labels
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handoff",
                             "seed_from_prior.npz"))
    cases = sorted({k.split("/")[0] for k in z.files})
    for name in cases:
        H, W, seed, nseg, n_nodes = (int(v) for v in z[f"{name}/meta"])
        img = geometric_sample(H, W, seed)[0]
        seg = slic_like_labels(H, W, nseg, seed)
        graph = gg.GraphBuilder(img, gg.SuperpixelGraphConfig(n_segments=nseg), segments=seg).build()
        assert graph.n_nodes == n_nodes
        np.testing.assert_allclose(graph.prior_features, z[f"{name}/prior"], rtol=1e-5, atol=2e-6)
        for tname in ("all_pr_bgd", "all_fgd", "no_fg", "no_bg", "mixed"):
            tri = z[f"{name}/{tname}/in"]
            for frac in (0.1, 0.5):
                got = gg.seed_from_prior(tri, graph, seed_frac=frac)
                assert got.dtype == np.uint8 and np.array_equal(got, z[f"{name}/{tname}/{frac}"]), (name, tname, frac)


def test_trimap_path_with_seeding(gg):
    """TrimapPath(seed_frac=0.1) == TrimapPath() followed by the oracle's seed_from_prior per
    image; thresholds of 2.0 make every trimap one-sided (no definite labels, and with the
    background posterior forced above the foreground one no foreground at all)."""
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port, trimap_port
    B, H, W, nseg = 5, 128, 160, 40
    imgs, labs = make_batch(B, H, W, nseg, seed0=300)
    state = model_port.random_state_dict(32, 2, seed=4)
    # bias the head towards background so that p_bg > p_fg everywhere: trimaps become all PR_BGD
    state = {k: v.clone() for k, v in state.items()}
    state["head.bias"] = state["head.bias"] + torch.tensor([8.0, 0.0, -8.0])
    cap = int(labs.max()) + 1
    plain = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=cap, threshold_fg=2.0, threshold_bg=2.0)
    tri0 = plain(imgs, labs)
    assert set(np.unique(tri0)) == {2}
    seeded = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=cap, threshold_fg=2.0, threshold_bg=2.0,
                           seed_frac=0.1)
    tri1 = seeded(imgs, labs)
    graphs = gg.build_graph_batch(imgs, labs, gg.SuperpixelGraphConfig()).to_graphs(labs)
    for b in range(B):
        want = trimap_port.seed_from_prior(tri0[b], graphs[b].prior_features, labs[b], graphs[b].n_nodes, 0.1)
        assert (want == 3).any() and np.array_equal(tri1[b], want)
    # a two-sided batch is left alone
    both = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=cap, seed_frac=0.1)
    ref = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=cap)
    st2 = model_port.random_state_dict(32, 2, seed=4)
    both = gg.TrimapPath(st2, gg.SuperpixelGraphConfig(), node_cap=cap, threshold_fg=0.34, threshold_bg=0.34, seed_frac=0.1)
    t_a = both(imgs, labs)
    ref = gg.TrimapPath(st2, gg.SuperpixelGraphConfig(), node_cap=cap, threshold_fg=0.34, threshold_bg=0.34)
    t_b = ref(imgs, labs)
    for b in range(B):
        want = trimap_port.seed_from_prior(t_b[b], graphs[b].prior_features, labs[b], graphs[b].n_nodes, 0.1)
        assert np.array_equal(t_a[b], want)


def test_trimap_path_uint16_label_transport(gg):
    """uint16 label maps on the host entry points (gg_path_config.label_bytes = 2) give the
    same trimaps and counts as the int32 maps of the reference layout."""
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port
    B, H, W = 9, 131, 150          # odd sizes: the widening kernel's tail path
    imgs, labs = make_batch(B, H, W, 50, seed0=500)
    state = model_port.random_state_dict(32, 2, seed=6)
    path = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=int(labs.max()) + 1, chunk=4)
    t32, n32, e32 = path(imgs, labs, return_counts=True)
    t16, n16, e16 = path(imgs, labs.astype(np.uint16), return_counts=True)
    assert np.array_equal(t32, t16) and np.array_equal(n32, n16) and np.array_equal(e32, e16)
    p = path.submit(imgs, labs.astype(np.uint16))
    assert np.array_equal(p.result(), t32)


# ----------------------------------------------------------------------------- training-data labels
def test_region_labels_vs_reference_golden(gg):
    """derive_trimap_labels / prepare_samples (gg_region_labels) against the reference's own
    dataset.py outputs (tests/golden/handoff/region_labels.npz): labels, fg_ratio and edge lists
    bit-exact, features within the feature tolerance; batched == per image."""
    import os
    from gcn_grabcut_b200.synthetic import geometric_sample, slic_like_labels
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handoff", "region_labels.npz"))
    cases = sorted({k.split("/")[0] for k in z.files})
    for name in cases:
        H, W, seed, nseg = (int(v) for v in z[f"{name}/meta"])
        img, mask = geometric_sample(H, W, seed)
        seg = slic_like_labels(H, W, nseg, seed)
        for key in [k for k in z.files if k.startswith(f"{name}/labels/")]:
            fg_thr, bg_thr = (float(v) for v in key.split("/")[-1].split("_"))
            got = gg.derive_trimap_labels(seg, mask, fg_thr, bg_thr)
            assert got.dtype == np.int64 and np.array_equal(got, z[key]), key
        data, y, segments = gg.prepare_sample({"image": img, "gt_mask": mask},
                                              gg.SuperpixelGraphConfig(n_segments=nseg), segments=seg)
        assert np.array_equal(segments, seg) and y is data.y
        assert data.y.dtype == torch.int64 and np.array_equal(data.y.numpy(), z[f"{name}/y"])
        assert data.fg_ratio.dtype == torch.float32 and np.array_equal(data.fg_ratio.numpy(), z[f"{name}/fg_ratio"])
        assert np.array_equal(data.edge_index.numpy(), z[f"{name}/edge_index"])
        np.testing.assert_allclose(data.x.numpy(), z[f"{name}/x"], rtol=FEAT_RTOL, atol=FEAT_ATOL)
        np.testing.assert_allclose(data.node_area.numpy(), z[f"{name}/node_area"], rtol=1e-6)
    # batched form, images of one shape: against the oracle per image
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import graph_port
    B, H, W, nseg = 7, 144, 176, 60
    imgs, labs = make_batch(B, H, W, nseg, seed0=900)
    masks = np.stack([geometric_sample(H, W, 900 + b)[1] for b in range(B)])
    res = gg.prepare_samples(imgs, masks, labs, gg.SuperpixelGraphConfig(n_segments=nseg))
    assert len(res) == B
    for b, (data, y, segm) in enumerate(res):
        want_y, want_r = graph_port.derive_trimap_labels(labs[b], masks[b], 0.70, 0.70)
        assert np.array_equal(y.numpy(), want_y) and np.array_equal(data.fg_ratio.numpy(), want_r)
        assert data.x.shape == (len(want_y), 19) and np.array_equal(segm, labs[b])


@pytest.mark.parametrize("D", [64, 256])
def test_resgcn_other_widths_tc_vs_simt_vs_oracle(gg, D):
    """Hidden widths other than 128: D = 256 runs the tcgen05 transforms as 2 x 2 blocks of 128
    (strided launches, the second K block accumulates), D = 64 as single 64-wide blocks; both
    against the SIMT fp32 transforms and the oracle."""
    from gcn_grabcut_b200 import _native as nat
    from oracle import model_port
    state = model_port.random_state_dict(D, 3, seed=8)
    net = gg.ResGCNNet(hidden_channels=D, n_layers=3)
    net.load_state_dict(state)
    net = net.to("cuda").eval()
    gen = torch.Generator().manual_seed(11)
    N = 700                                      # several 128-row tiles, a ragged last one
    x = torch.randn(N, 19, generator=gen)
    s = torch.randint(0, N, (4 * N,), generator=gen)
    d = torch.randint(0, N, (4 * N,), generator=gen)
    keep = s != d
    pairs = torch.unique(torch.stack([torch.minimum(s, d)[keep], torch.maximum(s, d)[keep]]), dim=1)
    ei = torch.cat([pairs, pairs.flip(0)], 1)
    ea = torch.rand(pairs.size(1), 5, generator=gen).repeat(2, 1)
    out = {}
    for impl in ("simt", "tc"):
        nat.handle(0).set_option("gemm_impl", 0 if impl == "simt" else 1)
        try:
            out[impl] = net(gg.Data(x=x, edge_index=ei, edge_attr=ea).to("cuda")).cpu()
        finally:
            nat.handle(0).set_option("gemm_impl", 1)
    ref = model_port.resgcn_forward(state, x, ei, ea, None)
    print(f"D={D}: tc-simt {float((out['tc'] - out['simt']).abs().max()):.3g}  tc-oracle "
          f"{float((out['tc'] - ref).abs().max()):.3g}")
    assert torch.allclose(out["tc"], out["simt"], atol=2e-4, rtol=1e-4)
    assert torch.allclose(out["tc"], ref, atol=5e-4, rtol=1e-4)


def test_streaming_sweep_is_chunking_invariant(gg):
    """Size-independent properties on a longer sweep (BASELINE config D in miniature: many
    batches streamed back to back): the trimap of an image does not depend on the batch it
    travels in, on its position in the batch, on the chunk size or on the number of batches in
    flight -- every reduction of the path is per image."""
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port
    H, W, nseg, n_img = 96 + 64, 128 + 64, 40, 96
    imgs, labs = make_batch(n_img, H, W, nseg, seed0=4000)
    state = model_port.random_state_dict(32, 2, seed=9)
    cap = int(labs.max()) + 1
    ref_path = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=cap, chunk=0)
    want = ref_path(imgs, labs)                                     # one call, default chunking
    rng = np.random.RandomState(0)
    perm = rng.permutation(n_img)
    path = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=cap, chunk=5)
    got = np.empty_like(want)
    pending, pos = [], 0
    for bs in (7, 1, 32, 13, 20, 23):                                # ragged batches of a shuffled order
        idx = perm[pos:pos + bs]
        pos += bs
        pending.append((idx, path.submit(imgs[idx], labs[idx])))
        if len(pending) > 3:
            i0, p0 = pending.pop(0)
            got[i0] = p0.result()
    for i0, p0 in pending:
        got[i0] = p0.result()
    assert pos == n_img and np.array_equal(got, want)


def test_models_share_the_device_handle(gg):
    """One device handle holds one set of weights: a ResGCNNet and two TrimapPaths with different
    weights can be used alternately, each reloads its own weights when another one used the handle."""
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port
    imgs, labs = make_batch(3, 128, 160, 40, seed0=77)
    cap = int(labs.max()) + 1
    s1, s2 = model_port.random_state_dict(32, 2, seed=1), model_port.random_state_dict(64, 3, seed=2)
    p1 = gg.TrimapPath(s1, gg.SuperpixelGraphConfig(), node_cap=cap)
    t1 = p1(imgs, labs)
    p2 = gg.TrimapPath(s2, gg.SuperpixelGraphConfig(), node_cap=cap)
    t2 = p2(imgs, labs)
    net = gg.ResGCNNet(hidden_channels=32, n_layers=2)
    net.load_state_dict(s1)
    net = net.to("cuda").eval()
    g = gg.GraphBuilder(imgs[0], gg.SuperpixelGraphConfig(), segments=labs[0]).build()
    data = gg.Data(x=torch.tensor(g.node_input()), edge_index=torch.tensor(g.edge_index),
                   edge_attr=torch.tensor(g.edge_attr)).to("cuda")
    probs = net.predict_probs(data)
    assert np.array_equal(p1(imgs, labs), t1) and np.array_equal(p2(imgs, labs), t2)
    assert np.array_equal(net.predict_probs(data), probs)
    assert np.array_equal(gg.refine_trimap(probs, labs[0], imgs[0]), t1[0])


def test_abi_error_paths(gg):
    """Misuse is reported through the status code and gg_last_error, never by a crash or a silent
    fallback: forward before gg_load_weights, bad shapes, a bad label width, NULL pointers,
    an unknown option, an unknown ticket."""
    import ctypes as C
    from gcn_grabcut_b200 import _native as nat
    L = nat.lib()
    hp = C.c_void_p()
    assert L.gg_create(C.byref(hp), 0) == 0
    try:
        img = torch.zeros((1, 16, 16, 3), dtype=torch.uint8)
        lab = torch.zeros((1, 16, 16), dtype=torch.int32)
        tri = torch.zeros((1, 16, 16), dtype=torch.uint8)
        pc = nat.PathConfig(nat.GraphConfig(4, 4, 8, 0), 8, 1e-3, 0.55, 0.55, 1, 0, 4, 0, 0.0)
        rc = L.gg_trimap_path_host(hp, nat.ptr(img), nat.ptr(lab), 1, 16, 16, C.byref(pc), nat.ptr(tri), None, None)
        assert rc == nat.GG_ERR_STATE and b"gg_load_weights" in L.gg_last_error()
        assert L.gg_set_option(hp, b"no_such_option", 1) == nat.GG_ERR_INVALID
        assert L.gg_trimap_path_host_wait(hp, 3) == nat.GG_ERR_INVALID
        assert L.gg_trimap_path_host(hp, None, nat.ptr(lab), 1, 16, 16, C.byref(pc), nat.ptr(tri), None, None) == nat.GG_ERR_INVALID
        out = (C.c_int64 * 4)()
        assert L.gg_selftest_math(hp, None) == nat.GG_ERR_INVALID and L.gg_selftest_math(hp, out) == 0
    finally:
        L.gg_destroy(hp)
    # through the Python mirror: shape / dtype mistakes raise ValueError like the reference's checks
    from oracle import model_port
    path = gg.TrimapPath(model_port.random_state_dict(32, 2, seed=0), gg.SuperpixelGraphConfig(), node_cap=8)
    with pytest.raises(ValueError):
        path(np.zeros((1, 16, 16), np.uint8), np.zeros((1, 16, 16), np.int32))
    with pytest.raises(ValueError):
        path(np.zeros((1, 16, 16, 3), np.uint8), np.zeros((1, 8, 16), np.int32))
    pc_bad = nat.PathConfig.from_buffer_copy(path.pc)
    pc_bad.label_bytes = 3
    img = torch.zeros((1, 16, 16, 3), dtype=torch.uint8)
    lab = torch.zeros((1, 16, 16), dtype=torch.int32)
    tri = torch.zeros((1, 16, 16), dtype=torch.uint8)
    assert nat.lib().gg_trimap_path_host(path.h.ptr, nat.ptr(img), nat.ptr(lab), 1, 16, 16, C.byref(pc_bad),
                                         nat.ptr(tri), None, None) == nat.GG_ERR_INVALID
    with pytest.raises(nat.NativeError):          # a label >= node_cap is a capacity error, not a wrong answer
        path(np.zeros((1, 16, 16, 3), np.uint8), np.full((1, 16, 16), 9, np.int32))
    assert path(np.zeros((2, 16, 16, 3), np.uint8), np.zeros((2, 16, 16), np.int32)).shape == (2, 16, 16)


def test_device_status_is_sticky_and_checked(gg):
    """The device-pointer entry points never synchronise, so overflow conditions surface through
    the sticky status word: (a) an edge index >= N given to forward(data) raises (the reference's
    scatter_add raises an index error) -- the bit set by gg_coo_to_csr must survive
    gg_resgcn_forward; (b) TrimapPath.run_device with a label >= node_cap: check=True /
    check_status() raise, and the word is cleared by the read so the next batch is clean;
    (c) in-place weight edits are picked up (tensor version counters)."""
    from gcn_grabcut_b200 import _native as nat
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port
    state = model_port.random_state_dict(32, 2, seed=3)
    net = gg.ResGCNNet(hidden_channels=32, n_layers=2)
    net.load_state_dict(state)
    net = net.to("cuda").eval()
    N = 40
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(N, 19, generator=gen)
    ei = torch.stack([torch.arange(N - 1), torch.arange(1, N)])
    ei = torch.cat([ei, ei.flip(0)], 1)
    ea = torch.rand(ei.shape[1], 5, generator=gen)
    good = net(gg.Data(x=x, edge_index=ei, edge_attr=ea).to("cuda")).cpu()
    bad_ei = ei.clone()
    bad_ei[0, 3] = N + 7
    with pytest.raises(nat.NativeError):
        net(gg.Data(x=x, edge_index=bad_ei, edge_attr=ea).to("cuda"))
    assert torch.equal(net(gg.Data(x=x, edge_index=ei, edge_attr=ea).to("cuda")).cpu(), good)
    # (c) in-place edit of a parameter: the device copy follows
    with torch.no_grad():
        net.head.bias.add_(1.0)
    moved = net(gg.Data(x=x, edge_index=ei, edge_attr=ea).to("cuda")).cpu()
    assert torch.allclose(moved, good + 1.0, atol=1e-5)
    # (b) whole path on device pointers
    imgs, labs = make_batch(20, 128, 160, 30, seed0=1)
    cap = int(labs.max()) + 1
    path = gg.TrimapPath(state, gg.SuperpixelGraphConfig(), node_cap=cap)
    it, lt = torch.from_numpy(imgs).cuda(), torch.from_numpy(labs).cuda()
    want = path.run_device(it, lt, check=True).cpu().numpy()
    bad = labs.copy()
    bad[17, 5, 5] = cap + 3                                   # lands in the second sub-batch
    path.run_device(it, torch.from_numpy(bad).cuda())         # no sync, no error yet
    path.run_device(it, lt)                                   # a clean batch after it: the bit stays
    with pytest.raises(nat.NativeError):
        path.check_status()
    path.check_status()                                       # read-and-clear: clean again
    assert np.array_equal(path.run_device(it, lt, check=True).cpu().numpy(), want)
    with pytest.raises(nat.NativeError):
        path.run_device(it, torch.from_numpy(bad).cuda(), check=True)


def test_reference_segment_through_the_drop_in(gg):
    """The boundary, exercised end to end: the reference's OWN ``GCNGrabCutPipeline.segment``
    (unmodified file from oracle/_ref, see oracle/make_ref.py) is run twice on the same seeded
    image, label map (SLIC shim hook) and checkpoint -- once as shipped (numpy / cv2 / torch on
    the CPU), once with the names swapped exactly as INTEGRATION.md §2 describes
    (GraphBuilder, SuperpixelGraphConfig, refine_trimap, guided_filter, _seed_from_prior,
    project_to_pixels, CLASS_*; the model is this repository's ResGCNNet loaded from the same
    state-dict).  Trimaps must agree pixel for pixel outside the tolerance band of a threshold,
    and the GrabCut masks computed from them by the unchanged cv2.grabCut must agree wherever
    the trimaps do."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference files not available (run python -m oracle.make_ref where /root/reference exists)")
    import gcn_grabcut_b200.pipeline as ggp
    from gcn_grabcut_b200.synthetic import geometric_sample, slic_like_labels
    from oracle import model_port, trimap_port
    ref = ref_loader.load()                                  # also puts the third-party shims on sys.path
    from skimage import segmentation as slic_shim            # the shim: returns the queued label map
    pm = ref.pipeline
    H, W, nseg, seed = 200, 264, 120, 21
    img = geometric_sample(H, W, seed)[0]
    seg = slic_like_labels(H, W, nseg, seed)
    state = model_port.random_state_dict(64, 3, seed=5)

    m_ref = ref.model.ResGCNNet(hidden_channels=64, n_layers=3)
    m_ref.load_state_dict(state)
    m_ref.eval()
    import cv2
    slic_shim.set_next_labels(seg)
    cv2.setRNGSeed(7)                      # cv2.grabCut seeds its GMMs with cv::kmeans, which draws from OpenCV's global RNG
    res_ref = pm.GCNGrabCutPipeline(m_ref, ref.graph_builder.SuperpixelGraphConfig(n_segments=nseg),
                                    device="cpu").segment(img)

    import gcn_grabcut_b200.graph_builder as ggb
    saved_backend = ggb.SLIC_BACKEND
    ggb.SLIC_BACKEND = "skimage"           # the same label map for both runs, through the shim's slic()
    saved = {k: getattr(pm, k) for k in ("GraphBuilder", "SuperpixelGraphConfig", "refine_trimap", "guided_filter",
                                          "_seed_from_prior", "project_to_pixels", "CLASS_BG", "CLASS_FG")}
    try:
        pm.GraphBuilder, pm.SuperpixelGraphConfig = gg.GraphBuilder, gg.SuperpixelGraphConfig
        pm.refine_trimap, pm.guided_filter = gg.refine_trimap, gg.guided_filter
        pm._seed_from_prior = ggp._seed_from_prior
        pm.project_to_pixels, pm.CLASS_BG, pm.CLASS_FG = gg.project_to_pixels, gg.CLASS_BG, gg.CLASS_FG
        m_gg = gg.ResGCNNet(hidden_channels=64, n_layers=3)
        m_gg.load_state_dict(state)
        m_gg.eval()
        slic_shim.set_next_labels(seg)
        cv2.setRNGSeed(7)
        res_gg = pm.GCNGrabCutPipeline(m_gg, gg.SuperpixelGraphConfig(n_segments=nseg), device="cuda").segment(img)
    finally:
        ggb.SLIC_BACKEND = saved_backend
        for k, v in saved.items():
            setattr(pm, k, v)
    assert set(res_gg.timing) == set(res_ref.timing) == {"graph_build", "data_prep", "gcn_inference", "grabcut",
                                                         "postprocess"}
    assert np.array_equal(res_gg.segments, res_ref.segments)
    assert res_gg.trimap.shape == (H, W) and res_gg.trimap.dtype == np.uint8
    # tolerance band: pixels whose filtered posterior sits within 2e-4 of a decision boundary
    g = gg.GraphBuilder(img, gg.SuperpixelGraphConfig(n_segments=nseg), segments=seg).build()
    probs = model_port.predict_probs(state, torch.tensor(g.node_input()), torch.tensor(g.edge_index),
                                     torch.tensor(g.edge_attr))
    _, pbg, pfg = trimap_port.refine_trimap(probs, seg, img, return_planes=True)
    near = trimap_port.near_threshold_mask(pbg, pfg, 0.55, 0.55, 2e-4)
    bad = int(((res_gg.trimap != res_ref.trimap) & ~near).sum())
    print(f"reference segment(): CPU as shipped vs drop-in on the GPU: {bad} trimap pixels differ outside the band "
          f"({int(near.sum())} inside); masks equal: {np.array_equal(res_gg.binary_mask, res_ref.binary_mask)}")
    assert bad == 0
    if np.array_equal(res_gg.trimap, res_ref.trimap):
        assert np.array_equal(res_gg.binary_mask, res_ref.binary_mask)


def test_fused_gcn_blocks_vs_layerwise_and_oracle(gg):
    """The per-graph fused residual-GCN kernel (gcn_fused.cu: LN -> tcgen05 transform -> x' kept in
    TMEM / shared memory -> CSR gather -> bias, gate, GELU, residual, JK, all layers in one launch)
    against the layer-wise kernels and the oracle: single graphs through forward(data) (N = 37, 128,
    129, 300, 384: one to three 128-row tiles, ragged and exact), isolated nodes and self loops, and
    the whole batched path (posteriors of every image)."""
    from gcn_grabcut_b200 import _native as nat
    from gcn_grabcut_b200.synthetic import make_batch
    from oracle import model_port
    h = nat.handle(0)
    state = model_port.random_state_dict(128, 6, seed=12)
    net = gg.ResGCNNet(hidden_channels=128, n_layers=6)
    net.load_state_dict(state)
    net = net.to("cuda").eval()
    gen = torch.Generator().manual_seed(3)
    for N in (37, 128, 129, 300, 384):
        x = torch.randn(N, 19, generator=gen)
        s = torch.randint(0, N, (5 * N,), generator=gen)
        d = torch.randint(0, N, (5 * N,), generator=gen)
        keep = (s != d) & (s != 3) & (d != 3)                   # node 3 stays isolated
        pairs = torch.unique(torch.stack([torch.minimum(s, d)[keep], torch.maximum(s, d)[keep]]), dim=1)
        ei = torch.cat([pairs, pairs.flip(0), torch.tensor([[5, 9], [5, 9]])], 1)   # two self loops
        ea = torch.rand(ei.size(1), 5, generator=gen)
        data = gg.Data(x=x, edge_index=ei, edge_attr=ea).to("cuda")
        out = {}
        for fused in (2, 0):
            h.set_option("gcn_fused", fused)
            try:
                l0 = h.launches()
                out[fused] = net(data).cpu()
                out[fused, "launches"] = h.launches() - l0
            finally:
                h.set_option("gcn_fused", 1)
        ref = model_port.resgcn_forward(state, x, ei, ea, None)
        assert out[2, "launches"] < out[0, "launches"], "the fused kernel did not run"
        print(f"N={N}: fused-layerwise {float((out[2] - out[0]).abs().max()):.3g}, fused-oracle "
              f"{float((out[2] - ref).abs().max()):.3g}; launches {out[2, 'launches']} vs {out[0, 'launches']}")
        assert torch.allclose(out[2], out[0], atol=5e-5, rtol=1e-5)
        assert torch.allclose(out[2], ref, atol=POST_ATOL, rtol=1e-4)
    # batched path: posteriors of every image
    imgs, labs = make_batch(12, 160, 192, 60, seed0=33)
    cap = int(labs.max()) + 1
    path = gg.TrimapPath(state, gg.SuperpixelGraphConfig(n_segments=60), node_cap=cap)
    it, lt = torch.from_numpy(imgs).cuda(), torch.from_numpy(labs).cuda()
    res = {}
    for fused in (2, 0):
        h.set_option("gcn_fused", fused)
        try:
            probs = torch.zeros(12 * cap, 3, device="cuda")
            noff = torch.zeros(13, dtype=torch.int64, device="cuda")
            tri = path.run_device(it, lt, probs_t=probs, node_off_t=noff, check=True)
            res[fused] = (tri.cpu().numpy(), probs.cpu().numpy(), noff.cpu().numpy())
        finally:
            h.set_option("gcn_fused", 1)
    assert np.array_equal(res[2][2], res[0][2])
    nt = int(res[2][2][-1])
    assert np.abs(res[2][1][:nt] - res[0][1][:nt]).max() < 2e-5
    assert np.mean(res[2][0] != res[0][0]) < 1e-4


def test_clean_mask_and_grabcut_guards_vs_reference_golden(gg):
    """gg_clean_masks / gg_grabcut_guards against the reference's own clean_mask and
    GrabCut.run_with_trimap guards (tests/golden/handoff/clean_and_guards.npz): bit-exact, single
    images and a batch; plus random masks against the oracle (cv2.connectedComponentsWithStats)."""
    import os
    from oracle import trimap_port
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handoff", "clean_and_guards.npz"))
    n = 0
    for key in z.files:
        parts = key.split("/")
        if parts[0] == "mask" and parts[2] != "in":
            ratio, largest = parts[2].split("_")
            got = gg.clean_mask(z[f"mask/{parts[1]}/in"].copy(), float(ratio), bool(int(largest)))
            assert got.dtype == np.uint8 and np.array_equal(got, z[key]), key
            n += 1
        if parts[0] == "tri" and parts[2] == "out":
            got, deg = gg.grabcut_guards(z[f"tri/{parts[1]}/in"])
            assert np.array_equal(got, z[key]) and deg == bool(z[f"tri/{parts[1]}/degenerate"]), key
            n += 1
    assert n >= 40
    rng = np.random.RandomState(5)
    batch = (rng.rand(6, 200, 264) < np.array([0.02, 0.2, 0.45, 0.55, 0.7, 0.0])[:, None, None]).astype(np.uint8)
    for ratio, largest in ((0.002, False), (0.0005, False), (0.3, False), (0.002, True)):
        got = gg.clean_mask(batch, ratio, largest)
        for b in range(6):
            want = trimap_port.clean_mask(batch[b].copy(), ratio, largest)
            assert np.array_equal(got[b], want), (b, ratio, largest)
    tri = rng.randint(0, 4, (5, 64, 80)).astype(np.uint8)
    tri[1][tri[1] == 1] = 3
    tri[2][tri[2] == 0] = 2
    tri[3][:] = 2
    tri[4][np.isin(tri[4], (1, 3))] = 0
    got, deg = gg.grabcut_guards(tri)
    for b in range(5):
        want, wd = trimap_port.grabcut_guards(tri[b])
        assert np.array_equal(got[b], want) and bool(deg[b]) == wd, b


def test_prepare_dataset_cache_roundtrip(gg, tmp_path):
    """prepare_dataset: misses are built in batches and written as <key>.pt blobs
    {"data", "segments"}; a second call is served from the cache without any kernel launch; a
    corrupt entry is rebuilt."""
    from gcn_grabcut_b200 import _native as nat
    from gcn_grabcut_b200.synthetic import geometric_sample, slic_like_labels
    samples, segs = [], []
    for i, (H, W) in enumerate(((128, 160), (128, 160), (144, 176), (128, 160), (144, 176))):
        img, mask = geometric_sample(H, W, 50 + i)
        samples.append({"image": img, "gt_mask": mask})
        segs.append(slic_like_labels(H, W, 40, 50 + i))
    cfg = gg.SuperpixelGraphConfig(n_segments=40)
    first = gg.prepare_dataset(samples, cfg, 0.7, 0.7, cache_dir=tmp_path, segments=segs, batch_size=2)
    files = sorted(p.name for p in tmp_path.iterdir())
    assert files == sorted(gg.cache_key(s, cfg, 0.7, 0.7) + ".pt" for s in samples)
    h = nat.handle(0)
    l0 = h.launches()
    second = gg.prepare_dataset(samples, cfg, 0.7, 0.7, cache_dir=tmp_path, segments=segs)
    assert h.launches() == l0, "a cache hit must not touch the GPU"
    for (d1, y1, s1), (d2, y2, s2), seg in zip(first, second, segs):
        assert torch.equal(d1.x, d2.x) and torch.equal(d1.edge_index, d2.edge_index) and torch.equal(y1, y2)
        assert torch.equal(d1.fg_ratio, d2.fg_ratio) and np.array_equal(s1, seg) and np.array_equal(s2, seg)
    blob = torch.load(tmp_path / files[0], map_location="cpu", weights_only=False)
    assert set(blob) == {"data", "segments"}
    (tmp_path / files[0]).write_bytes(b"truncated")
    third = gg.prepare_dataset(samples, cfg, 0.7, 0.7, cache_dir=tmp_path, segments=segs, keep_segments=False)
    assert h.launches() > l0 and all(t[2] is None for t in third)
    for (d1, _, _), (d3, _, _) in zip(first, third):
        assert torch.equal(d1.x, d3.x)


def test_compute_auto_prior_reference_signature(gg):
    """compute_auto_prior(segments, lab, centre_sigma, contrast_sigma) with a caller-supplied float32
    Lab plane -- the reference's own call shape (graph_builder.py:357-362) -- against the oracle,
    default and non-default sigmas; and the image= form against the builder's prior."""
    from gcn_grabcut_b200.synthetic import geometric_sample, slic_like_labels
    from oracle import graph_port
    for (H, W, nseg, seed) in ((160, 192, 48, 3), (97, 131, 20, 8), (320, 480, 300, 1)):
        img = geometric_sample(max(H, 128), max(W, 128), seed)[0][:H, :W]
        seg = slic_like_labels(H, W, nseg, seed)
        lab = graph_port.pixel_planes(np.ascontiguousarray(img))["lab"]
        for cs, ks in ((0.45, 0.40), (0.30, 0.25), (0.8, 1.1)):
            want = graph_port.auto_prior(seg, lab, cs, ks)
            got = gg.compute_auto_prior(seg, lab, cs, ks)
            assert got.dtype == np.float32 and got.shape == want.shape
            np.testing.assert_allclose(got, want, rtol=FEAT_RTOL, atol=FEAT_ATOL)
        via_image = gg.compute_auto_prior(seg, image=np.ascontiguousarray(img))
        np.testing.assert_allclose(via_image, graph_port.auto_prior(seg, lab), rtol=FEAT_RTOL, atol=FEAT_ATOL)


def test_slic_quality_vs_oracle(gg):
    """GPU SLIC (gg_slic: the reference's skimage.segmentation.slic call restated, graph_builder.py:
    177-188) against the numpy restatement of scikit-image's algorithm (oracle/slic_port.py).
    Label-for-label parity with scikit-image is unpinned (no build of it can run here), so the gate
    is (a) the invariants the reference tests -- labels contiguous 0..N-1, every label used
    (tests/test.py:112-117) -- plus 4-connected regions of at least half a nominal superpixel, and
    (b) segmentation quality on the generator's ground-truth masks: boundary recall and
    under-segmentation error no worse than the oracle's by more than a small margin."""
    from scipy import ndimage as ndi
    from gcn_grabcut_b200.synthetic import geometric_sample
    from oracle import graph_port, slic_port
    cases = [(160, 192, 48, 3), (200, 264, 120, 7), (320, 480, 300, 1), (320, 480, 300, 2)]
    br_g, br_o, ue_g, ue_o = [], [], [], []
    for (H, W, nseg, seed) in cases:
        img, mask = geometric_sample(H, W, seed)
        cfg = gg.SuperpixelGraphConfig(n_segments=nseg)
        lab, cnt = gg.slic_labels(img[None], cfg, return_counts=True)
        lab, n = lab[0], int(cnt[0])
        assert lab.dtype == np.int32 and lab.shape == (H, W)
        assert lab.min() == 0 and lab.max() == n - 1 and len(np.unique(lab)) == n, "labels must be 0..N-1, all used"
        sy, ty, sx, tx = slic_port.regular_grid_2d(H, W, nseg)
        nominal = len(range(sy, H, ty)) * len(range(sx, W, tx))
        assert 0.8 * nominal <= n <= 1.3 * nominal, (n, nominal)
        # every label is one 4-connected region; none is smaller than min_size (the first may be)
        sizes = np.bincount(lab.ravel())
        min_size = int(0.5 * H * W / nominal)
        assert (sizes[1:] >= min_size).all(), (sizes.min(), min_size)
        comp = 0
        for v in range(n):
            comp += ndi.label(lab == v)[1]
        assert comp == n, f"{comp} connected components for {n} labels"
        oracle = slic_port.slic(graph_port.pixel_planes(img)["lab"], nseg)
        br_g.append(slic_port.boundary_recall(lab, mask)); br_o.append(slic_port.boundary_recall(oracle, mask))
        ue_g.append(slic_port.undersegmentation_error(lab, mask)); ue_o.append(slic_port.undersegmentation_error(oracle, mask))
        same = float(np.mean(slic_port.boundary_map(lab) == slic_port.boundary_map(oracle)))
        print(f"{H}x{W} n={nseg}: GPU N={n} oracle N={int(oracle.max()) + 1}; boundary recall {br_g[-1]:.3f} vs "
              f"{br_o[-1]:.3f}; under-segmentation {ue_g[-1]:.5f} vs {ue_o[-1]:.5f}; boundary maps agree on {same:.3f}")
    assert np.mean(br_g) >= np.mean(br_o) - 0.03
    assert np.mean(ue_g) <= np.mean(ue_o) + 0.002


def test_slic_kernel_variants_and_odd_sizes(gg, tmp_path):
    """The SLIC kernels have shape-dependent variants: column-walk assignment with 64- or 32-row tiles
    (feature rows staged by asynchronous copies), the row-interleaved form for very small grid steps, and
    the connectivity passes with four pixels per thread when H*W % 4 == 0.  (a) Sizes that cut tiles and
    take the scalar connectivity kernels keep the invariants; (b) the alternative kernels, selected through
    the library's A/B environment switches in a fresh process, give the same segmentation: the scalar
    connectivity passes bit for bit, the row-interleaved assignment up to float32 rounding of near ties."""
    import os, subprocess, sys
    from scipy import ndimage as ndi
    from gcn_grabcut_b200.synthetic import geometric_sample
    from oracle import slic_port
    for (H, W, nseg, seed) in [(150, 203, 60, 5), (130, 171, 40, 9), (128, 320, 100, 2)]:
        img, _ = geometric_sample(H, W, seed)
        lab, cnt = gg.slic_labels(img[None], gg.SuperpixelGraphConfig(n_segments=nseg), return_counts=True)
        lab, n = lab[0], int(cnt[0])
        assert lab.min() == 0 and lab.max() == n - 1 and len(np.unique(lab)) == n, (H, W)
        assert sum(ndi.label(lab == v)[1] for v in range(n)) == n, f"{H}x{W}: labels are not 4-connected"
    imgs = np.stack([geometric_sample(320, 480, 60 + i)[0] for i in range(3)])
    np.save(tmp_path / "imgs.npy", imgs)
    base = gg.slic_labels(imgs, gg.SuperpixelGraphConfig(n_segments=300))
    script = ("import sys, numpy as np; sys.path.insert(0, %r); import gcn_grabcut_b200 as gg; "
              "imgs = np.load(sys.argv[1]); np.save(sys.argv[2], gg.slic_labels(imgs, gg.SuperpixelGraphConfig(n_segments=300)))"
              % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    for name, env, exact in [("cc_scalar", {"GG_SLIC_CC_SCALAR": "1"}, True), ("walk32", {"GG_SLIC_WALK": "32"}, True),
                             ("row_interleaved", {"GG_SLIC_WALK": "0"}, False),
                             # the column-walk kernels' own general path (taken when a centre drifted so far that
                             # its window no longer covers the tile), forced for every tile
                             ("walk_general_path", {"GG_SLIC_GENERAL": "1"}, False)]:
        out = tmp_path / f"{name}.npy"
        subprocess.run([sys.executable, "-c", script, str(tmp_path / "imgs.npy"), str(out)], check=True,
                       env={**os.environ, **env}, timeout=300)
        other = np.load(out)
        if exact:
            assert np.array_equal(other, base), name
        else:
            agree = np.mean([np.mean(slic_port.boundary_map(a) == slic_port.boundary_map(b)) for a, b in zip(other, base)])
            assert agree >= 0.995, (name, agree)


def test_slic_in_the_path(gg):
    """Superpixels produced on the device inside the whole-path entry points (only the images cross
    PCIe) == gg_slic followed by the path on those label maps; GraphBuilder(image, cfg).build()
    works without a label map; determinism (the centre sums are integers)."""
    from gcn_grabcut_b200.synthetic import geometric_sample
    from oracle import model_port
    B, H, W, nseg = 6, 160, 192, 48
    imgs = np.stack([geometric_sample(H, W, 40 + i)[0] for i in range(B)])
    cfg = gg.SuperpixelGraphConfig(n_segments=nseg)
    labs, cnt = gg.slic_labels(imgs, cfg, return_counts=True)
    labs2 = gg.slic_labels(imgs, cfg)
    assert np.array_equal(labs, labs2), "SLIC must be deterministic"
    state = model_port.random_state_dict(32, 2, seed=3)
    cap = int(cnt.max()) + 8
    with_labels = gg.TrimapPath(state, cfg, node_cap=cap)(imgs, labs)
    path = gg.TrimapPath(state, cfg, node_cap=cap, device_slic=True, chunk=4)
    tri, nn, ne = path(imgs, return_counts=True)
    assert np.array_equal(nn, cnt) and np.array_equal(tri, with_labels)
    td = path.run_device(torch.from_numpy(imgs).cuda(), check=True)
    assert np.array_equal(td.cpu().numpy(), with_labels)
    g = gg.GraphBuilder(imgs[0], cfg).build()
    assert g.n_nodes == int(cnt[0]) and np.array_equal(g.segments, labs[0])
    ref = gg.GraphBuilder(imgs[0], cfg, segments=labs[0]).build()
    assert np.array_equal(g.edge_index, ref.edge_index) and np.array_equal(g.node_features, ref.node_features)
