"""
CPU: the oracle port against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Integer outputs bit-exact; float outputs bit-exact where
the port uses the same numpy/cv2 calls, else within the stated tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import graph_port, model_port, trimap_port
from helpers import golden_cases, golden_inputs, load_golden

CASES = golden_cases()


def test_fixtures_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("name", CASES)
def test_graph_port_matches_reference(name):
    g = load_golden(name)
    img, seg = golden_inputs(g, name)
    out = graph_port.build_graph(img, seg, int(g["connectivity"]), int(g["n_nonlocal"]))
    assert out.n_nodes == int(g["n_nodes"]) and out.n_edges == int(g["n_edges"])
    assert np.array_equal(out.edge_index, g["edge_index"])                 # bit-exact structure
    assert out.edge_index.dtype == np.int64
    for key, mine in (("node_features", out.node_features), ("edge_attr", out.edge_attr),
                      ("prior_features", out.prior_features),
                      ("node_centroids", out.node_centroids), ("node_areas", out.node_areas)):
        assert mine.dtype == np.float32
        assert np.array_equal(mine, g[key]), f"{key}: max|d|={np.abs(mine - g[key]).max()}"


@pytest.mark.parametrize("name", CASES)
def test_model_port_matches_reference(name):
    g = load_golden(name)
    state = model_port.random_state_dict(int(g["hidden"]), int(g["n_layers"]), seed=int(g["seed"]))
    x = torch.tensor(np.concatenate([g["node_features"], g["prior_features"]], 1))
    ei = torch.tensor(g["edge_index"])
    ea = torch.tensor(g["edge_attr"])
    logits = model_port.resgcn_forward(state, x, ei, ea).numpy()
    # same torch ops in a different order of residual bookkeeping: fp32 tolerance
    np.testing.assert_allclose(logits, g["logits"], rtol=1e-5, atol=2e-6)
    probs = model_port.predict_probs(state, x, ei, ea)
    np.testing.assert_allclose(probs, g["probs"], rtol=1e-5, atol=1e-6)

    # batched: [graph, permuted graph, graph]
    n = x.shape[0]
    perm = torch.tensor(g["perm"])
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n)
    xb = torch.cat([x, x[perm], x])
    eib = torch.cat([ei, inv[ei] + n, ei + 2 * n], 1)
    eab = torch.cat([ea, ea, ea])
    batch = torch.arange(3).repeat_interleave(n)
    lb = model_port.resgcn_forward(state, xb, eib, eab, batch).numpy()
    np.testing.assert_allclose(lb, g["logits_batched"], rtol=1e-5, atol=2e-6)
    # the reference's own invariant (tests/test.py:294-306): batched == one by one @1e-4
    np.testing.assert_allclose(lb[:n], logits, atol=1e-4)
    np.testing.assert_allclose(lb[n:2 * n], logits[perm.numpy()], atol=1e-4)


@pytest.mark.parametrize("name", CASES)
def test_trimap_port_matches_reference(name):
    g = load_golden(name)
    img, seg = golden_inputs(g, name)
    tri = trimap_port.refine_trimap(g["probs"], seg, img, 0.55, 0.55, radius=8)
    assert tri.dtype == np.uint8 and np.array_equal(tri, g["trimap_refined"])
    tri4 = trimap_port.refine_trimap(g["probs"], seg, img, 0.4, 0.45, radius=4, eps=1e-2)
    assert np.array_equal(tri4, g["trimap_r4"])
    direct = model_port.probs_to_trimap(g["probs"], seg, 0.55, 0.55)
    assert np.array_equal(direct, g["trimap_direct"])
    assert set(np.unique(tri)).issubset({0, 1, 2, 3})


@pytest.mark.parametrize("name", CASES[:3])
def test_box_mean_restatement_matches_cv2(name):
    """cv2.blur(float32) == float64 window sum * 1/k^2 -> float32."""
    import cv2
    g = load_golden(name)
    img, seg = golden_inputs(g, name)
    guide = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY).astype(np.float32) / np.float32(255.0)
    for r in (1, 4, 8):
        if 2 * r + 1 > min(guide.shape):
            continue
        ref = cv2.blur(guide, (2 * r + 1, 2 * r + 1))
        mine = trimap_port.box_mean_f64(guide, r)
        frac = np.mean(ref == mine)
        assert frac > 0.999, f"r={r}: only {frac:.4f} of pixels bit-identical"
        np.testing.assert_allclose(mine, ref, rtol=0, atol=1e-7)


def test_thirdparty_colour_known_answers():
    """Known CIELAB / HSV values of sRGB primaries (D65, 2 deg)."""
    from oracle.thirdparty import rgb2lab, rgb2hsv
    px = np.array([[[255, 255, 255], [0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255],
                    [128, 128, 128]]], dtype=np.uint8)
    lab = rgb2lab(px)[0]
    np.testing.assert_allclose(lab[0], [100.0, 0.0, 0.0], atol=2e-2)
    np.testing.assert_allclose(lab[1], [0.0, 0.0, 0.0], atol=1e-9)
    np.testing.assert_allclose(lab[2], [53.24, 80.09, 67.20], atol=2e-2)
    np.testing.assert_allclose(lab[3], [87.73, -86.18, 83.18], atol=2e-2)
    np.testing.assert_allclose(lab[4], [32.30, 79.19, -107.86], atol=2e-2)
    np.testing.assert_allclose(lab[5][0], 53.585, atol=2e-2)
    hsv = rgb2hsv(px)[0]
    np.testing.assert_allclose(hsv[2], [0.0, 1.0, 1.0])
    np.testing.assert_allclose(hsv[3], [1 / 3, 1.0, 1.0])
    np.testing.assert_allclose(hsv[4], [2 / 3, 1.0, 1.0])
    np.testing.assert_allclose(hsv[5], [0.0, 0.0, 128 / 255])


def test_find_boundaries_label0_quirk():
    """mode='inner' treats label 0 as background: region 0 never owns boundary pixels."""
    from oracle.thirdparty import find_boundaries
    seg = np.zeros((6, 8), np.int32)
    seg[:, 4:] = 1
    b = find_boundaries(seg, mode="inner")
    assert not b[:, :4].any() and b[:, 4].all() and not b[:, 5:].any()


# ----------------------------------------------------------------------------- trimap hand-off
def _seed_fixture():
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handoff", "seed_from_prior.npz")
    z = np.load(path)
    cases = sorted({k.split("/")[0] for k in z.files})
    return z, cases


def test_seed_from_prior_port_matches_reference():
    """oracle.trimap_port.seed_from_prior against the reference's _seed_from_prior
    (tests/golden/make_golden_seed.py): one-sided trimaps are repaired identically, two-sided
    ones come back untouched."""
    from gcn_grabcut_b200.synthetic import slic_like_labels
    z, cases = _seed_fixture()
    assert len(cases) == 3
    n_checked = 0
    for name in cases:
        H, W, seed, nseg, n_nodes = (int(v) for v in z[f"{name}/meta"])
        seg = slic_like_labels(H, W, nseg, seed)
        prior = z[f"{name}/prior"]
        for tname in ("all_pr_bgd", "all_fgd", "no_fg", "no_bg", "mixed"):
            tri = z[f"{name}/{tname}/in"]
            for frac in (0.1, 0.5):
                want = z[f"{name}/{tname}/{frac}"]
                got = trimap_port.seed_from_prior(tri, prior, seg, n_nodes, frac)
                assert np.array_equal(got, want), (name, tname, frac)
                if tname == "mixed":
                    assert np.array_equal(got, tri)
                else:
                    assert not np.array_equal(got, tri)
                n_checked += 1
    assert n_checked == 30


def test_region_labels_port_matches_reference():
    """oracle.graph_port.derive_trimap_labels against the reference's derive_trimap_labels and
    prepare_sample outputs (tests/golden/make_golden_labels.py): labels and fg_ratio bit-exact."""
    import os
    from gcn_grabcut_b200.synthetic import geometric_sample, slic_like_labels
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handoff", "region_labels.npz"))
    cases = sorted({k.split("/")[0] for k in z.files})
    assert len(cases) == 3
    for name in cases:
        H, W, seed, nseg = (int(v) for v in z[f"{name}/meta"])
        mask = geometric_sample(H, W, seed)[1]
        seg = slic_like_labels(H, W, nseg, seed)
        for key in [k for k in z.files if k.startswith(f"{name}/labels/")]:
            fg_thr, bg_thr = (float(v) for v in key.split("/")[-1].split("_"))
            labels, _ = graph_port.derive_trimap_labels(seg, mask, fg_thr, bg_thr)
            assert labels.dtype == np.int64 and np.array_equal(labels, z[key]), key
        labels, ratio = graph_port.derive_trimap_labels(seg, mask, 0.70, 0.70)
        assert np.array_equal(labels, z[f"{name}/y"]) and np.array_equal(ratio, z[f"{name}/fg_ratio"])
        assert set(np.unique(labels)) <= {0, 1, 2} and len(np.unique(labels)) >= 2


def test_handoff_port_vs_reference_golden():
    """clean_mask / GrabCut guards: the oracle restatement against the reference's own outputs
    (tests/golden/handoff/clean_and_guards.npz, written by make_golden_handoff2.py)."""
    import os
    from oracle import trimap_port
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "handoff", "clean_and_guards.npz"))
    n = 0
    for key in z.files:
        parts = key.split("/")
        if parts[0] == "mask" and parts[2] != "in":
            ratio, largest = parts[2].split("_")
            got = trimap_port.clean_mask(z[f"mask/{parts[1]}/in"].copy(), float(ratio), bool(int(largest)))
            assert np.array_equal(got, z[key]), key
            n += 1
        if parts[0] == "tri" and parts[2] == "out":
            got, deg = trimap_port.grabcut_guards(z[f"tri/{parts[1]}/in"])
            assert np.array_equal(got, z[key]) and deg == bool(z[f"tri/{parts[1]}/degenerate"]), key
            n += 1
    assert n >= 40


def test_cache_key_known_answers():
    """The graph-cache key (reference dataset.py:363-377): values produced by the reference's own
    `_cache_key` for seeded samples (checked live against the reference when it is available)."""
    from gcn_grabcut_b200.dataset import cache_key
    from gcn_grabcut_b200.graph_builder import SuperpixelGraphConfig
    rng = np.random.RandomState(0)
    smp = {"image": rng.randint(0, 255, (40, 50, 3)).astype(np.uint8), "gt_mask": (rng.rand(40, 50) > 0.5).astype(np.uint8)}
    assert cache_key(smp, SuperpixelGraphConfig(), 0.7, 0.65) == "e4b39acfa81334ab16aa"
    assert cache_key(smp, SuperpixelGraphConfig(n_segments=77, n_nonlocal=0, connectivity=8), 0.7, 0.65) == "ae6ed2459bf60c50ce3d"
    lazy = {"image_path": "a/b.png", "mask_path": "a/b_m.png", "max_size": 512, "aug_seed": 3}
    from oracle import ref_loader
    if ref_loader.available():
        import importlib
        ref_loader.load()
        rd = importlib.import_module("gcn_grabcut.dataset")
        assert rd._cache_key(lazy, None, 0.7, 0.7) == cache_key(lazy, None, 0.7, 0.7)
        assert rd._cache_key(smp, None, 0.75, 0.75) == cache_key(smp, None, 0.75, 0.75)


# ----------------------------------------------------------------------------- GCNTrimapNet / GATTrimapNet
def variant_cases():
    import os
    from helpers import GOLDEN_DIR
    z = dict(np.load(os.path.join(GOLDEN_DIR, "variants", "reference.npz")))
    tags = sorted({k.split("/")[0] for k in z})
    return z, tags


def variant_state(z, tag):
    D, n, H, seed = (int(v) for v in z[f"{tag}/meta"])
    if tag.startswith("gcn"):
        state = model_port.random_gcn_trimap_state(D, n, seed=seed)
    else:
        state = model_port.random_gat_trimap_state(D, H, n, seed=seed)
    stored = {k[len(tag) + 7:]: v for k, v in z.items() if k.startswith(tag + "/state/")}
    for k, v in stored.items():                # the small cases pin the seeded generator itself
        assert np.array_equal(state[k].numpy(), v), f"{tag}: seeded state drifted at {k}"
    return state


@pytest.mark.parametrize("tag", variant_cases()[1])
def test_variant_ports_match_reference(tag):
    """oracle GCNTrimapNet / GATTrimapNet restatements against logits of the unmodified reference classes
    (tests/golden/make_golden_variants.py)."""
    z, _ = variant_cases()
    state = variant_state(z, tag)
    x, ei, ea = torch.tensor(z[f"{tag}/x"]), torch.tensor(z[f"{tag}/edge_index"]), torch.tensor(z[f"{tag}/edge_attr"])
    batch = torch.tensor(z[f"{tag}/batch"]) if f"{tag}/batch" in z else None
    if tag.startswith("gcn"):
        mine = model_port.gcn_trimap_forward(state, x, ei, ea)
    else:
        mine = model_port.gat_trimap_forward(state, x, ei, ea, batch)
    np.testing.assert_allclose(mine.numpy(), z[f"{tag}/logits"], atol=2e-5)


def test_slic_port_properties():
    """oracle/slic_port.py (the numpy restatement of skimage.segmentation.slic that gates the CUDA SLIC) has no
    golden vectors -- scikit-image cannot run here, parity with it is unpinned -- so the restatement is held
    to the properties scikit-image documents and the reference tests (tests/test.py:112-117): labels
    contiguous from start_label = 0 with every label used, 4-connected segments none of which is smaller than
    min_size = half a nominal superpixel (the first may absorb nothing and stay small), a segment count close
    to the grid, determinism, and superpixel boundaries that follow the object boundary of a two-region image."""
    from scipy import ndimage as ndi
    from gcn_grabcut_b200.synthetic import geometric_sample
    from oracle import graph_port, slic_port
    H, W, nseg = 120, 160, 40
    img, mask = geometric_sample(H, W, 4)
    lab_img = graph_port.pixel_planes(img)["lab"]
    seg = slic_port.slic(lab_img, nseg)
    assert seg.shape == (H, W) and np.issubdtype(seg.dtype, np.integer)
    n = int(seg.max()) + 1
    assert seg.min() == 0 and len(np.unique(seg)) == n
    sy, ty, sx, tx = slic_port.regular_grid_2d(H, W, nseg)
    nominal = len(range(sy, H, ty)) * len(range(sx, W, tx))
    assert 0.7 * nominal <= n <= 1.3 * nominal, (n, nominal)
    assert sum(ndi.label(seg == v)[1] for v in range(n)) == n, "every label must be one 4-connected region"
    sizes = np.bincount(seg.ravel())
    assert (sizes[1:] >= int(0.5 * H * W / nominal)).all()
    assert np.array_equal(seg, slic_port.slic(lab_img, nseg)), "deterministic"
    assert slic_port.boundary_recall(seg, mask) >= 0.95
    assert slic_port.undersegmentation_error(seg, mask) <= 0.03
    # the quality measures themselves: a segmentation equal to the mask is perfect, one blind to it is not
    assert slic_port.boundary_recall(mask.astype(np.int32), mask) == 1.0
    assert slic_port.undersegmentation_error(mask.astype(np.int32), mask) == 0.0
    assert slic_port.boundary_recall(np.zeros((H, W), np.int32), mask) == 0.0
