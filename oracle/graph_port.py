"""
CPU restatement of the reference graph construction (label map -> attributed region graph).

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for the CUDA path and the CPU
baseline that bench.py times.  Never imported by the product package.

Follows /root/reference/src/gcn_grabcut/graph_builder.py; every function cites the lines
it restates.  It is written stage by stage (pixel planes -> region sums -> node
attributes -> adjacency -> non-local pairs -> edge attributes -> prior) and keeps every
intermediate, so that a GPU mismatch can be localised to one stage.  dtype behaviour that
the reference gets implicitly from numpy-2 promotion rules is made explicit here.

Pinned against the reference itself: tests/golden/*.npz were produced by running the
unmodified reference files (oracle/ref_loader.py) and tests/test_oracle_golden.py checks
this port against them bit for bit.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import cv2
import numpy as np

from .thirdparty import find_boundaries, rgb2hsv, rgb2lab

N_IMAGE_FEATS = 16          # graph_builder.py:73
N_PRIOR_FEATS = 3           # graph_builder.py:74
N_NODE_FEATS = 19           # graph_builder.py:76
N_EDGE_FEATS = 5            # graph_builder.py:77

F32 = np.float32


@dataclass
class RegionGraph:
    """Field-for-field mirror of SuperpixelGraph (graph_builder.py:80-91) + intermediates."""
    segments: np.ndarray
    node_features: np.ndarray
    edge_index: np.ndarray
    edge_attr: np.ndarray
    n_nodes: int
    n_edges: int
    node_centroids: np.ndarray
    prior_features: np.ndarray
    node_areas: np.ndarray
    stages: Dict[str, np.ndarray] = field(default_factory=dict)

    def node_input(self) -> np.ndarray:                      # graph_builder.py:93-98
        return np.concatenate([self.node_features, self.prior_features], axis=1).astype(F32)


# --------------------------------------------------------------------------- pixel planes

def pixel_planes(bgr: np.ndarray) -> Dict[str, np.ndarray]:
    """graph_builder.py:142-154 -- Lab / HSV (float64 math -> float32), gray, Sobel magnitude."""
    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    lab = rgb2lab(rgb).astype(F32)
    hsv = rgb2hsv(rgb).astype(F32)
    gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY).astype(F32)
    gx = cv2.Sobel(gray, cv2.CV_32F, 1, 0, ksize=3)
    gy = cv2.Sobel(gray, cv2.CV_32F, 0, 1, ksize=3)
    grad = np.sqrt(gx ** 2 + gy ** 2)
    return {"lab": lab, "hsv": hsv, "gray": gray, "grad": grad}


# --------------------------------------------------------------------------- region sums

def _segsum(flat: np.ndarray, weights: np.ndarray, n: int) -> np.ndarray:
    """np.bincount with weights: float64 sequential accumulation, then float32 (…:197-199)."""
    return np.bincount(flat, weights=weights.ravel(), minlength=n).astype(F32)


def region_statistics(seg: np.ndarray, planes: Dict[str, np.ndarray], n: int) -> Dict[str, np.ndarray]:
    """graph_builder.py:190-226."""
    H, W = seg.shape
    flat = seg.ravel()
    lab, hsv, grad = planes["lab"], planes["hsv"], planes["grad"]

    counts = np.bincount(flat, minlength=n).astype(F32)
    safe = np.maximum(counts, F32(1.0))

    mean_lab = np.stack([_segsum(flat, lab[:, :, c], n) for c in range(3)], 1) / safe[:, None]
    sq_lab = np.stack([_segsum(flat, lab[:, :, c] ** 2, n) for c in range(3)], 1) / safe[:, None]
    std_lab = np.sqrt(np.maximum(sq_lab - mean_lab ** 2, F32(0.0)))
    mean_hsv = np.stack([_segsum(flat, hsv[:, :, c], n) for c in range(3)], 1) / safe[:, None]

    yy, xx = np.mgrid[0:H, 0:W]
    cy = _segsum(flat, yy.astype(F32) / F32(H), n) / safe      # float32 coordinates (:207)
    cx = _segsum(flat, xx.astype(F32) / F32(W), n) / safe
    centroids = np.stack([cy, cx], 1).astype(F32)

    boundary_px = _segsum(flat, find_boundaries(seg, mode="inner").astype(F32), n)

    gmax_eps = F32(grad.max()) + F32(1e-6)                      # float32 + weak python float (:214)
    grad_scaled = grad / gmax_eps
    return {
        "counts": counts, "safe": safe,
        "area_ratio": (counts / F32(float(H * W))).astype(F32),
        "mean_lab": mean_lab.astype(F32), "std_lab": std_lab.astype(F32),
        "mean_hsv": mean_hsv.astype(F32), "centroids": centroids,
        "boundary_px": boundary_px,
        "mean_grad": (_segsum(flat, grad, n) / safe).astype(F32),
        "mean_grad_n": (_segsum(flat, grad_scaled, n) / safe).astype(F32),
    }


# --------------------------------------------------------------------------- node attributes

def node_features(st: Dict[str, np.ndarray]) -> np.ndarray:
    """graph_builder.py:228-255 -- the 16 image-derived columns of x_i."""
    n = st["counts"].shape[0]
    f = np.zeros((n, N_IMAGE_FEATS), dtype=F32)
    f[:, 0:3] = st["mean_lab"]
    f[:, 3:6] = st["std_lab"]
    f[:, 6:9] = st["mean_hsv"]
    f[:, 9] = st["centroids"][:, 0]
    f[:, 10] = st["centroids"][:, 1]
    f[:, 11] = st["area_ratio"]
    perim = np.maximum(st["boundary_px"], F32(1.0))
    f[:, 12] = np.clip((F32(4 * np.pi) * st["counts"]) / (perim ** 2), F32(0.0), F32(1.0))
    f[:, 13] = st["mean_grad"] / F32(255.0)
    f[:, 14] = st["boundary_px"] / st["safe"]
    d = st["centroids"] - F32(0.5)
    f[:, 15] = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) / F32(0.707)
    for lo in (0, 3):                                           # per-image min-max (:250-253)
        col = f[:, lo:lo + 3]
        mn, mx = col.min(0), col.max(0)
        f[:, lo:lo + 3] = (col - mn) / (mx - mn + F32(1e-6))
    return np.nan_to_num(f, nan=0.0, posinf=1.0, neginf=0.0)


# --------------------------------------------------------------------------- adjacency

def adjacency_pairs(seg: np.ndarray, n: int, connectivity: int = 4):
    """graph_builder.py:265-286 -- sorted unique (lo,hi) pairs and shared-boundary counts."""
    views = [(seg[:, :-1], seg[:, 1:]), (seg[:-1, :], seg[1:, :])]
    if connectivity == 8:
        views += [(seg[:-1, :-1], seg[1:, 1:]), (seg[:-1, 1:], seg[1:, :-1])]
    a = np.concatenate([v[0].ravel() for v in views])
    b = np.concatenate([v[1].ravel() for v in views])
    keep = a != b
    a, b = a[keep], b[keep]
    lo, hi = np.minimum(a, b).astype(np.int64), np.maximum(a, b).astype(np.int64)
    codes, cnt = np.unique(lo * n + hi, return_counts=True)
    pairs = np.stack([codes // n, codes % n], 1)
    return pairs, cnt


def _adjacency_csr(adj: np.ndarray, n: int):
    """Symmetric adjacency as CSR (indptr, indices) -- the row-blocked form of the reference's
    N x N boolean mask (graph_builder.py:337-340)."""
    a = np.concatenate([adj[:, 0], adj[:, 1]]).astype(np.int64)
    b = np.concatenate([adj[:, 1], adj[:, 0]]).astype(np.int64)
    order = np.argsort(a, kind="stable")
    a, b = a[order], b[order]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, a + 1, 1)
    return np.cumsum(indptr), b


def nonlocal_pairs(adj: np.ndarray, mean_lab: np.ndarray, n: int, k: int,
                   tie_break: str = "argpartition", block: int = 512):
    """
    graph_builder.py:324-350 -- k nearest neighbours in mean-Lab space, spatially adjacent
    pairs excluded, symmetrised.  Returns (pairs, n_ties): ``n_ties`` counts rows whose
    k-th and (k+1)-th candidate distances are equal -- there np.argpartition's choice is
    implementation-defined and a different (valid) selection is possible.

    The reference materialises the N x N distance matrix (400 MB at N = 10^4); every row of
    it is independent, so the same arithmetic is evaluated here in blocks of ``block`` rows
    (identical float32 operations per element, identical result).

    ``tie_break``: "argpartition" = the reference's call (whatever numpy's introselect picks
    among equal distances); "lower_index" = the rule the CUDA path documents (equal
    distances -> the lower region index first), i.e. a stable sort on (distance, index).
    Both agree whenever ``n_ties == 0``.
    """
    indptr, indices = _adjacency_csr(adj, n)
    kth = min(k, n - 1) - 1
    codes, n_ties = [], 0
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        diff = mean_lab[r0:r1, None, :] - mean_lab[None, :, :]
        sq = diff * diff
        d = np.sqrt((sq[:, :, 0] + sq[:, :, 1]) + sq[:, :, 2])     # float32, left-to-right (:334)
        rr = np.arange(r0, r1)
        d[rr - r0, rr] = np.inf                                     # fill_diagonal
        reps = indptr[r0 + 1:r1 + 1] - indptr[r0:r1]
        d[np.repeat(rr - r0, reps), indices[indptr[r0]:indptr[r1]]] = np.inf      # adjacent pairs
        if tie_break == "lower_index":
            nbrs = np.argsort(d, axis=1, kind="stable")[:, :k]
        else:
            nbrs = np.argpartition(d, kth=kth, axis=1)[:, :k]
        rows = np.repeat(rr, nbrs.shape[1])
        cols = nbrs.ravel()
        ok = np.isfinite(d[rows - r0, cols])
        rows, cols = rows[ok], cols[ok]
        lo, hi = np.minimum(rows, cols).astype(np.int64), np.maximum(rows, cols).astype(np.int64)
        codes.append(lo * n + hi)
        if n > k:
            part = np.partition(d, (k - 1, k), axis=1)
            n_ties += int(np.sum((part[:, k - 1] == part[:, k]) & np.isfinite(part[:, k - 1])))
    codes = np.unique(np.concatenate(codes)) if codes else np.zeros(0, np.int64)
    return np.stack([codes // n, codes % n], 1), n_ties


def pair_features(pairs, st, shared, flag) -> np.ndarray:
    """graph_builder.py:309-322 -- e_ij (5-d), normalisers taken over THIS pair set."""
    i, j = pairs[:, 0], pairs[:, 1]
    dl = st["mean_lab"][i] - st["mean_lab"][j]
    de = np.sqrt((dl[:, 0] * dl[:, 0] + dl[:, 1] * dl[:, 1]) + dl[:, 2] * dl[:, 2])
    de = de / (de.max() + F32(1e-6))
    dc = st["centroids"][i] - st["centroids"][j]
    dxy = np.sqrt(dc[:, 0] * dc[:, 0] + dc[:, 1] * dc[:, 1])
    dxy = dxy / (dxy.max() + F32(1e-6))
    gc = np.abs(st["mean_grad_n"][i] - st["mean_grad_n"][j])
    return np.stack([de, dxy, shared, gc, flag], axis=1).astype(F32)


def compute_edges(seg, st, n, connectivity=4, n_nonlocal=4, tie_break="argpartition"):
    """graph_builder.py:257-307 -- COO order [adj | nl | adj reversed | nl reversed]."""
    adj, cnt = adjacency_pairs(seg, n, connectivity)
    shared = cnt.astype(F32) / (np.float64(cnt.max()) + 1e-6)   # float64 division (:286)
    attr = pair_features(adj, st, shared, np.zeros(len(adj), F32))
    pairs, n_ties = adj, 0
    nl = np.zeros((0, 2), dtype=np.int64)
    if n_nonlocal > 0 and n > n_nonlocal + 1:
        nl, n_ties = nonlocal_pairs(adj, st["mean_lab"], n, int(n_nonlocal), tie_break)
        if len(nl):
            nl_attr = pair_features(nl, st, np.zeros(len(nl), F32), np.ones(len(nl), F32))
            pairs = np.concatenate([adj, nl], 0)
            attr = np.concatenate([attr, nl_attr], 0)
    src = np.concatenate([pairs[:, 0], pairs[:, 1]])
    dst = np.concatenate([pairs[:, 1], pairs[:, 0]])
    edge_index = np.stack([src, dst], 0).astype(np.int64)
    edge_attr = np.concatenate([attr, attr], 0).astype(F32)
    return edge_index, edge_attr, {"adj_pairs": adj, "adj_counts": cnt, "nl_pairs": nl,
                                   "knn_ties": np.int64(n_ties)}


# --------------------------------------------------------------------------- prior

def _unit_norm(v: np.ndarray) -> np.ndarray:
    """graph_builder.py:447-454."""
    v = v.astype(F32)
    mn, mx = float(v.min()), float(v.max())
    if mx - mn < 1e-8:
        return np.zeros_like(v)
    return (v - F32(mn)) / F32(mx - mn)


def auto_prior(seg: np.ndarray, lab: np.ndarray, centre_sigma: float = 0.45,
               contrast_sigma: float = 0.40, return_stages: bool = False, block: int = 512):
    """graph_builder.py:357-444 -- [fg-ness, bg-ness, ambiguity] per region."""
    H, W = seg.shape
    n = int(seg.max()) + 1
    flat = seg.ravel()
    counts = np.bincount(flat, minlength=n).astype(F32)
    safe = np.maximum(counts, F32(1.0))
    mean_lab = np.stack([np.bincount(flat, weights=lab[:, :, c].ravel(), minlength=n)
                         for c in range(3)], axis=1).astype(F32) / safe[:, None]
    yy, xx = np.mgrid[0:H, 0:W]
    cy = np.bincount(flat, weights=(yy.ravel() / H), minlength=n) / safe      # float64 path (:401)
    cx = np.bincount(flat, weights=(xx.ravel() / W), minlength=n) / safe
    cen = np.stack([cy, cx], axis=1).astype(F32)

    # the two N x N matrices of :406-411, evaluated in row blocks (rows are independent; numpy
    # reduces each row of a C-contiguous float32 block with the same pairwise order whatever
    # the number of rows, so the blocked sum is bit-identical to the reference's full matrix)
    area_w = counts / F32(max(counts.sum(), 1.0))
    contrast = np.empty(n, dtype=F32)
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        dl = mean_lab[r0:r1, None, :] - mean_lab[None, :, :]
        colour_d = np.sqrt((dl[..., 0] * dl[..., 0] + dl[..., 1] * dl[..., 1]) + dl[..., 2] * dl[..., 2])
        dc = cen[r0:r1, None, :] - cen[None, :, :]
        spatial_d = np.sqrt(dc[..., 0] * dc[..., 0] + dc[..., 1] * dc[..., 1])
        spatial_w = np.exp(-(spatial_d ** 2) / F32(2 * contrast_sigma ** 2))
        contrast[r0:r1] = (colour_d * spatial_w * area_w[None, :]).sum(axis=1)
    contrast = _unit_norm(contrast)

    c0 = cen - F32(0.5)
    centre_d = np.sqrt(c0[:, 0] * c0[:, 0] + c0[:, 1] * c0[:, 1])
    centre_w = np.exp(-(centre_d ** 2) / F32(2 * centre_sigma ** 2))
    fg = _unit_norm(contrast * centre_w)

    border = np.concatenate([seg[0, :], seg[-1, :], seg[:, 0], seg[:, -1]])   # corners twice
    bcount = np.bincount(border, minlength=n).astype(F32)
    bratio = bcount / safe
    if bcount.sum() > 0:
        w_bg = bcount / bcount.sum()
        mu = (mean_lab * w_bg[:, None]).sum(axis=0)
        var = (((mean_lab - mu) ** 2) * w_bg[:, None]).sum(axis=0).sum()
        sigma = float(np.sqrt(max(var, 1e-6)))
        dm = mean_lab - mu
        d_bg = np.sqrt((dm[:, 0] * dm[:, 0] + dm[:, 1] * dm[:, 1]) + dm[:, 2] * dm[:, 2])
        bg = np.exp(-(d_bg ** 2) / F32(2 * (sigma + 1e-6) ** 2))
    else:
        bg = np.zeros(n, dtype=F32)
    bg = _unit_norm(np.maximum(bg, np.clip(bratio * F32(4.0), F32(0.0), F32(1.0))))
    amb = F32(1.0) - np.abs(fg - bg)
    prior = np.nan_to_num(np.stack([fg, bg, amb], axis=1).astype(F32),
                          nan=0.0, posinf=1.0, neginf=0.0)
    if return_stages:
        return prior, {"prior_centroids": cen, "prior_contrast": contrast,
                       "border_count": bcount}
    return prior


# --------------------------------------------------------------------------- driver

def build_graph(bgr: np.ndarray, seg: np.ndarray, connectivity: int = 4,
                n_nonlocal: int = 4, keep_stages: bool = True,
                tie_break: str = "argpartition") -> RegionGraph:
    """GraphBuilder(image, cfg).build() with the label map supplied (graph_builder.py:156-175).
    ``tie_break``: see nonlocal_pairs."""
    seg = np.ascontiguousarray(seg, dtype=np.int32)
    n = int(seg.max()) + 1
    planes = pixel_planes(bgr)
    st = region_statistics(seg, planes, n)
    feats = node_features(st)
    edge_index, edge_attr, est = compute_edges(seg, st, n, connectivity, n_nonlocal, tie_break)
    prior, pst = auto_prior(seg, planes["lab"], return_stages=True)
    stages: Dict[str, np.ndarray] = {}
    if keep_stages:
        stages.update(st)
        stages.update(est)
        stages.update(pst)
        stages["gray"] = planes["gray"]
        stages["grad_max"] = np.float32(planes["grad"].max())
    return RegionGraph(
        segments=seg, node_features=feats.astype(F32), edge_index=edge_index,
        edge_attr=edge_attr, n_nodes=n, n_edges=edge_index.shape[1],
        node_centroids=st["centroids"], prior_features=prior,
        node_areas=st["area_ratio"], stages=stages)


# ----------------------------------------------------------------------------- training labels
def derive_trimap_labels(segments: np.ndarray, gt_mask: np.ndarray, fg_threshold: float = 0.75,
                         bg_threshold: float = 0.75):
    """dataset.py:175-205 (labels) and :239-249 (fg_ratio of prepare_sample): per-region foreground
    coverage by exact counts; returns (labels int64 (N,), fg_ratio float32 (N,))."""
    n_nodes = int(segments.max()) + 1
    flat = segments.ravel()
    counts = np.bincount(flat, minlength=n_nodes).astype(np.float64)
    fg_sum = np.bincount(flat, weights=(gt_mask.ravel() > 0).astype(np.float64), minlength=n_nodes)
    fg_ratio = fg_sum / np.maximum(counts, 1.0)
    labels = np.full(n_nodes, 1, dtype=np.int64)
    labels[fg_ratio >= fg_threshold] = 2
    labels[fg_ratio <= 1 - bg_threshold] = 0
    labels[counts == 0] = 1
    return labels, fg_ratio.astype(np.float32)
