"""
CPU restatement of ``skimage.segmentation.slic`` as the reference calls it
(/root/reference/src/gcn_grabcut/graph_builder.py:177-188):

    slic(lab_float32, n_segments, compactness=10, sigma=1, start_label=0, channel_axis=-1)

TEST INFRASTRUCTURE (see oracle/__init__.py).  scikit-image is a third-party dependency of the
reference that is absent from /root/reference and not installable here (pyproject.toml:10-23 lists
the bare name ``scikit-image``, version unpinned); this file restates the published algorithm of
scikit-image 0.19-0.22 (``skimage/segmentation/slic_superpixels.py`` and ``_slic.pyx``):

  1. ``img_as_float``; global min-max rescale of the whole array to [0, 1];
  2. Gaussian smoothing of the two spatial axes (``scipy.ndimage.gaussian_filter``: radius
     ``int(4 sigma + 0.5)``, ``mode="reflect"``), applied when ``sigma > 0``;
  3. a 3-channel input is converted with ``rgb2lab`` (``convert2lab`` defaults to true for three
     channels -- the reference passes an image that is ALREADY CIELAB, so its rescaled L, a, b are
     read as R, G, B and converted once more; the quirk is part of the behaviour and is kept);
  4. cluster centres on ``regular_grid(shape, n_segments)``; features scaled by ``1/compactness``;
  5. ``max_num_iter = 10`` rounds of k-means: every centre claims the pixels of its
     ``(4 step + 1)^2`` window whose distance ``|dc|^2 + |dxy|^2 / step^2`` it strictly improves
     (centres in index order, so ties go to the lower index), then centres = feature means;
  6. ``_enforce_label_connectivity_cython``: raster-order breadth-first search over 4-connected
     components of equal label (cut at ``max_size``); components smaller than ``min_size`` take
     the label of the last previously labelled neighbour seen; surviving components are numbered
     consecutively from ``start_label``.

PARITY UNPINNED against a real scikit-image build (none can be run here); the CUDA SLIC is gated on
segmentation quality against this restatement (boundary recall / under-segmentation error on the
generator's ground-truth masks) and on the invariants the reference tests (labels 0..N-1, all
used), not on label-for-label equality -- see tests/test_gpu_parity.py::test_slic_quality.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage as ndi

from .thirdparty import rgb2lab

try:                                    # the connectivity pass is a sequential flood fill
    import numba
    _njit = numba.njit(cache=False)
except Exception:                       # pragma: no cover
    def _njit(f):
        return f


def regular_grid_2d(height: int, width: int, n_points: int):
    """skimage.util.regular_grid for the (1, H, W) volume slic builds: (start_y, step_y, start_x, step_x)."""
    shape = np.array([1, height, width], dtype=float)
    order = np.argsort(shape)
    sorted_dims = shape[order]
    space = float(np.prod(shape))
    steps = np.full(3, (space / n_points) ** (1.0 / 3))
    if (sorted_dims < steps).any():
        for dim in range(3):
            steps[dim] = sorted_dims[dim]
            space = float(np.prod(sorted_dims[dim + 1:]))
            steps[dim + 1:] = (space / n_points) ** (1.0 / (3 - dim - 1)) if dim < 2 else steps[dim + 1:]
            if (sorted_dims >= steps).all():
                break
    starts = (steps // 2).astype(int)
    steps_i = np.round(steps).astype(int)
    unsort = np.argsort(order)
    starts, steps_i = starts[unsort], steps_i[unsort]
    return int(starts[1]), max(int(steps_i[1]), 1), int(starts[2]), max(int(steps_i[2]), 1)


def preprocess(image: np.ndarray, compactness: float, sigma: float) -> np.ndarray:
    """Steps 1-4 (feature image): float32 (H, W, 3)."""
    img = np.array(image, dtype=np.float32, copy=True)
    img -= img.min()
    mx = img.max()
    if mx != 0:
        img /= mx
    if sigma > 0:
        img = ndi.gaussian_filter(img, [sigma, sigma, 0], mode="reflect")
    if img.shape[-1] == 3:
        img = rgb2lab(img).astype(np.float32)
    return np.ascontiguousarray(img * np.float32(1.0 / compactness), dtype=np.float32)


def kmeans(feat: np.ndarray, n_segments: int, max_num_iter: int = 10):
    """Step 5.  Returns (nearest (H,W) int64, centres (K, 5) [y, x, c0, c1, c2], step)."""
    H, W, C = feat.shape
    sy, ty, sx, tx = regular_grid_2d(H, W, n_segments)
    ys, xs = np.arange(sy, H, ty), np.arange(sx, W, tx)
    cy, cx = np.meshgrid(ys, xs, indexing="ij")
    K = cy.size
    cen = np.zeros((K, 2 + C), dtype=np.float32)
    cen[:, 0], cen[:, 1] = cy.ravel(), cx.ravel()
    step = max(ty, tx)
    w_sp = np.float32(1.0 / (step * step))
    nearest = np.zeros((H, W), dtype=np.int64)
    yy, xx = np.mgrid[0:H, 0:W]
    for _ in range(max_num_iter):
        dist = np.full((H, W), np.finfo(np.float32).max, dtype=np.float32)
        changed = False
        for k in range(K):
            y0 = int(max(cen[k, 0] - 2 * ty, 0)); y1 = int(min(cen[k, 0] + 2 * ty + 1, H))
            x0 = int(max(cen[k, 1] - 2 * tx, 0)); x1 = int(min(cen[k, 1] + 2 * tx + 1, W))
            if y1 <= y0 or x1 <= x0 or not np.isfinite(cen[k]).all():
                continue
            dy = (cen[k, 0] - yy[y0:y1, x0:x1].astype(np.float32))
            dx = (cen[k, 1] - xx[y0:y1, x0:x1].astype(np.float32))
            d = (dy * dy + dx * dx) * w_sp
            dc = feat[y0:y1, x0:x1] - cen[k, 2:]
            d = d + (dc * dc).sum(-1)
            win = dist[y0:y1, x0:x1]
            better = win > d
            if better.any():
                changed = True
                win[better] = d[better]
                nearest[y0:y1, x0:x1][better] = k
        if not changed:
            break
        flat = nearest.ravel()
        n = np.bincount(flat, minlength=K).astype(np.float32)
        new = np.zeros_like(cen)
        new[:, 0] = np.bincount(flat, weights=yy.ravel(), minlength=K)
        new[:, 1] = np.bincount(flat, weights=xx.ravel(), minlength=K)
        for c in range(C):
            new[:, 2 + c] = np.bincount(flat, weights=feat[:, :, c].ravel(), minlength=K)
        with np.errstate(invalid="ignore", divide="ignore"):
            cen = (new / n[:, None]).astype(np.float32)
    return nearest, cen, step


@_njit
def _enforce_connectivity(seg, min_size, max_size, start_label):
    H, W = seg.shape
    mask_label = start_label - 1
    out = np.full((H, W), mask_label, dtype=np.int64)
    ddx = np.array([1, -1, 0, 0]); ddy = np.array([0, 0, 1, -1])
    coords = np.empty((max(max_size, 1), 2), dtype=np.int64)
    cur = start_label
    for y in range(H):
        for x in range(W):
            if seg[y, x] == mask_label or out[y, x] > mask_label:
                continue
            adjacent = 0
            label = seg[y, x]
            out[y, x] = cur
            size = 1
            visited = 0
            coords[0, 0] = y; coords[0, 1] = x
            while visited < size and size < max_size:
                for i in range(4):
                    yy = coords[visited, 0] + ddy[i]; xx = coords[visited, 1] + ddx[i]
                    if 0 <= yy < H and 0 <= xx < W:
                        if seg[yy, xx] == label and out[yy, xx] == mask_label:
                            out[yy, xx] = cur
                            coords[size, 0] = yy; coords[size, 1] = xx
                            size += 1
                            if size >= max_size:
                                break
                        elif out[yy, xx] > mask_label and out[yy, xx] != cur:
                            adjacent = out[yy, xx]
                visited += 1
            if size < min_size:
                for i in range(size):
                    out[coords[i, 0], coords[i, 1]] = adjacent
            else:
                cur += 1
    return out


def slic(image: np.ndarray, n_segments: int = 300, compactness: float = 10.0, sigma: float = 1.0,
         max_num_iter: int = 10, start_label: int = 0, min_size_factor: float = 0.5,
         max_size_factor: float = 3.0) -> np.ndarray:
    """The reference's slic call on a (H, W, 3) float image -> int64 (H, W) labels."""
    feat = preprocess(image, compactness, sigma)
    nearest, cen, _ = kmeans(feat, n_segments, max_num_iter)
    H, W = nearest.shape
    seg_size = H * W / cen.shape[0]
    labels = _enforce_connectivity(nearest + start_label, int(min_size_factor * seg_size),
                                   int(max_size_factor * seg_size), start_label)
    return labels


# ----------------------------------------------------------------------------- quality measures
def boundary_map(labels: np.ndarray) -> np.ndarray:
    b = np.zeros(labels.shape, dtype=bool)
    b[:, :-1] |= labels[:, :-1] != labels[:, 1:]
    b[:-1, :] |= labels[:-1, :] != labels[1:, :]
    return b


def boundary_recall(labels: np.ndarray, gt_mask: np.ndarray, tol: int = 2) -> float:
    """Fraction of ground-truth boundary pixels with a superpixel boundary within ``tol`` pixels."""
    gt_b = boundary_map(gt_mask.astype(np.int64))
    if not gt_b.any():
        return 1.0
    sp_b = ndi.binary_dilation(boundary_map(labels), iterations=tol)
    return float((gt_b & sp_b).sum() / gt_b.sum())


def undersegmentation_error(labels: np.ndarray, gt_mask: np.ndarray) -> float:
    """Min-based under-segmentation error (Neubert & Protzel): leakage of superpixels across the
    ground-truth regions, as a fraction of the image."""
    n = int(labels.max()) + 1
    gt = (gt_mask > 0).astype(np.int64)
    inside = np.bincount(labels.ravel(), weights=gt.ravel(), minlength=n)
    total = np.bincount(labels.ravel(), minlength=n).astype(np.float64)
    return float(np.minimum(inside, total - inside).sum() / labels.size)
