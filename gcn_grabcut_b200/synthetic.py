"""
Synthetic inputs for the trimap path: benchmark images and label maps.

These are INPUT PRODUCERS, not part of the accelerated path (the path starts at the
label map, BASELINE.json north_star), so they run on the host with numpy/cv2.

* ``geometric_sample``  -- seeded restatement of the reference's benchmark image
  generator ``GeometricDataset.sample`` (parametric_geom_dataset.py:54-69): one filled
  circle or 3..8-gon on a black canvas plus integer uniform noise.  The reference
  draws from the global ``random`` / ``np.random`` state; here both streams are
  explicit (``random.Random(seed)``, ``np.random.RandomState(seed)``), which yields
  the same image as ``random.seed(seed); np.random.seed(seed); ds.sample()``.
* ``slic_like_labels``  -- deterministic stand-in for ``skimage.segmentation.slic``
  (graph_builder.py:177-188; scikit-image is not installable here): a jittered-grid
  Voronoi partition with contiguous labels 0..N-1 in raster order of the cells, every
  label non-empty, regions compact and of SLIC-like size.
"""

from __future__ import annotations

import random as _random

import cv2
import numpy as np


def geometric_sample(height: int, width: int, seed: int, scale: float = 1.0):
    """
    One synthetic BGR uint8 image (H, W, 3) and its binary mask (H, W).

    ``scale`` multiplies the reference's pixel constants (centre margin 60, radius
    20..70); 1.0 reproduces parametric_geom_dataset.py exactly and is what configs
    A/B/D use.  Configs C/E pass ``min(H, W) / 320`` so that the object keeps its
    relative size on the larger canvases.
    """
    rng = _random.Random(seed)
    nrng = np.random.RandomState(seed)
    H, W = int(height), int(width)
    m = int(round(60 * scale))

    shape_type = rng.choice(["circle", "polygon"])
    center = (rng.randint(m, W - m), rng.randint(m, H - m))
    radius = rng.randint(int(round(20 * scale)), int(round(70 * scale)))
    color = tuple(int(rng.uniform(50, 255)) for _ in range(3))

    img = np.zeros((H, W, 3), dtype=np.uint8)
    mask = np.zeros((H, W), dtype=np.uint8)
    if shape_type == "circle":
        cv2.circle(img, center, radius, color, -1)
        cv2.circle(mask, center, radius, 1, -1)
    else:
        nv = rng.randint(3, 8)
        rot = rng.uniform(0, np.pi)
        angles = np.linspace(0, 2 * np.pi, nv, endpoint=False) + rot
        pts = np.vstack([center[0] + radius * np.cos(angles),
                         center[1] + radius * np.sin(angles)]).T.astype(np.int32)
        cv2.fillPoly(img, [pts], color)
        cv2.fillPoly(mask, [pts], 1)

    level = rng.randint(3, 20)
    noise = nrng.randint(-level, level, img.shape, dtype=np.int16)
    img = np.clip(img.astype(np.int16) + noise, 0, 255).astype(np.uint8)
    return img, mask


def grid_shape(height: int, width: int, n_segments: int):
    """Cell grid (ny, nx) with ny*nx ~ n_segments and roughly square cells."""
    ny = max(1, int(round(np.sqrt(n_segments * height / float(width)))))
    nx = max(1, int(round(n_segments / float(ny))))
    return min(ny, height), min(nx, width)


def slic_like_labels(height: int, width: int, n_segments: int, seed: int = 0,
                     jitter: float = 0.35) -> np.ndarray:
    """
    Jittered-grid Voronoi label map, int32 (H, W), labels 0..ny*nx-1, all used.

    Each grid cell owns one site, displaced from the cell centre by at most
    ``jitter`` cells; a pixel takes the label of its nearest site (ties -> lower
    label).  Sites sit on distinct integer pixels, so every label owns at least its
    own site pixel.
    """
    H, W = int(height), int(width)
    ny, nx = grid_shape(H, W, n_segments)
    rs = np.random.RandomState(seed)
    ch, cw = H / ny, W / nx
    jy = (rs.rand(ny, nx) * 2 - 1) * jitter * ch
    jx = (rs.rand(ny, nx) * 2 - 1) * jitter * cw
    sy = (np.arange(ny)[:, None] + 0.5) * ch + jy
    sx = (np.arange(nx)[None, :] + 0.5) * cw + jx
    # snap to distinct pixels inside the owning cell
    sy = np.clip(np.floor(sy), np.floor(np.arange(ny) * ch)[:, None],
                 np.ceil((np.arange(ny) + 1) * ch)[:, None] - 1).clip(0, H - 1)
    sx = np.clip(np.floor(sx), np.floor(np.arange(nx) * cw)[None, :],
                 np.ceil((np.arange(nx) + 1) * cw)[None, :] - 1).clip(0, W - 1)

    yy = np.arange(H, dtype=np.float32)[:, None]
    xx = np.arange(W, dtype=np.float32)[None, :]
    cy = np.minimum((np.arange(H) / ch).astype(np.int64), ny - 1)[:, None]
    cx = np.minimum((np.arange(W) / cw).astype(np.int64), nx - 1)[None, :]

    best_d = np.full((H, W), np.inf, dtype=np.float32)
    best_l = np.zeros((H, W), dtype=np.int32)
    syf, sxf = sy.astype(np.float32), sx.astype(np.float32)
    for dy in (-2, -1, 0, 1, 2):
        for dx in (-2, -1, 0, 1, 2):
            ny_i = np.clip(cy + dy, 0, ny - 1)
            nx_i = np.clip(cx + dx, 0, nx - 1)
            ny_b = np.broadcast_to(ny_i, (H, W))
            nx_b = np.broadcast_to(nx_i, (H, W))
            d = (yy - syf[ny_b, nx_b]) ** 2 + (xx - sxf[ny_b, nx_b]) ** 2
            lab = (ny_b * nx + nx_b).astype(np.int32)
            better = (d < best_d) | ((d == best_d) & (lab < best_l))
            best_d = np.where(better, d, best_d)
            best_l = np.where(better, lab, best_l)
    return np.ascontiguousarray(best_l, dtype=np.int32)


def make_batch(n_images: int, height: int, width: int, n_segments: int,
               seed0: int = 0, scale: float = 1.0):
    """(B,H,W,3) uint8 BGR images and (B,H,W) int32 label maps, seeds seed0..seed0+B-1."""
    imgs = np.empty((n_images, height, width, 3), dtype=np.uint8)
    labs = np.empty((n_images, height, width), dtype=np.int32)
    for i in range(n_images):
        imgs[i], _ = geometric_sample(height, width, seed0 + i, scale)
        labs[i] = slic_like_labels(height, width, n_segments, seed0 + i)
    return imgs, labs


# ----------------------------------------------------------------------------- weights
N_PRIOR_FEATS = 3


def random_state_dict(hidden: int = 128, n_layers: int = 6, seed: int = 0,
                      in_channels: int = 19, edge_channels: int = 5, n_classes: int = 3,
                      randomize_norms: bool = True) -> dict:
    """
    A random-init ResGCNNet state-dict with the reference's keys and shapes
    (model.py:449-499).  nn.Linear layers: kaiming-normal(relu) weights (model.py:501-506);
    biases, LayerNorm/BatchNorm affine terms and running statistics are drawn non-trivially
    (``randomize_norms``) so that every term of the forward pass is exercised -- a freshly
    constructed reference model has zero biases and identity norms.
    """
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    D, q, c = hidden, max(hidden // 4, 8), max(hidden // 2, 8)

    def lin(o, i):
        return torch.randn(o, i, generator=g) * math.sqrt(2.0 / i)

    def vec(n, scale=0.1, shift=0.0):
        if not randomize_norms:
            return torch.full((n,), float(shift))
        return torch.randn(n, generator=g) * scale + shift

    s = {}
    s["jk_logits"] = vec(n_layers + 2, 0.5)
    s["in_norm.norm.weight"] = vec(in_channels, 0.1, 1.0)
    s["in_norm.norm.bias"] = vec(in_channels)
    s["in_norm.norm.running_mean"] = vec(in_channels, 0.2, 0.3)
    s["in_norm.norm.running_var"] = (vec(in_channels, 0.1, 0.5).abs() + 0.05)
    s["in_norm.norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    s["input_proj.0.weight"], s["input_proj.0.bias"] = lin(D, in_channels), vec(D)
    s["input_proj.1.weight"], s["input_proj.1.bias"] = vec(D, 0.1, 1.0), vec(D)
    s["prior_booster.0.weight"], s["prior_booster.0.bias"] = lin(q, N_PRIOR_FEATS), vec(q)
    s["prior_booster.2.weight"], s["prior_booster.2.bias"] = lin(D, q), vec(D)
    s["edge_ctx.encode.0.weight"], s["edge_ctx.encode.0.bias"] = lin(c, edge_channels), vec(c)
    s["edge_ctx.encode.2.weight"], s["edge_ctx.encode.2.bias"] = lin(c, c), vec(c)
    s["edge_ctx.to_gate.0.weight"], s["edge_ctx.to_gate.0.bias"] = vec(c, 0.1, 1.0), vec(c)
    s["edge_ctx.to_gate.1.weight"], s["edge_ctx.to_gate.1.bias"] = lin(D, c), vec(D)
    a = math.sqrt(6.0 / (D + D))
    for i in range(n_layers):
        s[f"gcn_layers.{i}.bias"] = vec(D)
        s[f"gcn_layers.{i}.lin.weight"] = (torch.rand(D, D, generator=g) * 2 - 1) * a
        s[f"norms.{i}.weight"], s[f"norms.{i}.bias"] = vec(D, 0.1, 1.0), vec(D)
    b = 1.0 / math.sqrt(D)
    s["sage.lin_l.weight"] = (torch.rand(D, D, generator=g) * 2 - 1) * b
    s["sage.lin_l.bias"] = (torch.rand(D, generator=g) * 2 - 1) * b
    s["sage.lin_r.weight"] = (torch.rand(D, D, generator=g) * 2 - 1) * b
    s["sage_norm.weight"], s["sage_norm.bias"] = vec(D, 0.1, 1.0), vec(D)
    s["ctx.attn.weight"], s["ctx.attn.bias"] = lin(1, D), vec(1)
    s["ctx.compress.weight"], s["ctx.compress.bias"] = lin(D // 2, D), vec(D // 2)
    s["ctx.expand.weight"], s["ctx.expand.bias"] = lin(D, D // 2), vec(D)
    s["fuse.0.weight"], s["fuse.0.bias"] = vec(D, 0.1, 1.0), vec(D)
    s["fuse.1.weight"], s["fuse.1.bias"] = lin(D, D), vec(D)
    s["head.weight"], s["head.bias"] = lin(n_classes, D), vec(n_classes)
    return {k: v.contiguous() for k, v in s.items()}


