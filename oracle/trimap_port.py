"""
CPU restatement of the region->pixel projection: guided filter + thresholds -> trimap.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
/root/reference/src/gcn_grabcut/pipeline.py:71-146.  ``cv2.blur`` is the reference's own
backend and is called directly; ``box_mean_f64`` restates its float32 numerics (float64
window sums over BORDER_REFLECT_101, times 1/k^2, cast to float32) so that the CUDA
kernel's arithmetic can be checked without OpenCV in the loop.
"""

from __future__ import annotations

import cv2
import numpy as np

from .model_port import CLASS_BG, CLASS_FG, project_to_pixels

F32 = np.float32


def guided_filter(guide: np.ndarray, src: np.ndarray, radius: int = 8, eps: float = 1e-3) -> np.ndarray:
    """pipeline.py:71-100 -- He et al. box-filter formulation, all planes float32."""
    k = (2 * radius + 1, 2 * radius + 1)
    mean_g = cv2.blur(guide, k)
    mean_s = cv2.blur(src, k)
    cov = cv2.blur(guide * src, k) - mean_g * mean_s
    var = cv2.blur(guide * guide, k) - mean_g * mean_g
    a = cov / (var + F32(eps))
    b = mean_s - a * mean_g
    return cv2.blur(a, k) * guide + cv2.blur(b, k)


def refine_trimap(probs: np.ndarray, seg: np.ndarray, bgr: np.ndarray, thr_fg: float = 0.55,
                  thr_bg: float = 0.55, radius: int = 8, eps: float = 1e-3,
                  return_planes: bool = False):
    """pipeline.py:103-146 -- (H,W) uint8 trimap in cv2.GC_* label space."""
    guide = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY).astype(F32) / F32(255.0)
    p_bg = project_to_pixels(probs[:, CLASS_BG].astype(F32), seg)
    p_fg = project_to_pixels(probs[:, CLASS_FG].astype(F32), seg)
    p_bg = np.clip(guided_filter(guide, p_bg, radius, eps), 0.0, 1.0)
    p_fg = np.clip(guided_filter(guide, p_fg, radius, eps), 0.0, 1.0)
    tri = np.where(p_fg > p_bg, 3, 2).astype(np.uint8)
    tri[p_bg >= thr_bg] = 0
    tri[p_fg >= thr_fg] = 1
    if return_planes:
        return tri, p_bg, p_fg
    return tri


def box_mean_f64(img: np.ndarray, radius: int) -> np.ndarray:
    """cv2.blur(float32, (2r+1,2r+1)) restated: float64 window sums, REFLECT_101 border."""
    k = 2 * radius + 1
    pad = cv2.copyMakeBorder(img, radius, radius, radius, radius, cv2.BORDER_REFLECT_101)
    ii = np.zeros((pad.shape[0] + 1, pad.shape[1] + 1), dtype=np.float64)
    ii[1:, 1:] = pad.astype(np.float64).cumsum(0).cumsum(1)
    H, W = img.shape
    s = ii[k:k + H, k:k + W] - ii[0:H, k:k + W] - ii[k:k + H, 0:W] + ii[0:H, 0:W]
    return (s * (1.0 / (k * k))).astype(F32)


def near_threshold_mask(p_bg: np.ndarray, p_fg: np.ndarray, thr_fg: float, thr_bg: float,
                        tol: float) -> np.ndarray:
    """Pixels whose label decision lies within ``tol`` of a decision boundary (eq. 27)."""
    return ((np.abs(p_fg - F32(thr_fg)) <= tol) | (np.abs(p_bg - F32(thr_bg)) <= tol) |
            (np.abs(p_fg - p_bg) <= tol))


def seed_from_prior(trimap: np.ndarray, prior: np.ndarray, seg: np.ndarray, n_nodes: int,
                    seed_frac: float = 0.1) -> np.ndarray:
    """pipeline.py:149-186 (_seed_from_prior) -- promote the most confident regions of the
    automatic prior when the trimap lacks a foreground or a background label.  Equal prior values
    are taken larger index first (stable ascending argsort, reversed); the reference's default
    argsort leaves the order of ties to numpy's introsort."""
    if prior is None or prior.size == 0:
        return trimap
    has_fg = np.isin(trimap, (1, 3)).any()
    has_bg = np.isin(trimap, (0, 2)).any()
    if has_fg and has_bg:
        return trimap
    n_seed = max(1, int(round(seed_frac * n_nodes)))
    trimap = trimap.copy()
    if not has_fg:
        ids = np.argsort(prior[:, 0], kind="stable")[::-1][:n_seed]
        trimap[np.isin(seg, ids)] = 3
    if not has_bg:
        ids = np.argsort(prior[:, 1], kind="stable")[::-1][:n_seed]
        trimap[np.isin(seg, ids)] = 2
    return trimap


def grabcut_guards(trimap: np.ndarray):
    """grabcut.py:127-140 -- promotion of probable to definite labels when a definite side is
    missing, and the single-class short-circuit.  Returns (trimap, degenerate)."""
    trimap = trimap.astype(np.uint8)
    if not (trimap == cv2.GC_FGD).any():
        trimap = trimap.copy()
        trimap[trimap == cv2.GC_PR_FGD] = cv2.GC_FGD
    if not (trimap == cv2.GC_BGD).any():
        trimap = trimap.copy()
        trimap[trimap == cv2.GC_PR_BGD] = cv2.GC_BGD
    degenerate = (not (trimap == cv2.GC_FGD).any()) or (not (trimap == cv2.GC_BGD).any())
    return trimap, bool(degenerate)


def clean_mask(mask: np.ndarray, min_area_ratio: float = 0.002, keep_largest: bool = False) -> np.ndarray:
    """pipeline.py:189-227 -- connected-component clean-up of a {0,1} mask (8-connectivity)."""
    if mask.sum() == 0 or (min_area_ratio <= 0 and not keep_largest):
        return mask
    n_labels, labels, stats, _ = cv2.connectedComponentsWithStats(mask.astype(np.uint8), connectivity=8)
    if n_labels <= 1:
        return mask
    areas = stats[1:, cv2.CC_STAT_AREA]
    min_area = min_area_ratio * mask.size
    if keep_largest:
        keep = np.array([int(areas.argmax()) + 1])
    else:
        keep = np.flatnonzero(areas >= min_area) + 1
        if keep.size == 0:
            keep = np.array([int(areas.argmax()) + 1])
    return np.isin(labels, keep).astype(np.uint8)
