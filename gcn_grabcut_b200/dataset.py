"""
Training-data path of the graph builder -- host-side mirror of the reference's
``src/gcn_grabcut/dataset.py`` label derivation (``derive_trimap_labels`` :175-205,
``prepare_sample`` :212-260), SURVEY 8(f) rank 3: the second caller of the builder.

The per-region pixel / foreground counts are exact integers, so ``fg_ratio`` and the labels
are bit-identical to the reference's.  ``prepare_samples`` is the batched form ("ten thousand
graphs built in minutes", reference README): one ``gg_build_graphs`` + one
``gg_region_labels`` call for B images.  Dataset discovery, augmentation and the ``.pt`` cache
(``dataset.py:60-170, 363-441``) are file IO and stay with the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import _native as nat
from .graph_builder import GraphBuilder, SuperpixelGraphConfig, build_graph_batch
from .model import CLASS_BG, CLASS_FG, CLASS_UNK, Data  # noqa: F401  (same names as the reference imports)


def _region_labels_device(seg_t, mask_t, node_off_t, n_rows: int, fg_threshold: float,
                          bg_threshold: float, dev: int):
    import torch
    h = nat.handle(dev)
    B, H, W = (int(v) for v in seg_t.shape)
    ratio = torch.empty(n_rows, dtype=torch.float32, device=seg_t.device)
    y = torch.empty(n_rows, dtype=torch.int64, device=seg_t.device)
    stream = nat.current_stream(dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_region_labels(h.ptr, nat.ptr(seg_t), nat.ptr(mask_t), nat.ptr(node_off_t), B, H, W,
                                             n_rows, float(fg_threshold), float(bg_threshold), nat.ptr(ratio),
                                             nat.ptr(y), C.c_void_p(stream)))
        h.check_status(stream)
    return ratio, y


def derive_trimap_labels(segments: np.ndarray, gt_mask: np.ndarray, fg_threshold: float = 0.75,
                         bg_threshold: float = 0.75, device=None) -> np.ndarray:
    """3-class label per superpixel by foreground coverage (reference dataset.py:175-205).
    Returns (N,) int64 with CLASS_BG = 0, CLASS_UNK = 1, CLASS_FG = 2."""
    import torch
    dev = nat.device_index(device if device is not None else "cuda")
    tdev = torch.device("cuda", dev)
    if gt_mask.shape != segments.shape:
        raise ValueError(f"gt_mask shape {gt_mask.shape} != segments shape {segments.shape}")
    n = int(segments.max()) + 1
    seg = torch.from_numpy(np.ascontiguousarray(segments, dtype=np.int32)[None]).to(tdev)
    msk = torch.from_numpy(np.ascontiguousarray(gt_mask > 0).astype(np.uint8)[None]).to(tdev)
    off = torch.tensor([0, n], dtype=torch.int64, device=tdev)
    _, y = _region_labels_device(seg, msk, off, n, fg_threshold, bg_threshold, dev)
    return y.cpu().numpy()


def prepare_sample(sample: dict, sp_config: Optional[SuperpixelGraphConfig] = None,
                   fg_threshold: float = 0.70, bg_threshold: float = 0.70, segments=None) -> tuple:
    """Raw sample dict (keys ``image``, ``gt_mask``) -> (Data, labels tensor, segments array), the
    reference's ``prepare_sample`` (dataset.py:212-260).  ``segments`` supplies the label map when
    scikit-image's SLIC is not installed."""
    res = prepare_samples(sample["image"][None], np.asarray(sample["gt_mask"])[None],
                          None if segments is None else np.asarray(segments)[None], sp_config, fg_threshold,
                          bg_threshold)
    return res[0]


def prepare_samples(images: np.ndarray, gt_masks: np.ndarray, segments: Optional[np.ndarray] = None,
                    sp_config: Optional[SuperpixelGraphConfig] = None, fg_threshold: float = 0.70,
                    bg_threshold: float = 0.70, device=None) -> List[Tuple]:
    """Batched ``prepare_sample``: images uint8 (B,H,W,3), gt_masks (B,H,W) (> 0 = foreground),
    segments int32 (B,H,W) or None (then SLIC runs per image like the reference).  Returns a list
    of (Data(x, edge_index, edge_attr, node_area, fg_ratio, y), y, segments) per image."""
    import torch
    cfg = sp_config or SuperpixelGraphConfig()
    images = np.ascontiguousarray(images, dtype=np.uint8)
    if segments is None:
        segments = np.stack([GraphBuilder(img, cfg)._compute_superpixels() for img in images])
    segments = np.ascontiguousarray(segments, dtype=np.int32)
    if gt_masks.shape != segments.shape:
        raise ValueError(f"gt_masks shape {gt_masks.shape} != segments shape {segments.shape}")
    dev = nat.device_index(device if device is not None else "cuda")
    batch = build_graph_batch(images, segments, cfg, device=dev)
    seg_t = batch._keepalive[1]
    msk_t = torch.from_numpy(np.ascontiguousarray(gt_masks > 0).astype(np.uint8)).to(seg_t.device)
    ratio, y = _region_labels_device(seg_t, msk_t, batch.node_off, batch.B * batch.node_cap, fg_threshold,
                                     bg_threshold, dev)
    ratio, y = ratio.cpu(), y.cpu()
    node_off = batch.node_off.cpu().numpy()
    out = []
    for b, g in enumerate(batch.to_graphs(segments)):
        lo, hi = int(node_off[b]), int(node_off[b + 1])
        yb = y[lo:hi].clone()
        data = Data(x=torch.tensor(g.node_input(), dtype=torch.float32),
                    edge_index=torch.tensor(g.edge_index, dtype=torch.long),
                    edge_attr=torch.tensor(g.edge_attr, dtype=torch.float32),
                    node_area=torch.tensor(g.node_areas, dtype=torch.float32),
                    fg_ratio=ratio[lo:hi].clone(), y=yb)
        out.append((data, data.y, g.segments))
    return out
