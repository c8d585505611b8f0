"""
Training-data path of the graph builder -- host-side mirror of the reference's
``src/gcn_grabcut/dataset.py`` label derivation (``derive_trimap_labels`` :175-205,
``prepare_sample`` :212-260), SURVEY 8(f) rank 3: the second caller of the builder.

The per-region pixel / foreground counts are exact integers, so ``fg_ratio`` and the labels
are bit-identical to the reference's.  ``prepare_samples`` is the batched form ("ten thousand
graphs built in minutes", reference README): one ``gg_build_graphs`` + one
``gg_region_labels`` call for B images.  ``prepare_dataset`` adds the reference's on-disk graph
cache (``dataset.py:363-441``): the same sha1 key over the sample and the configuration, the same
blob ``{"data": Data(x, edge_index, edge_attr, node_area, fg_ratio, y), "segments"}`` written
atomically as ``<key>.pt`` -- with the cache misses of a call built in batches on the GPU instead
of one image per worker process.  Dataset discovery and augmentation (``dataset.py:60-170``) are
file IO and stay with the reference.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
from pathlib import Path
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


# ----------------------------------------------------------------------------- graph cache
def cache_key(sample: dict, sp_config: Optional[SuperpixelGraphConfig], fg_threshold: float,
              bg_threshold: float) -> str:
    """The reference's cache key (dataset.py:363-377), byte for byte: sha1 over the image and mask
    bytes (or, for a lazily decoded sample, the repr of its paths / max_size / aug_seed) followed
    by the repr of the configuration tuple; first 20 hex digits.  Entries written by either
    implementation are found by the other."""
    cfg = sp_config or SuperpixelGraphConfig()
    digest = hashlib.sha1()
    if "image" in sample:
        for key in ("image", "gt_mask"):
            digest.update(np.ascontiguousarray(sample[key]))
    else:
        digest.update(repr((sample["image_path"], sample["mask_path"], sample.get("max_size"),
                            sample.get("aug_seed"))).encode())
    digest.update(repr((cfg.n_segments, cfg.compactness, cfg.sigma, cfg.use_lab, cfg.connectivity,
                        cfg.n_nonlocal, fg_threshold, bg_threshold)).encode())
    return digest.hexdigest()[:20]


_cache_key = cache_key      # the reference's (private) name


def _write_blob(path: Path, data, segments) -> None:
    """torch.save to a temporary name, then rename: an interrupted run cannot leave a truncated
    entry behind (dataset.py:431-440)."""
    import torch
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_suffix(f".{os.getpid()}.tmp")
    try:
        torch.save({"data": data, "segments": segments}, tmp)
        os.replace(tmp, path)
    except Exception:
        tmp.unlink(missing_ok=True)


def prepare_dataset(samples: List[dict], sp_config: Optional[SuperpixelGraphConfig] = None,
                    fg_threshold: float = 0.70, bg_threshold: float = 0.70, cache_dir=None,
                    keep_segments: bool = True, segments: Optional[List[np.ndarray]] = None,
                    batch_size: int = 64, device=None) -> List[Optional[Tuple]]:
    """
    Cached, batched graph preparation -- the role of the reference's ``prepare_dataset`` /
    ``_prepare_one`` (dataset.py:402-441) for in-memory samples (dicts with ``image`` and
    ``gt_mask``).  Per sample: look the key up in ``cache_dir`` (a hit never touches the GPU; a
    corrupt entry is rebuilt); the misses are grouped by image shape and built ``batch_size`` at a
    time with ``prepare_samples``; every new graph is written as ``<key>.pt``.  Returns one
    ``(data, labels, segments or None)`` per sample, in order.  ``segments`` optionally supplies the
    label maps (SLIC is the input producer).
    """
    import torch
    cfg = sp_config or SuperpixelGraphConfig()
    out: List[Optional[Tuple]] = [None] * len(samples)
    paths: List[Optional[Path]] = [None] * len(samples)
    misses = []
    for i, smp in enumerate(samples):
        if cache_dir is not None:
            paths[i] = Path(cache_dir) / f"{cache_key(smp, cfg, fg_threshold, bg_threshold)}.pt"
            if paths[i].exists():
                try:
                    blob = torch.load(paths[i], map_location="cpu", weights_only=False)
                    data = blob["data"]
                    out[i] = (data, data.y, blob["segments"] if keep_segments else None)
                    continue
                except Exception:
                    pass                      # corrupt or stale entry: rebuild it
        misses.append(i)
    by_shape = {}
    for i in misses:
        by_shape.setdefault(tuple(samples[i]["image"].shape[:2]), []).append(i)
    for idxs in by_shape.values():
        for k in range(0, len(idxs), batch_size):
            chunk = idxs[k:k + batch_size]
            imgs = np.stack([samples[i]["image"] for i in chunk])
            msks = np.stack([np.asarray(samples[i]["gt_mask"]) for i in chunk])
            segs = None if segments is None else np.stack([segments[i] for i in chunk])
            for i, (data, y, seg) in zip(chunk, prepare_samples(imgs, msks, segs, cfg, fg_threshold, bg_threshold,
                                                                 device=device)):
                if paths[i] is not None:
                    _write_blob(paths[i], data, seg)
                out[i] = (data, y, seg if keep_segments else None)
    return out
