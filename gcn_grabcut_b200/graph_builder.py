"""
Region-graph construction on the B200 -- host-side mirror of the reference's
``src/gcn_grabcut/graph_builder.py`` public interface (GraphBuilder, SuperpixelGraph,
SuperpixelGraphConfig, compute_auto_prior, N_* constants; reference lines 64-175, 357-362).

All arithmetic runs in libgcn_grabcut_b200.so (``gg_build_graphs``): one pass over the
pixels for the per-region sums and label transitions, sort/dedup of the adjacency, kNN in
mean-Lab space, node / edge attribute epilogues and the automatic prior.  The array layouts
(19-d node input, 5-d edge attributes, int64 COO in the reference's edge order) are those of
the reference, so the result feeds ``ResGCNNet`` / ``refine_trimap`` / ``cv2.grabCut``
unchanged.

SLIC (``GraphBuilder._compute_superpixels``, reference :177-188) is the input producer of the
path.  A label map passed as ``segments=`` is used as it is; otherwise the superpixels are
computed on the GPU (``slic_labels`` -> ``gg_slic``: scikit-image's algorithm restated, see
csrc/slic.cu -- label-for-label parity with a scikit-image build is unpinned, none can be run
here).  ``SLIC_BACKEND = "skimage"`` (env ``GG_SLIC=skimage``) calls
``skimage.segmentation.slic`` exactly as the reference does when it is importable.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _native as nat

SLIC_BACKEND = os.environ.get("GG_SLIC", "gpu")      # "gpu" | "skimage"

N_IMAGE_FEATS = 16
N_PRIOR_FEATS = 3
N_HINT_FEATS = N_PRIOR_FEATS
N_NODE_FEATS = N_IMAGE_FEATS + N_PRIOR_FEATS
N_EDGE_FEATS = 5


@dataclass
class SuperpixelGraphConfig:
    n_segments: int = 300
    compactness: float = 10.0
    sigma: float = 1.0
    use_lab: bool = True
    connectivity: int = 4
    n_nonlocal: int = 4


@dataclass
class SuperpixelGraph:
    """Container for a built superpixel graph (same fields as the reference's)."""
    segments: np.ndarray
    node_features: np.ndarray
    edge_index: np.ndarray
    edge_attr: np.ndarray
    n_nodes: int = 0
    n_edges: int = 0
    node_centroids: np.ndarray = field(default_factory=lambda: np.empty((0, 2)))
    prior_features: np.ndarray = field(default_factory=lambda: np.empty((0, N_PRIOR_FEATS)))
    node_areas: np.ndarray = field(default_factory=lambda: np.empty((0,)))

    def node_input(self, prior_features: Optional[np.ndarray] = None) -> np.ndarray:
        prior = self.prior_features if prior_features is None else prior_features
        if prior is None or prior.size == 0:
            prior = np.zeros((self.n_nodes, N_PRIOR_FEATS), dtype=np.float32)
        return np.concatenate([self.node_features, prior], axis=1).astype(np.float32)

    def to_pyg(self, prior_features: Optional[np.ndarray] = None):
        import torch
        from .model import Data
        area = self.node_areas
        if area is None or area.size == 0:
            area = np.full(self.n_nodes, 1.0 / max(self.n_nodes, 1), dtype=np.float32)
        return Data(x=torch.tensor(self.node_input(prior_features), dtype=torch.float32),
                    edge_index=torch.tensor(self.edge_index, dtype=torch.long),
                    edge_attr=torch.tensor(self.edge_attr, dtype=torch.float32),
                    node_area=torch.tensor(area, dtype=torch.float32))


# --------------------------------------------------------------------------- batched API
@dataclass
class BatchedRegionGraphs:
    """
    Device-resident ragged batch produced by ``build_graph_batch`` (torch CUDA tensors).
    Image b owns nodes [node_off[b], node_off[b+1]) and edges [edge_off[b], edge_off[b+1]).
    """
    B: int
    node_cap: int
    pair_cap: int
    n_nodes: "torch.Tensor"      # int32 [B]
    n_edges: "torch.Tensor"      # int32 [B]
    node_off: "torch.Tensor"     # int64 [B+1]
    edge_off: "torch.Tensor"     # int64 [B+1]
    x: "torch.Tensor"            # float32 [B*node_cap, 19]
    edge_index: "torch.Tensor"   # int64 [2, 2*B*pair_cap]   (image-local ids)
    edge_attr: "torch.Tensor"    # float32 [2*B*pair_cap, 5]
    centroids: "torch.Tensor"    # float32 [B*node_cap, 2]
    areas: "torch.Tensor"        # float32 [B*node_cap]
    csr_rowptr: "torch.Tensor"   # int32 [B*node_cap+1]
    csr_src: "torch.Tensor"      # int32 [2*B*pair_cap]
    csr_eid: "torch.Tensor"      # int32 [2*B*pair_cap]
    n_adj_pairs: "torch.Tensor"  # int32 [B]
    n_nl_pairs: "torch.Tensor"   # int32 [B]
    shared_cnt: "torch.Tensor"   # int32 [B*pair_cap]

    def to_graphs(self, segments: np.ndarray) -> List[SuperpixelGraph]:
        """Copy to the host and slice into per-image ``SuperpixelGraph`` objects."""
        no = self.node_off.cpu().numpy()
        eo = self.edge_off.cpu().numpy()
        nt, et = int(no[-1]), int(eo[-1])
        x = self.x[:nt].cpu().numpy()
        ei = self.edge_index[:, :et].cpu().numpy()
        ea = self.edge_attr[:et].cpu().numpy()
        cen = self.centroids[:nt].cpu().numpy()
        ar = self.areas[:nt].cpu().numpy()
        out = []
        for b in range(self.B):
            n0, n1, e0, e1 = int(no[b]), int(no[b + 1]), int(eo[b]), int(eo[b + 1])
            out.append(SuperpixelGraph(
                segments=segments[b],
                node_features=np.ascontiguousarray(x[n0:n1, :N_IMAGE_FEATS]),
                edge_index=np.ascontiguousarray(ei[:, e0:e1]),
                edge_attr=np.ascontiguousarray(ea[e0:e1]),
                n_nodes=n1 - n0, n_edges=e1 - e0,
                node_centroids=np.ascontiguousarray(cen[n0:n1]),
                prior_features=np.ascontiguousarray(x[n0:n1, N_IMAGE_FEATS:]),
                node_areas=np.ascontiguousarray(ar[n0:n1])))
        return out


def build_graph_batch(images, segments, config: Optional[SuperpixelGraphConfig] = None,
                      node_cap: Optional[int] = None, pair_cap: Optional[int] = None,
                      device=None, check: bool = True) -> BatchedRegionGraphs:
    """
    Build the region graphs of a batch of images on the GPU.

    images   : uint8 (B,H,W,3) BGR -- numpy array or CUDA tensor
    segments : int32 (B,H,W) label maps -- numpy array or CUDA tensor
    node_cap : capacity per image (> max label); default: max label + 1 (needs a host scan
               for numpy input, a device reduction + sync for tensors)
    check    : synchronise and raise if a label exceeded node_cap or a table overflowed
    """
    import torch
    cfg = config or SuperpixelGraphConfig()
    dev = nat.device_index(device if device is not None else "cuda")
    h = nat.handle(dev)
    tdev = torch.device("cuda", dev)
    img_t = images if torch.is_tensor(images) else torch.from_numpy(np.ascontiguousarray(images))
    seg_t = segments if torch.is_tensor(segments) else torch.from_numpy(np.ascontiguousarray(segments, dtype=np.int32))
    if img_t.dim() != 4 or img_t.shape[-1] != 3 or img_t.dtype != torch.uint8:
        raise ValueError(f"images must be uint8 (B,H,W,3), got {tuple(img_t.shape)} {img_t.dtype}")
    if seg_t.dtype != torch.int32 or tuple(seg_t.shape) != tuple(img_t.shape[:3]):
        raise ValueError(f"segments must be int32 (B,H,W) matching images, got {tuple(seg_t.shape)} {seg_t.dtype}")
    B, H, W = int(img_t.shape[0]), int(img_t.shape[1]), int(img_t.shape[2])
    if node_cap is None:
        node_cap = int(seg_t.max().item()) + 1
    node_cap = max(int(node_cap), 1)
    pair_cap = int(pair_cap) if pair_cap else max(8, node_cap * max(6, 3 + cfg.n_nonlocal))
    img_t = img_t.to(tdev, non_blocking=True).contiguous()
    seg_t = seg_t.to(tdev, non_blocking=True).contiguous()

    SN, SE = B * node_cap, 2 * B * pair_cap
    i32, i64, f32 = torch.int32, torch.int64, torch.float32
    mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=tdev)
    out = BatchedRegionGraphs(
        B=B, node_cap=node_cap, pair_cap=pair_cap,
        n_nodes=mk((B,), i32), n_edges=mk((B,), i32), node_off=mk((B + 1,), i64), edge_off=mk((B + 1,), i64),
        x=mk((SN, N_NODE_FEATS), f32), edge_index=mk((2, SE), i64), edge_attr=mk((SE, N_EDGE_FEATS), f32),
        centroids=mk((SN, 2), f32), areas=mk((SN,), f32), csr_rowptr=mk((SN + 1,), i32),
        csr_src=mk((SE,), i32), csr_eid=mk((SE,), i32), n_adj_pairs=mk((B,), i32),
        n_nl_pairs=mk((B,), i32), shared_cnt=mk((B * pair_cap,), i32))
    gcfg = nat.GraphConfig(int(cfg.connectivity), int(cfg.n_nonlocal), node_cap, pair_cap)
    gout = nat.GraphOut(*[nat.ptr(getattr(out, n)) for n, _ in nat.GraphOut._fields_])
    stream = nat.current_stream(dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_build_graphs(h.ptr, nat.ptr(img_t), nat.ptr(seg_t), B, H, W,
                                            C.byref(gcfg), C.byref(gout), C.c_void_p(stream)))
        if check:
            h.check_status(stream)
    out._keepalive = (img_t, seg_t)
    return out


def slic_labels(images, config: Optional[SuperpixelGraphConfig] = None, max_num_iter: int = 10, device=None,
                return_counts: bool = False):
    """
    SLIC superpixels of a batch of BGR uint8 images (B,H,W,3) on the GPU (``gg_slic``): the
    reference's ``slic(lab, n_segments, compactness, sigma, start_label=0)`` call
    (graph_builder.py:177-188).  Returns int32 (B,H,W) label maps, contiguous 0..N-1 per image
    (and the label counts with ``return_counts``).  Accepts numpy arrays or CUDA tensors (then
    CUDA tensors are returned).
    """
    import torch
    cfg = config or SuperpixelGraphConfig()
    as_tensor = torch.is_tensor(images)
    dev = nat.device_index(images.device if as_tensor and images.is_cuda else (device if device is not None else "cuda"))
    h = nat.handle(dev)
    tdev = torch.device("cuda", dev)
    img_t = images if as_tensor else torch.from_numpy(np.ascontiguousarray(images, dtype=np.uint8))
    img_t = img_t.to(tdev).contiguous()
    if img_t.dtype != torch.uint8 or img_t.dim() != 4 or img_t.shape[-1] != 3:
        raise ValueError("images must be uint8 (B,H,W,3)")
    B, H, W = (int(v) for v in img_t.shape[:3])
    lab = torch.empty((B, H, W), dtype=torch.int32, device=tdev)
    cnt = torch.empty(B, dtype=torch.int32, device=tdev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_slic(h.ptr, nat.ptr(img_t), B, H, W, int(cfg.n_segments), float(cfg.compactness),
                                    float(cfg.sigma), int(max_num_iter), nat.ptr(lab), nat.ptr(cnt),
                                    C.c_void_p(nat.current_stream(dev))))
    if as_tensor:
        return (lab, cnt) if return_counts else lab
    return (lab.cpu().numpy(), cnt.cpu().numpy()) if return_counts else lab.cpu().numpy()


# --------------------------------------------------------------------------- reference API
class GraphBuilder:
    """
    Builds the attributed superpixel graph of one BGR image (reference: graph_builder.py:131-175).

        graph = GraphBuilder(image, cfg, segments=labels).build()
    """

    def __init__(self, image: np.ndarray, config: Optional[SuperpixelGraphConfig] = None,
                 segments: Optional[np.ndarray] = None, device=None):
        import cv2
        self.bgr = image
        self.rgb = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
        self.config = config or SuperpixelGraphConfig()
        self._segments = segments
        self._device = device

    def _compute_superpixels(self) -> np.ndarray:
        """Input producer (reference :177-188): the supplied label map, else SLIC -- on the GPU, or
        scikit-image's when ``SLIC_BACKEND == "skimage"``."""
        if self._segments is not None:
            return np.ascontiguousarray(self._segments, dtype=np.int32)
        cfg = self.config
        if SLIC_BACKEND == "skimage":
            from skimage.segmentation import slic
            from skimage.color import rgb2lab
            img = rgb2lab(self.rgb).astype(np.float32) if cfg.use_lab else self.rgb.astype(float)
            return slic(img, n_segments=cfg.n_segments, compactness=cfg.compactness, sigma=cfg.sigma,
                        start_label=0, channel_axis=-1).astype(np.int32)
        if not cfg.use_lab:
            raise NotImplementedError("the GPU SLIC clusters in CIELAB (use_lab=True, the reference's default)")
        return slic_labels(self.bgr[None], cfg, device=self._device)[0]

    def build(self, segments: Optional[np.ndarray] = None) -> SuperpixelGraph:
        if segments is not None:
            self._segments = segments
        seg = self._compute_superpixels()
        if seg.shape != self.bgr.shape[:2]:
            raise ValueError(f"segments shape {seg.shape} != image shape {self.bgr.shape[:2]}")
        batch = build_graph_batch(self.bgr[None], seg[None], self.config, device=self._device)
        return batch.to_graphs(seg[None])[0]


def compute_auto_prior(segments: np.ndarray, lab: Optional[np.ndarray] = None,
                       centre_sigma: float = 0.45, contrast_sigma: float = 0.40, *,
                       image: Optional[np.ndarray] = None, device=None) -> np.ndarray:
    """
    Per-superpixel [fg-ness, bg-ness, ambiguity] prior (reference: graph_builder.py:357-444), on
    the GPU.  Same call as the reference: ``compute_auto_prior(segments, lab)`` with the float32
    CIELAB image (H,W,3) -- its region sums are taken on the device (``gg_auto_prior``).
    Alternatively pass ``image=`` (BGR uint8) and no ``lab``: Lab is then derived from the image
    by the graph builder's own pixel kernel (default sigmas only).
    """
    import torch
    seg = np.ascontiguousarray(segments, dtype=np.int32)
    if lab is None:
        if image is None:
            raise ValueError("compute_auto_prior needs the Lab plane (or image=<BGR uint8>)")
        if abs(centre_sigma - 0.45) > 1e-12 or abs(contrast_sigma - 0.40) > 1e-12:
            raise NotImplementedError("with image= only centre_sigma=0.45, contrast_sigma=0.40 are supported")
        cfg = SuperpixelGraphConfig(n_nonlocal=0)
        g = build_graph_batch(image[None], seg[None], cfg, device=device).to_graphs(seg[None])[0]
        return g.prior_features
    lab = np.ascontiguousarray(lab, dtype=np.float32)
    if lab.shape != seg.shape + (3,):
        raise ValueError(f"lab shape {lab.shape} != segments shape {seg.shape} + (3,)")
    dev = nat.device_index(device if device is not None else "cuda")
    h = nat.handle(dev)
    tdev = torch.device("cuda", dev)
    n = int(seg.max()) + 1
    H, W = seg.shape
    seg_t, lab_t = torch.from_numpy(seg).to(tdev), torch.from_numpy(lab).to(tdev)
    prior = torch.empty(n, 3, dtype=torch.float32, device=tdev)
    stream = nat.current_stream(dev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_auto_prior(h.ptr, nat.ptr(seg_t), nat.ptr(lab_t), 1, H, W, n, float(centre_sigma),
                                          float(contrast_sigma), nat.ptr(prior), C.c_void_p(0), C.c_void_p(stream)))
        h.check_status(stream)
    return prior.cpu().numpy()


def encode_user_hints(segments: np.ndarray, fg_points, bg_points) -> np.ndarray:
    """Legacy click channels (reference :457-494); host-side, not on the trimap path."""
    n = int(segments.max()) + 1
    hints = np.zeros((n, 3), dtype=np.float32)
    hints[:, 2] = 1.0
    for col, pts in ((0, fg_points), (1, bg_points)):
        for r, c in pts:
            r, c = int(r), int(c)
            if 0 <= r < segments.shape[0] and 0 <= c < segments.shape[1]:
                nid = int(segments[r, c])
                hints[nid, col] = 1.0
                hints[nid, 2] = 0.0
    return hints
