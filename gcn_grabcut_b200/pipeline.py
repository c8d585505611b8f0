"""
Region -> pixel projection and the batched trimap path -- host-side mirror of the trimap part
of the reference's ``src/gcn_grabcut/pipeline.py`` (guided_filter :71-100, refine_trimap
:103-146, and the graph -> network -> trimap section of GCNGrabCutPipeline.segment :296-322).

``guided_filter`` / ``refine_trimap`` keep the reference's signatures and numpy in/out, so the
reference's ``segment()`` can import them unchanged; ``TrimapPath`` is the batched form that
the throughput harness uses: B images + label maps in, B OpenCV trimaps out, everything in
between on the GPU (``gg_trimap_path_host`` / ``gg_trimap_path_device``).  The GrabCut
refinement downstream (``cv2.grabCut``) is not part of the path and stays as it is.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _native as nat
from .graph_builder import SuperpixelGraphConfig
from .model import CLASS_BG, CLASS_FG, project_to_pixels  # noqa: F401  (re-exported like the reference)


def guided_filter(guide: np.ndarray, src: np.ndarray, radius: int = 8, eps: float = 1e-3,
                  device=None) -> np.ndarray:
    """Edge-preserving filter of ``src`` under ``guide`` (float32 (H,W) planes), on the GPU."""
    import torch
    dev = nat.device_index(device if device is not None else "cuda")
    h = nat.handle(dev)
    tdev = torch.device("cuda", dev)
    g = torch.from_numpy(np.ascontiguousarray(guide, dtype=np.float32)).to(tdev)
    s = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32)).to(tdev)
    if g.dim() != 2 or g.shape != s.shape:
        raise ValueError("guide and src must be (H,W) planes of equal shape")
    out = torch.empty_like(g)
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_guided_filter(h.ptr, nat.ptr(g), nat.ptr(s), int(g.shape[0]), int(g.shape[1]),
                                             int(radius), float(eps), nat.ptr(out),
                                             C.c_void_p(nat.current_stream(dev))))
    return out.cpu().numpy()


def refine_trimap(probs: np.ndarray, segments: np.ndarray, image: np.ndarray,
                  threshold_fg: float = 0.55, threshold_bg: float = 0.55, radius: int = 8,
                  eps: float = 1e-3, device=None, return_planes: bool = False):
    """
    Per-region class probabilities -> pixel-level trimap whose boundaries follow image edges
    (reference pipeline.py:103-146).  probs (N,3) [BG,UNK,FG]; segments (H,W); image BGR uint8.
    Returns (H,W) uint8 in cv2.GC_* label space.
    """
    import torch
    dev = nat.device_index(device if device is not None else "cuda")
    h = nat.handle(dev)
    tdev = torch.device("cuda", dev)
    if image.shape[:2] != segments.shape:
        raise ValueError(f"segments shape {segments.shape} != image shape {image.shape[:2]}")
    H, W = segments.shape
    p = torch.from_numpy(np.ascontiguousarray(probs, dtype=np.float32)).to(tdev)
    seg = torch.from_numpy(np.ascontiguousarray(segments, dtype=np.int32)).to(tdev)
    img = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8)).to(tdev)
    goff = torch.tensor([0, p.shape[0]], dtype=torch.int64, device=tdev)
    tri = torch.empty(H, W, dtype=torch.uint8, device=tdev)
    pb = torch.empty(H, W, dtype=torch.float32, device=tdev) if return_planes else None
    pf = torch.empty(H, W, dtype=torch.float32, device=tdev) if return_planes else None
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_refine_trimap(h.ptr, nat.ptr(img), nat.ptr(seg), nat.ptr(p), nat.ptr(goff),
                                             1, H, W, int(radius), float(eps), float(threshold_fg),
                                             float(threshold_bg), nat.ptr(tri), nat.ptr(pb), nat.ptr(pf),
                                             C.c_void_p(nat.current_stream(dev))))
    if return_planes:
        return tri.cpu().numpy(), pb.cpu().numpy(), pf.cpu().numpy()
    return tri.cpu().numpy()


def seed_from_prior(trimap: np.ndarray, graph, seed_frac: float = 0.1, device=None) -> np.ndarray:
    """
    Guarantee the trimap contains both a foreground and a background seed (reference
    ``_seed_from_prior``, pipeline.py:149-186): if one side is missing, the
    ``max(1, round(seed_frac * n_nodes))`` regions with the largest automatic prior for that side
    are promoted to FG_PROBABLE / BG_PROBABLE.  Returns a new (H,W) uint8 array.
    """
    import torch
    prior = graph.prior_features
    if prior is None or prior.size == 0:
        return trimap
    dev = nat.device_index(device if device is not None else "cuda")
    h = nat.handle(dev)
    tdev = torch.device("cuda", dev)
    seg = graph.segments
    if trimap.shape != seg.shape:
        raise ValueError(f"trimap shape {trimap.shape} != segments shape {seg.shape}")
    H, W = seg.shape
    n = int(graph.n_nodes)
    x = torch.from_numpy(np.ascontiguousarray(graph.node_input(), dtype=np.float32)).to(tdev)
    tri = torch.from_numpy(np.ascontiguousarray(trimap, dtype=np.uint8)).to(tdev)
    lab = torch.from_numpy(np.ascontiguousarray(seg, dtype=np.int32)).to(tdev)
    goff = torch.tensor([0, n], dtype=torch.int64, device=tdev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_seed_from_prior(h.ptr, nat.ptr(tri), nat.ptr(lab), nat.ptr(x), nat.ptr(goff),
                                               1, H, W, n, float(seed_frac),
                                               C.c_void_p(nat.current_stream(dev))))
    return tri.cpu().numpy()


_seed_from_prior = seed_from_prior      # the reference's (private) name


def grabcut_guards(trimaps, device=None):
    """
    The guards ``GrabCut.run_with_trimap`` applies before ``cv2.grabCut`` (reference
    grabcut.py:127-140), for one (H,W) trimap or a (B,H,W) batch: without a definite-foreground
    pixel the probable-foreground pixels are promoted to definite, likewise for the background.
    Returns ``(trimaps, degenerate)``: ``degenerate[b]`` is True where a side is still missing --
    the reference then returns the trimap's own labelling instead of calling ``cv2.grabCut``.
    """
    import torch
    t = np.ascontiguousarray(trimaps, dtype=np.uint8)
    single = t.ndim == 2
    if single:
        t = t[None]
    dev = nat.device_index(device if device is not None else "cuda")
    h = nat.handle(dev)
    tdev = torch.device("cuda", dev)
    tt = torch.from_numpy(t).to(tdev)
    deg = torch.zeros(t.shape[0], dtype=torch.int32, device=tdev)
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_grabcut_guards(h.ptr, nat.ptr(tt), int(t.shape[0]), int(t.shape[1]), int(t.shape[2]),
                                              nat.ptr(deg), C.c_void_p(nat.current_stream(dev))))
    out, d = tt.cpu().numpy(), deg.cpu().numpy().astype(bool)
    return (out[0], bool(d[0])) if single else (out, d)


def clean_mask(mask: np.ndarray, min_area_ratio: float = 0.002, keep_largest: bool = False,
               device=None) -> np.ndarray:
    """
    Remove spurious connected components from an automatic mask (reference pipeline.py:189-227):
    mask (H,W) uint8 {0,1}; components (8-connectivity) smaller than ``min_area_ratio`` of the
    image are dropped -- the largest one is kept if none survives -- or only the largest is kept.
    A (B,H,W) batch is cleaned image by image in one call (``gg_clean_masks``).
    """
    import torch
    m = np.asarray(mask)
    if min_area_ratio <= 0 and not keep_largest:               # reference early return
        return mask
    single = m.ndim == 2
    mb = np.ascontiguousarray(m[None] if single else m).astype(np.uint8)
    if single and mb.sum() == 0:
        return mask
    dev = nat.device_index(device if device is not None else "cuda")
    h = nat.handle(dev)
    tdev = torch.device("cuda", dev)
    mt = torch.from_numpy(mb).to(tdev)
    out = torch.empty_like(mt)
    with torch.cuda.device(dev):
        nat.check(nat.lib().gg_clean_masks(h.ptr, nat.ptr(mt), nat.ptr(out), int(mb.shape[0]), int(mb.shape[1]),
                                           int(mb.shape[2]), float(min_area_ratio), int(bool(keep_largest)),
                                           C.c_void_p(nat.current_stream(dev))))
    res = out.cpu().numpy()
    return res[0] if single else res


class TrimapPath:
    """
    The whole per-image trimap path, batched:  BGR images + label maps -> OpenCV trimaps.

        path = TrimapPath(state_dict_or_model, SuperpixelGraphConfig(), node_cap=320)
        trimaps = path(images_uint8_BHW3, labels_int32_BHW)          # host in / host out

    Equivalent, per image, to the reference's
        graph  = GraphBuilder(image, cfg).build()                    (label map supplied)
        probs  = model.predict_probs(Data(graph.node_input(), graph.edge_index, graph.edge_attr))
        trimap = refine_trimap(probs, graph.segments, image, thr_fg, thr_bg, radius)
    (pipeline.py:298-317), or ``model.predict_trimap`` when ``edge_aware=False``; with
    ``seed_frac=0.1`` also the ``_seed_from_prior`` repair that ``segment()`` applies next
    (pipeline.py:324), i.e. everything between ``cv2.imread`` + SLIC and ``cv2.grabCut``.
    """

    def __init__(self, model, sp_config: Optional[SuperpixelGraphConfig] = None, node_cap: int = 512,
                 pair_cap: int = 0, threshold_fg: float = 0.55, threshold_bg: float = 0.55,
                 filter_radius: int = 8, eps: float = 1e-3, edge_aware: bool = True, chunk: int = 0,
                 seed_frac: float = 0.0, device=None, device_slic: bool = False):
        self.cfg = sp_config or SuperpixelGraphConfig()
        self.dev = nat.device_index(device if device is not None else "cuda")
        self.h = nat.handle(self.dev)
        state = model.state_dict() if hasattr(model, "state_dict") else model
        self._state = {k: v.detach().clone() if hasattr(v, "detach") else v for k, v in state.items()}
        nat.load_state_dict(self.h, self._state)
        self.h.weights_token = self
        self.node_cap = int(node_cap)
        pair_cap = int(pair_cap) if pair_cap else self.node_cap * max(6, 3 + self.cfg.n_nonlocal)
        self.pc = nat.PathConfig(
            nat.GraphConfig(int(self.cfg.connectivity), int(self.cfg.n_nonlocal), self.node_cap, pair_cap),
            int(filter_radius), float(eps), float(threshold_fg), float(threshold_bg), int(bool(edge_aware)),
            int(chunk), 4, 0, float(seed_frac),
            int(self.cfg.n_segments) if device_slic else 0, 10, float(self.cfg.compactness), float(self.cfg.sigma))
        self.device_slic = bool(device_slic)

    def _ensure_weights(self):
        # one device handle holds one set of weights: if another model (a ResGCNNet, another
        # TrimapPath) used the handle in between, this path's weights are loaded again (the
        # library waits for queued work before it replaces them)
        if self.h.weights_token is not self:
            nat.load_state_dict(self.h, self._state)
            self.h.weights_token = self

    def __call__(self, images, labels=None, out: Optional[np.ndarray] = None,
                 return_counts: bool = False):
        """Host buffers (numpy or pinned torch CPU tensors) in, host trimaps out; copies included
        (``gg_trimap_path_host``: synchronous, one batch at a time)."""
        return self.submit(images, labels, out, return_counts, _sync=True).result()

    def submit(self, images, labels=None, out=None, return_counts: bool = False, _sync: bool = False) -> "PendingTrimaps":
        """
        Asynchronous ``__call__`` for streaming many batches: enqueues the copies and kernels of
        this batch (``gg_trimap_path_host_submit``) and returns at once; ``.result()`` of the
        returned object blocks until the trimaps are in the host buffer.  With two or three
        batches in flight the copy-in of the next batch overlaps the kernels of the current one.
        The input and output buffers must not be modified until ``.result()`` returns (pinned
        buffers are needed for the copies to overlap).
        """
        import torch
        self._ensure_weights()
        img = images if torch.is_tensor(images) else torch.from_numpy(np.ascontiguousarray(images))
        if labels is None:
            # superpixels on the device (TrimapPath(..., device_slic=True)): only the images cross PCIe
            if not self.device_slic:
                raise ValueError("labels=None needs TrimapPath(..., device_slic=True)")
            lab = None
        elif torch.is_tensor(labels):
            lab = labels
        elif isinstance(labels, np.ndarray) and labels.dtype == np.uint16:
            lab = torch.from_numpy(np.ascontiguousarray(labels))     # compact transport, see below
        else:
            lab = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32))
        if img.is_cuda or (lab is not None and lab.is_cuda):
            raise ValueError("TrimapPath.__call__ takes host buffers; use run_device for CUDA tensors")
        if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3:
            raise ValueError("images must be uint8 (B,H,W,3)")
        # int32 label maps are the reference layout (graph_builder.py:188); uint16 maps (labels
        # < 65536) are accepted as a compact transport: 5 instead of 7 bytes per pixel over PCIe
        if lab is not None and (lab.dtype not in (torch.int32, torch.uint16) or tuple(lab.shape) != tuple(img.shape[:3])):
            raise ValueError("labels must be int32 (or uint16) (B,H,W) matching images")
        B, H, W = int(img.shape[0]), int(img.shape[1]), int(img.shape[2])
        tri = out if out is not None else torch.empty((B, H, W), dtype=torch.uint8)
        tri_t = tri if torch.is_tensor(tri) else torch.from_numpy(tri)
        nn_ = torch.empty(B, dtype=torch.int32) if return_counts else None
        ne_ = torch.empty(B, dtype=torch.int32) if return_counts else None
        img = img.contiguous()
        lab = lab.contiguous() if lab is not None else None
        pc = self.pc
        if lab is not None and lab.dtype == torch.uint16:
            pc = nat.PathConfig.from_buffer_copy(self.pc)
            pc.label_bytes = 2
        ticket = C.c_int(-1)
        with torch.cuda.device(self.dev):
            if _sync:
                nat.check(nat.lib().gg_trimap_path_host(self.h.ptr, nat.ptr(img), nat.ptr(lab), B, H, W,
                                                        C.byref(pc), nat.ptr(tri_t), nat.ptr(nn_), nat.ptr(ne_)))
                return PendingTrimaps(self, -1, None, tri_t, torch.is_tensor(out), nn_, ne_)
            nat.check(nat.lib().gg_trimap_path_host_submit(self.h.ptr, nat.ptr(img), nat.ptr(lab), B, H, W,
                                                           C.byref(pc), nat.ptr(tri_t), nat.ptr(nn_),
                                                           nat.ptr(ne_), C.byref(ticket)))
        return PendingTrimaps(self, ticket.value, (img, lab), tri_t, torch.is_tensor(out), nn_, ne_)

    def check_status(self) -> None:
        """Read (and clear) the device status word of the ``run_device`` calls enqueued so far on
        the current stream; raises ``NativeError`` (GG_ERR_CAPACITY) if any of them saw a label
        >= ``node_cap`` or overflowed the pair / edge capacity -- such a batch's trimaps are NOT
        valid (out-of-range labels are clamped on the device).  Synchronises the stream."""
        import torch
        with torch.cuda.device(self.dev):
            self.h.check_status(nat.current_stream(self.dev))

    def run_device(self, images_t, labels_t=None, trimap_t=None, probs_t=None, node_off_t=None, check: bool = False):
        """CUDA tensors in, CUDA trimaps out, on the current stream, no host synchronisation
        (``check=False``).  The device status is sticky: call ``check_status()`` once after a run of
        calls, or pass ``check=True`` to synchronise and verify this call."""
        import torch
        self._ensure_weights()
        B, H, W = int(images_t.shape[0]), int(images_t.shape[1]), int(images_t.shape[2])
        if trimap_t is None:
            trimap_t = torch.empty((B, H, W), dtype=torch.uint8, device=images_t.device)
        with torch.cuda.device(self.dev):
            nat.check(nat.lib().gg_trimap_path_device(
                self.h.ptr, nat.ptr(images_t, torch.uint8),
                nat.ptr(labels_t, torch.int32) if labels_t is not None else C.c_void_p(0), B, H, W,
                C.byref(self.pc), nat.ptr(trimap_t), nat.ptr(probs_t), nat.ptr(node_off_t),
                C.c_void_p(nat.current_stream(self.dev))))
        if check:
            self.check_status()
        return trimap_t

    def shard(self, n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
        """Contiguous block [lo, hi) of a batch of ``n_items`` images owned by ``rank``."""
        return shard_range(n_items, rank, world_size)


class PendingTrimaps:
    """A batch submitted with ``TrimapPath.submit``; ``result()`` waits for it (once)."""

    def __init__(self, path: TrimapPath, ticket: int, keep, tri_t, as_tensor: bool, nn_, ne_):
        self._path, self._ticket, self._keep = path, ticket, keep
        self._tri, self._as_tensor, self._nn, self._ne = tri_t, as_tensor, nn_, ne_
        self._done = ticket < 0

    def result(self):
        if not self._done:
            import torch
            with torch.cuda.device(self._path.dev):
                self._done = True
                nat.check(nat.lib().gg_trimap_path_host_wait(self._path.h.ptr, self._ticket))
            self._keep = None
        res = self._tri if self._as_tensor else self._tri.numpy()
        if self._nn is not None:
            return res, self._nn.numpy(), self._ne.numpy()
        return res

    def __del__(self):
        # an abandoned submission still has to release its ticket (and the buffers it writes to)
        if not self._done:
            try:
                self.result()
            except Exception:
                pass


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Images shard across GPUs with no exchange step: rank r owns a contiguous block; blocks
    differ by at most one image."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank / world_size")
    base, rem = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
