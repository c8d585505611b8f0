"""
Trimap network on the B200 -- host-side mirror of the reference's ``src/gcn_grabcut/model.py``
interface for the residual GCN (ResGCNNet, build_model, probs_to_node_trimap,
project_to_pixels, _probs_to_trimap, label constants; reference lines 57-64, 421-557,
593-678).

``ResGCNNet`` is an ``nn.Module`` only as a parameter container: its ``state_dict`` has
exactly the reference's keys, so reference checkpoints load unchanged
(``inference.py:76-89``), but ``forward`` does not run PyTorch ops -- it hands the
parameters and the graph to libgcn_grabcut_b200.so (``gg_resgcn_forward``: CSR
gather/scatter message passing + tensor-core node transforms).  Inference only (eval-mode
semantics: dropout off, BatchNorm running statistics); there is no autograd and no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

try:
    import torch
    import torch.nn as nn
    _TORCH = True
except ImportError:                                   # pragma: no cover
    _TORCH = False

from . import _native as nat
from .graph_builder import N_EDGE_FEATS, N_NODE_FEATS, N_PRIOR_FEATS

TRIMAP_BG = 0
TRIMAP_FG = 1
TRIMAP_PROB_BG = 2
TRIMAP_PROB_FG = 3

CLASS_BG = 0
CLASS_UNK = 1
CLASS_FG = 2


class Data:
    """Minimal stand-in for ``torch_geometric.data.Data``: x, edge_index, edge_attr[, batch]."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, batch=None, **kwargs):
        self.x, self.edge_index, self.edge_attr, self.batch = x, edge_index, edge_attr, batch
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device, *args, **kwargs):
        for k, v in list(self.__dict__.items()):
            if _TORCH and torch.is_tensor(v):
                setattr(self, k, v.to(device, *args, **kwargs))
        return self


if _TORCH:

    def _data_device(data):
        x = data.x
        if not x.is_cuda:
            return None
        return x.device.index if x.device.index is not None else torch.cuda.current_device()

    def _prepare_graph(h, data):
        """What every forward(data) of the reference accepts (model.py:287-291, 384-390, 508-514) ->
        device tensors + the dst-sorted CSR the kernels walk (gg_coo_to_csr)."""
        x, edge_index = data.x, data.edge_index
        edge_attr = getattr(data, "edge_attr", None)
        batch = getattr(data, "batch", None)
        tdev = torch.device("cuda", h.device)
        x = x.to(tdev, torch.float32).contiguous()
        N, E = int(x.shape[0]), int(edge_index.shape[1])
        if x.shape[1] != N_NODE_FEATS:
            raise ValueError(f"x must be (N,{N_NODE_FEATS}), got {tuple(x.shape)}")
        ei = edge_index.to(tdev, torch.int64).contiguous()
        if edge_attr is None:
            edge_attr = torch.zeros(E, N_EDGE_FEATS, device=tdev)       # model.py:511-512
        ea = edge_attr.to(tdev, torch.float32).contiguous()
        if batch is None:
            goff = torch.tensor([0, N], dtype=torch.int64, device=tdev)
            n_graphs = 1
        else:
            b = batch.to(tdev, torch.int64)
            n_graphs = int(b.max().item()) + 1 if N > 0 else 1
            counts = torch.bincount(b, minlength=n_graphs)
            goff = torch.zeros(n_graphs + 1, dtype=torch.int64, device=tdev)
            goff[1:] = torch.cumsum(counts, 0)
            if N > 1 and bool((b[1:] < b[:-1]).any()):
                raise ValueError("data.batch must be sorted (graphs stored contiguously)")
        rowptr = torch.empty(N + 1, dtype=torch.int32, device=tdev)
        src = torch.empty(max(E, 1), dtype=torch.int32, device=tdev)
        eid = torch.empty(max(E, 1), dtype=torch.int32, device=tdev)
        stream = nat.current_stream(h.device)
        with torch.cuda.device(h.device):
            nat.check(nat.lib().gg_coo_to_csr(h.ptr, nat.ptr(ei), E, N, nat.ptr(rowptr), nat.ptr(src),
                                              nat.ptr(eid), C.c_void_p(stream)))
        return x, rowptr, src, eid, ea, goff, n_graphs, N, E, stream

    class _GraphLinear(nn.Module):
        """Parameter holder with PyG ``Linear`` naming (``.weight`` / ``.bias``)."""

        def __init__(self, in_dim, out_dim, bias, glorot=False):
            super().__init__()
            self.weight = nn.Parameter(torch.empty(out_dim, in_dim))
            if glorot:
                a = math.sqrt(6.0 / (in_dim + out_dim))
                nn.init.uniform_(self.weight, -a, a)
            else:
                nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            if bias:
                b = 1.0 / math.sqrt(in_dim)
                self.bias = nn.Parameter(torch.empty(out_dim).uniform_(-b, b))
            else:
                self.register_parameter("bias", None)

    class _GCNLayer(nn.Module):
        """Keys ``lin.weight`` and ``bias`` of a PyG GCNConv (reference model.py:480)."""

        def __init__(self, dim):
            super().__init__()
            self.lin = _GraphLinear(dim, dim, bias=False, glorot=True)
            self.bias = nn.Parameter(torch.zeros(dim))

    class _SAGELayer(nn.Module):
        """Keys ``lin_l.{weight,bias}`` and ``lin_r.weight`` of a PyG SAGEConv (model.py:483)."""

        def __init__(self, dim):
            super().__init__()
            self.lin_l = _GraphLinear(dim, dim, bias=True)
            self.lin_r = _GraphLinear(dim, dim, bias=False)

    class _InputNorm(nn.Module):
        def __init__(self, n, momentum=0.05):
            super().__init__()
            self.norm = nn.BatchNorm1d(n, momentum=momentum, affine=True)

    class _EdgeContext(nn.Module):
        def __init__(self, edge_dim, hidden):
            super().__init__()
            c = max(hidden // 2, 8)
            self.encode = nn.Sequential(nn.Linear(edge_dim, c), nn.GELU(), nn.Linear(c, c))
            self.to_gate = nn.Sequential(nn.LayerNorm(c), nn.Linear(c, hidden), nn.Sigmoid())

    class _GlobalContext(nn.Module):
        def __init__(self, hidden):
            super().__init__()
            self.attn = nn.Linear(hidden, 1)
            self.compress = nn.Linear(hidden, hidden // 2)
            self.expand = nn.Linear(hidden // 2, hidden)

    class ResGCNNet(nn.Module):
        """
        Residual GCN with jumping-knowledge fusion (reference model.py:421-590), executed by
        the CUDA library.  Constructor arguments, state-dict keys, ``eval`` / ``to`` /
        ``forward(data)`` / ``predict_probs`` / ``predict_trimap`` / ``layer_weights`` follow
        the reference.  ``hidden_channels`` must be a multiple of 32 in [32, 256].
        """

        def __init__(self, in_channels: int = N_NODE_FEATS, edge_channels: int = N_EDGE_FEATS,
                     hidden_channels: int = 128, n_layers: int = 6, n_classes: int = 3,
                     dropout: float = 0.15):
            super().__init__()
            if in_channels != N_NODE_FEATS or edge_channels != N_EDGE_FEATS or n_classes != 3:
                raise ValueError("the CUDA path is built for 19-d nodes, 5-d edges, 3 classes")
            if hidden_channels % 32 or not 32 <= hidden_channels <= 256:
                raise ValueError("hidden_channels must be a multiple of 32 in [32, 256]")
            self.n_classes, self.n_layers, self.dropout = n_classes, n_layers, dropout
            D = hidden_channels
            self.in_norm = _InputNorm(in_channels)
            self.input_proj = nn.Sequential(nn.Linear(in_channels, D), nn.LayerNorm(D), nn.GELU())
            q = max(D // 4, 8)
            self.prior_booster = nn.Sequential(nn.Linear(N_PRIOR_FEATS, q), nn.GELU(),
                                               nn.Linear(q, D), nn.Sigmoid())
            self.edge_ctx = _EdgeContext(edge_channels, D)
            self.gcn_layers = nn.ModuleList(_GCNLayer(D) for _ in range(n_layers))
            self.norms = nn.ModuleList(nn.LayerNorm(D) for _ in range(n_layers))
            self.sage = _SAGELayer(D)
            self.sage_norm = nn.LayerNorm(D)
            self.jk_logits = nn.Parameter(torch.zeros(n_layers + 2))
            self.ctx = _GlobalContext(D)
            self.fuse = nn.Sequential(nn.LayerNorm(D), nn.Linear(D, D), nn.GELU(), nn.Dropout(dropout))
            self.head = nn.Linear(D, n_classes)
            for m in self.modules():                       # reference _init_weights (model.py:501-506)
                if isinstance(m, nn.Linear):
                    nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
                    if m.bias is not None:
                        nn.init.zeros_(m.bias)
            self._gg_device: Optional[int] = None
            self._gg_dirty = True
            self._gg_version = -1
            self.requires_grad_(False)
            self.eval()

        # -- parameter plumbing
        def load_state_dict(self, state_dict, strict: bool = True, **kw):
            out = super().load_state_dict(state_dict, strict=strict, **kw)
            self._gg_dirty = True
            return out

        def _apply(self, fn, *a, **kw):
            out = super()._apply(fn, *a, **kw)
            self._gg_dirty = True
            p = next(self.parameters())
            self._gg_device = p.device.index if p.is_cuda else None
            if p.is_cuda and self._gg_device is None:
                self._gg_device = torch.cuda.current_device()
            return out

        def _handle(self, device: Optional[int] = None) -> "nat.Handle":
            if device is None:
                device = self._gg_device if self._gg_device is not None else torch.cuda.current_device() \
                    if torch.cuda.is_available() else None
            if device is None:
                raise nat.NativeError(nat.GG_ERR_CUDA, "ResGCNNet needs a CUDA device (no CPU fallback); "
                                                       "call model.to('cuda')")
            h = nat.handle(device)
            # in-place edits (p.data.copy_, an optimiser step over param_groups(),
            # vector_to_parameters) bump the tensors' version counters: reload on any change
            version = sum(int(t._version) for t in self.state_dict().values())
            if self._gg_dirty or h.weights_token is not self or version != self._gg_version:
                nat.load_state_dict(h, self.state_dict())
                h.weights_token = self
                self._gg_dirty = False
                self._gg_version = version
            return h

        # -- inference
        @torch.no_grad()
        def forward(self, data) -> "torch.Tensor":
            logits, _ = self._run(data, want_logits=True, want_probs=False)
            return logits

        def _run(self, data, want_logits: bool, want_probs: bool):
            h = self._handle(_data_device(data))
            x, rowptr, src, eid, ea, goff, n_graphs, N, E, stream = _prepare_graph(h, data)
            tdev = x.device
            logits = torch.empty(N, 3, dtype=torch.float32, device=tdev) if want_logits else None
            probs = torch.empty(N, 3, dtype=torch.float32, device=tdev) if want_probs else None
            with torch.cuda.device(h.device):
                nat.check(nat.lib().gg_resgcn_forward(h.ptr, nat.ptr(x), nat.ptr(rowptr), nat.ptr(src), nat.ptr(eid),
                                                      nat.ptr(ea), nat.ptr(goff), n_graphs, N, E, nat.ptr(logits),
                                                      nat.ptr(probs), C.c_void_p(stream)))
                h.check_status(stream)
            return logits, probs

        @torch.no_grad()
        def layer_weights(self) -> np.ndarray:
            return torch.softmax(self.jk_logits.detach(), dim=0).cpu().numpy()

        @torch.no_grad()
        def predict_probs(self, data) -> np.ndarray:
            self.eval()
            _, probs = self._run(data, want_logits=False, want_probs=True)
            return probs.float().cpu().numpy()

        @torch.no_grad()
        def predict_trimap(self, data, segments: np.ndarray, threshold_fg: float = 0.55,
                           threshold_bg: float = 0.55) -> np.ndarray:
            self.eval()
            _, probs = self._run(data, want_logits=False, want_probs=True)
            return _project_trimap_device(probs, segments, threshold_fg, threshold_bg)

        def param_groups(self, base_lr: float) -> list:
            """Layer-wise LR decay groups (reference model.py:559-590); training is out of scope."""
            n = self.n_layers
            groups = [{"params": list(g.parameters()) + list(nm.parameters()), "lr": base_lr * (0.8 ** (n - i))}
                      for i, (g, nm) in enumerate(zip(self.gcn_layers, self.norms))]
            groups.append({"params": list(self.in_norm.parameters()) + list(self.input_proj.parameters()) +
                           list(self.prior_booster.parameters()), "lr": base_lr * 0.5})
            groups.append({"params": list(self.edge_ctx.parameters()) + list(self.sage.parameters()) +
                           list(self.sage_norm.parameters()) + list(self.ctx.parameters()), "lr": base_lr * 0.9})
            groups.append({"params": [self.jk_logits] + list(self.fuse.parameters()) +
                           list(self.head.parameters()), "lr": base_lr})
            return groups

    # ------------------------------------------------------------------ variants
    class _VariantNet(nn.Module):
        """Parameter container + forward plumbing shared by GCNTrimapNet / GATTrimapNet: the
        state-dict has the reference's keys; forward runs gg_variant_forward (csrc/variants.cu)."""

        _variant = 0

        def _finish_init(self):
            self._gg_dirty = True
            self._gg_version = -1
            self._gg_device: Optional[int] = None
            self.requires_grad_(False)
            self.eval()

        def load_state_dict(self, state_dict, strict: bool = True, **kw):
            out = super().load_state_dict(state_dict, strict=strict, **kw)
            self._gg_dirty = True
            return out

        def _apply(self, fn, *a, **kw):
            out = super()._apply(fn, *a, **kw)
            self._gg_dirty = True
            p = next(self.parameters())
            self._gg_device = (p.device.index if p.device.index is not None else torch.cuda.current_device()) \
                if p.is_cuda else None
            return out

        def _tensor_keys(self):
            raise NotImplementedError

        def _handle(self, device: Optional[int] = None) -> "nat.Handle":
            if device is None:
                device = self._gg_device if self._gg_device is not None else \
                    (torch.cuda.current_device() if torch.cuda.is_available() else None)
            if device is None:
                raise nat.NativeError(nat.GG_ERR_CUDA, f"{type(self).__name__} needs a CUDA device (no CPU fallback); "
                                                       "call model.to('cuda')")
            h = nat.handle(device)
            sd = self.state_dict()
            version = sum(int(t._version) for t in sd.values())
            if self._gg_dirty or getattr(h, "variant_token", None) is not self or version != self._gg_version:
                nat.load_variant_state_dict(h, self._variant, self.hidden_channels, self.n_layers,
                                            getattr(self, "n_heads", 0), sd, self._tensor_keys())
                h.variant_token = self
                self._gg_dirty = False
                self._gg_version = version
            return h

        def _run(self, data, want_logits: bool, want_probs: bool):
            h = self._handle(_data_device(data))
            x, rowptr, src, eid, ea, goff, n_graphs, N, E, stream = _prepare_graph(h, data)
            tdev = x.device
            logits = torch.empty(N, 3, dtype=torch.float32, device=tdev) if want_logits else None
            probs = torch.empty(N, 3, dtype=torch.float32, device=tdev) if want_probs else None
            with torch.cuda.device(h.device):
                nat.check(nat.lib().gg_variant_forward(h.ptr, self._variant, nat.ptr(x), nat.ptr(rowptr), nat.ptr(src),
                                                       nat.ptr(eid), nat.ptr(ea), nat.ptr(goff), n_graphs, N, E,
                                                       nat.ptr(logits), nat.ptr(probs), C.c_void_p(stream)))
                h.check_status(stream)
            return logits, probs

        @torch.no_grad()
        def forward(self, data) -> "torch.Tensor":
            return self._run(data, want_logits=True, want_probs=False)[0]

        @torch.no_grad()
        def predict_probs(self, data) -> np.ndarray:
            self.eval()
            return self._run(data, want_logits=False, want_probs=True)[1].float().cpu().numpy()

        @torch.no_grad()
        def predict_trimap(self, data, segments: np.ndarray, threshold_fg: float = 0.55,
                           threshold_bg: float = 0.55) -> np.ndarray:
            self.eval()
            probs = self._run(data, want_logits=False, want_probs=True)[1]
            return _project_trimap_device(probs, segments, threshold_fg, threshold_bg)

    _BN = ("weight", "bias", "running_mean", "running_var")

    class _EdgeInjection(nn.Module):
        """Keys of EdgeInjectionLayer (model.py:142-162): proj.0 / proj.2."""

        def __init__(self, edge_dim, hidden):
            super().__init__()
            self.proj = nn.Sequential(nn.Linear(edge_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.Sigmoid())

    class _ResGCNBlock(nn.Module):
        """Keys of ResGCNBlock (model.py:216-223) with in_dim == out_dim (skip = Identity)."""

        def __init__(self, dim, edge_dim):
            super().__init__()
            self.conv = _GCNLayer(dim)
            self.bn = nn.BatchNorm1d(dim)
            self.skip = nn.Identity()
            self.edge_inject = _EdgeInjection(edge_dim, dim)

    class GCNTrimapNet(_VariantNet):
        """Baseline GCN (reference model.py:239-316) executed by the CUDA library: constructor
        arguments, state-dict keys, forward(data) / predict_probs / predict_trimap as the reference."""

        _variant = 1

        def __init__(self, in_channels: int = N_NODE_FEATS, edge_channels: int = N_EDGE_FEATS,
                     hidden_channels: int = 128, n_layers: int = 6, n_classes: int = 3, dropout: float = 0.2):
            super().__init__()
            if in_channels != N_NODE_FEATS or edge_channels != N_EDGE_FEATS or n_classes != 3:
                raise ValueError("the CUDA path is built for 19-d nodes, 5-d edges, 3 classes")
            if hidden_channels % 32 or not 32 <= hidden_channels <= 256:
                raise ValueError("hidden_channels must be a multiple of 32 in [32, 256]")
            D = hidden_channels
            self.n_classes, self.n_layers, self.hidden_channels = n_classes, n_layers, D
            self.in_norm = _InputNorm(in_channels)
            self.input_proj = nn.Sequential(nn.Linear(in_channels, D), nn.BatchNorm1d(D), nn.ReLU())
            self.blocks = nn.ModuleList(_ResGCNBlock(D, edge_channels) for _ in range(n_layers))
            self.head = nn.Sequential(nn.Linear(D * (n_layers + 1), D), nn.BatchNorm1d(D), nn.ReLU(),
                                      nn.Dropout(dropout), nn.Linear(D, D // 2), nn.ReLU(), nn.Linear(D // 2, n_classes))
            self._finish_init()

        def _tensor_keys(self):
            keys = [f"in_norm.norm.{k}" for k in _BN] + ["input_proj.0.weight", "input_proj.0.bias"]
            keys += [f"input_proj.1.{k}" for k in _BN]
            for i in range(self.n_layers):
                p = f"blocks.{i}."
                keys += [p + "conv.bias", p + "conv.lin.weight"] + [p + f"bn.{k}" for k in _BN]
                keys += [p + "edge_inject.proj.0.weight", p + "edge_inject.proj.0.bias",
                         p + "edge_inject.proj.2.weight", p + "edge_inject.proj.2.bias"]
            keys += ["head.0.weight", "head.0.bias"] + [f"head.1.{k}" for k in _BN]
            keys += ["head.4.weight", "head.4.bias", "head.6.weight", "head.6.bias"]
            return keys

    class _GATv2Layer(nn.Module):
        """Keys of a PyG GATv2Conv(in, D/H, heads=H, edge_dim=5, share_weights=False) (model.py:360-366)."""

        def __init__(self, dim, heads, edge_dim):
            super().__init__()
            self.att = nn.Parameter(torch.empty(1, heads, dim // heads))
            a = math.sqrt(6.0 / (heads + dim // heads))
            nn.init.uniform_(self.att, -a, a)
            self.bias = nn.Parameter(torch.zeros(dim))
            self.lin_l = _GraphLinear(dim, dim, bias=True, glorot=True)
            self.lin_r = _GraphLinear(dim, dim, bias=True, glorot=True)
            self.lin_edge = _GraphLinear(edge_dim, dim, bias=False, glorot=True)

    class GATTrimapNet(_VariantNet):
        """GATv2 variant (reference model.py:323-414) executed by the CUDA library.  ``n_heads`` must
        divide both 32 and ``hidden_channels``."""

        _variant = 2

        def __init__(self, in_channels: int = N_NODE_FEATS, edge_channels: int = N_EDGE_FEATS,
                     hidden_channels: int = 128, n_heads: int = 8, n_layers: int = 5, n_classes: int = 3,
                     dropout: float = 0.2):
            super().__init__()
            if in_channels != N_NODE_FEATS or edge_channels != N_EDGE_FEATS or n_classes != 3:
                raise ValueError("the CUDA path is built for 19-d nodes, 5-d edges, 3 classes")
            if hidden_channels % 32 or not 32 <= hidden_channels <= 256:
                raise ValueError("hidden_channels must be a multiple of 32 in [32, 256]")
            if n_heads < 1 or 32 % n_heads or hidden_channels % n_heads:
                raise ValueError("n_heads must divide 32 and hidden_channels")
            D = hidden_channels
            self.n_classes, self.n_heads, self.n_layers, self.hidden_channels = n_classes, n_heads, n_layers, D
            self.dropout = dropout
            self.in_norm = _InputNorm(in_channels)
            self.input_proj = nn.Sequential(nn.Linear(in_channels, D), nn.LayerNorm(D), nn.GELU())
            self.convs = nn.ModuleList(_GATv2Layer(D, n_heads, edge_channels) for _ in range(n_layers))
            self.lns = nn.ModuleList(nn.LayerNorm(D) for _ in range(n_layers))
            self.edge_gates = nn.ModuleList(_EdgeInjection(edge_channels, D) for _ in range(n_layers))
            self.skip_proj = nn.Linear(D, D, bias=False)
            self.ctx = _GlobalContext(D)
            self.head = nn.Sequential(nn.Linear(D, D), nn.GELU(), nn.Dropout(dropout), nn.Linear(D, n_classes))
            self._finish_init()

        def _tensor_keys(self):
            keys = [f"in_norm.norm.{k}" for k in _BN]
            keys += ["input_proj.0.weight", "input_proj.0.bias", "input_proj.1.weight", "input_proj.1.bias"]
            for i in range(self.n_layers):
                p, q = f"convs.{i}.", f"edge_gates.{i}.proj."
                keys += [p + "att", p + "bias", p + "lin_l.weight", p + "lin_l.bias", p + "lin_r.weight",
                         p + "lin_r.bias", p + "lin_edge.weight", f"lns.{i}.weight", f"lns.{i}.bias",
                         q + "0.weight", q + "0.bias", q + "2.weight", q + "2.bias"]
            keys += ["skip_proj.weight", "ctx.attn.weight", "ctx.attn.bias", "ctx.compress.weight",
                     "ctx.compress.bias", "ctx.expand.weight", "ctx.expand.bias", "head.0.weight", "head.0.bias",
                     "head.3.weight", "head.3.bias"]
            return keys

    def build_model(variant: str = "resgcn", in_channels: int = N_NODE_FEATS,
                    edge_channels: int = N_EDGE_FEATS, hidden_channels: int = 128, n_layers: int = 6,
                    n_classes: int = 3, dropout: float = 0.2):
        """Factory (reference model.py:593-620): resgcn (the path's network) | gcn | gat."""
        if variant == "resgcn":
            return ResGCNNet(in_channels, edge_channels, hidden_channels, n_layers, n_classes, dropout)
        if variant == "gcn":
            return GCNTrimapNet(in_channels, edge_channels, hidden_channels, n_layers, n_classes, dropout)
        if variant == "gat":
            return GATTrimapNet(in_channels, edge_channels, hidden_channels, n_layers=n_layers,
                                n_classes=n_classes, dropout=dropout)
        raise ValueError(f"Unknown variant '{variant}'. Choose: resgcn | gcn | gat")

    def _project_trimap_device(probs_t, segments: np.ndarray, thr_fg: float, thr_bg: float) -> np.ndarray:
        dev = probs_t.device.index
        h = nat.handle(dev)
        seg = torch.from_numpy(np.ascontiguousarray(segments, dtype=np.int32)).to(probs_t.device)
        H, W = seg.shape
        goff = torch.tensor([0, probs_t.shape[0]], dtype=torch.int64, device=probs_t.device)
        tri = torch.empty(H, W, dtype=torch.uint8, device=probs_t.device)
        with torch.cuda.device(dev):
            nat.check(nat.lib().gg_project_trimap(h.ptr, nat.ptr(seg), nat.ptr(probs_t.contiguous()),
                                                  nat.ptr(goff), 1, H, W, float(thr_fg), float(thr_bg),
                                                  nat.ptr(tri), C.c_void_p(nat.current_stream(dev))))
        return tri.cpu().numpy()


def probs_to_node_trimap(probs: np.ndarray, threshold_fg: float = 0.55,
                         threshold_bg: float = 0.55) -> np.ndarray:
    """Per-region GrabCut labels from class probabilities (reference model.py:623-645); N values,
    host-side (the pixel-level projection is what runs on the GPU)."""
    bg_p, fg_p = probs[:, CLASS_BG], probs[:, CLASS_FG]
    labels = np.where(fg_p > bg_p, TRIMAP_PROB_FG, TRIMAP_PROB_BG).astype(np.uint8)
    labels[bg_p >= threshold_bg] = TRIMAP_BG
    labels[fg_p >= threshold_fg] = TRIMAP_FG
    return labels


def project_to_pixels(node_values: np.ndarray, segments: np.ndarray) -> np.ndarray:
    """values[segments] with zero padding (reference model.py:648-661); a plain gather kept on
    the host for API parity -- the trimap path itself gathers inside gg_refine_trimap."""
    need = int(segments.max()) + 1
    values = node_values
    if values.shape[0] < need:
        pad = np.zeros((need - values.shape[0], *values.shape[1:]), dtype=values.dtype)
        values = np.concatenate([values, pad], axis=0)
    return values[segments]


def _probs_to_trimap(probs: np.ndarray, segments: np.ndarray, threshold_fg: float,
                     threshold_bg: float) -> np.ndarray:
    """Pixel trimap from region probabilities (reference model.py:664-678) via gg_project_trimap."""
    import torch
    p = torch.from_numpy(np.ascontiguousarray(probs, dtype=np.float32)).to("cuda")
    return _project_trimap_device(p, segments, threshold_fg, threshold_bg)
