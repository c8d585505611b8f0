"""
CPU restatement of the reference trimap network ``ResGCNNet`` (eval mode, fp32).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
/root/reference/src/gcn_grabcut/model.py:69-139, 165-213, 449-557, 623-678 and the PyG
layer semantics restated in oracle/thirdparty.py.  Functional: the network is a plain
``state_dict`` (the keys a reference checkpoint holds, inference.py:76-89), so the same
tensors can be handed to the reference, to this port and to the CUDA path.
"""

from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

TRIMAP_BG, TRIMAP_FG, TRIMAP_PROB_BG, TRIMAP_PROB_FG = 0, 1, 2, 3     # model.py:57-60
CLASS_BG, CLASS_UNK, CLASS_FG = 0, 1, 2                               # model.py:62-64
N_PRIOR_FEATS = 3


def infer_dims(state: Dict[str, torch.Tensor]):
    """(D, n_layers) recovered from the checkpoint keys, as inference.py:81-86 does."""
    D = int(state["input_proj.0.weight"].shape[0])
    n = sum(1 for k in state if k.startswith("gcn_layers.") and k.endswith(".bias"))
    return D, n


from gcn_grabcut_b200.synthetic import random_state_dict  # noqa: E402,F401  (seeded random-init weights: shared with bench.py)


def _scatter_mean(src: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    """model.py:69-74 -- grouped mean, empty groups -> 0 (count clamped to 1)."""
    out = torch.zeros(n, src.size(1), dtype=src.dtype)
    out.index_add_(0, index, src)
    cnt = torch.bincount(index, minlength=n).to(src.dtype).clamp(min=1)
    return out / cnt.unsqueeze(1)


def _graph_softmax(scores: torch.Tensor, batch: Optional[torch.Tensor]) -> torch.Tensor:
    """model.py:90-108 -- softmax over the nodes of each graph."""
    if batch is None:
        return torch.softmax(scores.float(), dim=0)
    g = int(batch.max()) + 1
    s = scores.float()
    peak = torch.full((g, 1), float("-inf")).index_reduce(0, batch, s, "amax", include_self=True)
    ex = torch.exp(s - peak[batch])
    tot = torch.zeros_like(peak).index_add_(0, batch, ex)
    return ex / (tot[batch] + 1e-12)


def _gcn_conv(x, edge_index, weight, bias):
    """PyG GCNConv defaults (thirdparty.GCNConv); model.py:480, 524."""
    n = x.size(0)
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    loop = torch.arange(n, dtype=src.dtype)
    src, dst = torch.cat([src[keep], loop]), torch.cat([dst[keep], loop])
    deg = torch.zeros(n, dtype=x.dtype).scatter_add_(0, dst, torch.ones(dst.numel(), dtype=x.dtype))
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    w = dis[src] * dis[dst]
    h = F.linear(x, weight)
    out = torch.zeros(n, h.size(1), dtype=h.dtype).index_add_(0, dst, h[src] * w.unsqueeze(1))
    return out + bias


def _sage_conv(x, edge_index, w_l, b_l, w_r):
    """PyG SAGEConv defaults (thirdparty.SAGEConv); model.py:483, 530."""
    agg = _scatter_mean(x[edge_index[0]], edge_index[1], x.size(0))
    return F.linear(agg, w_l, b_l) + F.linear(x, w_r)


@torch.no_grad()
def resgcn_forward(state: Dict[str, torch.Tensor], x: torch.Tensor, edge_index: torch.Tensor,
                   edge_attr: torch.Tensor, batch: Optional[torch.Tensor] = None,
                   return_stages: bool = False):
    """ResGCNNet.forward in eval mode (model.py:508-536) -> logits (N, 3)."""
    s = {k: v.float() if v.is_floating_point() else v for k, v in state.items()}
    D, n_layers = infer_dims(s)
    x = x.float()
    n = x.size(0)
    stages = {}

    prior = x[:, -N_PRIOR_FEATS:]                                              # :516
    xn = F.batch_norm(x, s["in_norm.norm.running_mean"], s["in_norm.norm.running_var"],
                      s["in_norm.norm.weight"], s["in_norm.norm.bias"], False, 0.0, 1e-5)
    h = F.gelu(F.layer_norm(F.linear(xn, s["input_proj.0.weight"], s["input_proj.0.bias"]),
                            (D,), s["input_proj.1.weight"], s["input_proj.1.bias"], 1e-5))
    boost = torch.sigmoid(F.linear(F.gelu(F.linear(prior, s["prior_booster.0.weight"],
                                                   s["prior_booster.0.bias"])),
                                   s["prior_booster.2.weight"], s["prior_booster.2.bias"]))
    h = h * (1.0 + boost)                                                      # :518

    c = s["edge_ctx.encode.0.weight"].shape[0]
    enc = F.linear(F.gelu(F.linear(edge_attr.float(), s["edge_ctx.encode.0.weight"],
                                   s["edge_ctx.encode.0.bias"])),
                   s["edge_ctx.encode.2.weight"], s["edge_ctx.encode.2.bias"])
    ctx = _scatter_mean(enc, edge_index[1], n)                                 # :138
    gate = torch.sigmoid(F.linear(F.layer_norm(ctx, (c,), s["edge_ctx.to_gate.0.weight"],
                                               s["edge_ctx.to_gate.0.bias"], 1e-5),
                                  s["edge_ctx.to_gate.1.weight"], s["edge_ctx.to_gate.1.bias"]))
    stages["h0"], stages["gate"] = h, gate

    states = [h]
    for i in range(n_layers):                                                  # :523-528
        u = _gcn_conv(F.layer_norm(h, (D,), s[f"norms.{i}.weight"], s[f"norms.{i}.bias"], 1e-5),
                      edge_index, s[f"gcn_layers.{i}.lin.weight"], s[f"gcn_layers.{i}.bias"])
        h = h + F.gelu(u * gate)
        states.append(h)
    stages["h_last"] = h
    sg = _sage_conv(h, edge_index, s["sage.lin_l.weight"], s["sage.lin_l.bias"],
                    s["sage.lin_r.weight"])
    states.append(F.gelu(F.layer_norm(sg, (D,), s["sage_norm.weight"], s["sage_norm.bias"], 1e-5)))

    w = torch.softmax(s["jk_logits"], dim=0)                                   # :532-533
    z = torch.stack(states, dim=0).mul(w[:, None, None]).sum(dim=0)
    stages["z"] = z

    a = _graph_softmax(F.linear(z, s["ctx.attn.weight"], s["ctx.attn.bias"]), batch)   # :165-188
    if batch is None:
        g = (a * z).sum(dim=0, keepdim=True)
    else:
        ng = int(batch.max()) + 1
        g = torch.zeros(ng, D).index_add_(0, batch, a * z)[batch]
    g = torch.sigmoid(F.linear(F.relu(F.linear(g, s["ctx.compress.weight"], s["ctx.compress.bias"])),
                               s["ctx.expand.weight"], s["ctx.expand.bias"]))
    z = z * g
    f = F.gelu(F.linear(F.layer_norm(z, (D,), s["fuse.0.weight"], s["fuse.0.bias"], 1e-5),
                        s["fuse.1.weight"], s["fuse.1.bias"]))
    logits = F.linear(f, s["head.weight"], s["head.bias"])                     # :536
    if return_stages:
        return logits, stages
    return logits


@torch.no_grad()
def predict_probs(state, x, edge_index, edge_attr, batch=None) -> np.ndarray:
    """ResGCNNet.predict_probs (model.py:543-546) -> float32 (N,3), columns [BG, UNK, FG]."""
    return F.softmax(resgcn_forward(state, x, edge_index, edge_attr, batch), dim=-1).float().numpy()


def probs_to_node_trimap(probs: np.ndarray, thr_fg: float = 0.55, thr_bg: float = 0.55) -> np.ndarray:
    """model.py:623-645 -- FG overrides BG when both thresholds are met."""
    bg, fg = probs[:, CLASS_BG], probs[:, CLASS_FG]
    lab = np.where(fg > bg, TRIMAP_PROB_FG, TRIMAP_PROB_BG).astype(np.uint8)
    lab[bg >= thr_bg] = TRIMAP_BG
    lab[fg >= thr_fg] = TRIMAP_FG
    return lab


def project_to_pixels(values: np.ndarray, seg: np.ndarray) -> np.ndarray:
    """model.py:648-661 -- values[seg], zero-padded if seg references missing rows."""
    need = int(seg.max()) + 1
    if values.shape[0] < need:
        pad = np.zeros((need - values.shape[0], *values.shape[1:]), dtype=values.dtype)
        values = np.concatenate([values, pad], axis=0)
    return values[seg]


def probs_to_trimap(probs: np.ndarray, seg: np.ndarray, thr_fg: float, thr_bg: float) -> np.ndarray:
    """model.py:664-678 -- node labels gathered through the label map (PR_BGD padding)."""
    lab = probs_to_node_trimap(probs, thr_fg, thr_bg)
    need = int(seg.max()) + 1
    if lab.shape[0] < need:
        lab = np.concatenate([lab, np.full(need - lab.shape[0], TRIMAP_PROB_BG, dtype=np.uint8)])
    return lab[seg].astype(np.uint8)


# ----------------------------------------------------------------------------- GCNTrimapNet (baseline variant)
def _bn_eval(x, state, prefix):
    """BatchNorm1d in eval mode (running statistics, eps 1e-5)."""
    return (x - state[prefix + ".running_mean"]) / torch.sqrt(state[prefix + ".running_var"] + 1e-5) * \
        state[prefix + ".weight"] + state[prefix + ".bias"]


def gcn_trimap_dims(state: Dict[str, torch.Tensor]):
    D = int(state["input_proj.0.weight"].shape[0])
    n = sum(1 for k in state if k.startswith("blocks.") and k.endswith(".conv.bias"))
    return D, n


@torch.no_grad()
def gcn_trimap_forward(state: Dict[str, torch.Tensor], x: torch.Tensor, edge_index: torch.Tensor,
                       edge_attr: torch.Tensor) -> torch.Tensor:
    """GCNTrimapNet.forward in eval mode (model.py:290-304) with ResGCNBlock (:216-233) and
    EdgeInjectionLayer (:142-162): logits (N, 3)."""
    D, n = gcn_trimap_dims(state)
    x, edge_attr = x.float(), edge_attr.float()
    N = x.size(0)
    h = F.relu(_bn_eval(F.linear(_bn_eval(x, state, "in_norm.norm"), state["input_proj.0.weight"],
                                 state["input_proj.0.bias"]), state, "input_proj.1"))
    all_h = [h]
    for i in range(n):
        p = f"blocks.{i}."
        c = _gcn_conv(h, edge_index, state[p + "conv.lin.weight"], state[p + "conv.bias"])
        c = F.relu(_bn_eval(c, state, p + "bn")) + h                       # skip = Identity (in_dim == out_dim)
        e = torch.sigmoid(F.linear(F.relu(F.linear(edge_attr, state[p + "edge_inject.proj.0.weight"],
                                                   state[p + "edge_inject.proj.0.bias"])),
                                   state[p + "edge_inject.proj.2.weight"], state[p + "edge_inject.proj.2.bias"]))
        h = c * _scatter_mean(e, edge_index[1], N)
        all_h.append(h)
    t = F.relu(_bn_eval(F.linear(torch.cat(all_h, -1), state["head.0.weight"], state["head.0.bias"]), state, "head.1"))
    t = F.relu(F.linear(t, state["head.4.weight"], state["head.4.bias"]))
    return F.linear(t, state["head.6.weight"], state["head.6.bias"])


def random_gcn_trimap_state(hidden: int = 128, n_layers: int = 6, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random GCNTrimapNet state-dict with the reference's keys (non-trivial norms / biases)."""
    g = torch.Generator().manual_seed(seed)
    D = hidden

    def lin(o, i):
        return torch.randn(o, i, generator=g) * (2.0 / i) ** 0.5

    def vec(n, scale=0.1, shift=0.0):
        return torch.randn(n, generator=g) * scale + shift

    def bn(prefix, n, s):
        s[prefix + ".weight"], s[prefix + ".bias"] = vec(n, 0.1, 1.0), vec(n)
        s[prefix + ".running_mean"], s[prefix + ".running_var"] = vec(n, 0.2, 0.1), vec(n, 0.1, 0.6).abs() + 0.05
        s[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    s: Dict[str, torch.Tensor] = {}
    bn("in_norm.norm", 19, s)
    s["input_proj.0.weight"], s["input_proj.0.bias"] = lin(D, 19), vec(D)
    bn("input_proj.1", D, s)
    for i in range(n_layers):
        p = f"blocks.{i}."
        s[p + "conv.bias"] = vec(D)
        s[p + "conv.lin.weight"] = (torch.rand(D, D, generator=g) * 2 - 1) * (6.0 / (2 * D)) ** 0.5
        bn(p + "bn", D, s)
        s[p + "edge_inject.proj.0.weight"], s[p + "edge_inject.proj.0.bias"] = lin(D, 5), vec(D)
        s[p + "edge_inject.proj.2.weight"], s[p + "edge_inject.proj.2.bias"] = lin(D, D), vec(D)
    s["head.0.weight"], s["head.0.bias"] = lin(D, D * (n_layers + 1)), vec(D)
    bn("head.1", D, s)
    s["head.4.weight"], s["head.4.bias"] = lin(D // 2, D), vec(D // 2)
    s["head.6.weight"], s["head.6.bias"] = lin(3, D // 2), vec(3)
    return {k: v.contiguous() for k, v in s.items()}


# ----------------------------------------------------------------------------- GATTrimapNet (attention variant)
def _gatv2_conv(x, edge_index, edge_attr, state, p, heads):
    """PyG GATv2Conv as the reference configures it (model.py:360-366; thirdparty.GATv2Conv)."""
    n = x.size(0)
    xl = F.linear(x, state[p + "lin_l.weight"], state[p + "lin_l.bias"])
    xr = F.linear(x, state[p + "lin_r.weight"], state[p + "lin_r.bias"])
    D = xl.size(1)
    C_ = D // heads
    src, dst = edge_index[0], edge_index[1]
    keep = src != dst
    src, dst, ea = src[keep], dst[keep], edge_attr[keep]
    loop_attr = _scatter_mean(ea, dst, n)                                          # fill_value="mean"
    loop = torch.arange(n, dtype=src.dtype)
    src, dst, ea = torch.cat([src, loop]), torch.cat([dst, loop]), torch.cat([ea, loop_attr])
    m = F.leaky_relu(xl[src] + xr[dst] + F.linear(ea, state[p + "lin_edge.weight"]), 0.2).view(-1, heads, C_)
    score = (m * state[p + "att"].view(1, heads, C_)).sum(-1)
    out = torch.zeros(n, heads, C_)
    for i in range(n):                                                              # softmax over the edges entering i
        sel = (dst == i).nonzero().flatten()
        s = score[sel]
        e = torch.exp(s - s.max(0, keepdim=True).values)
        a = e / (e.sum(0, keepdim=True) + 1e-16)
        out[i] = (xl[src[sel]].view(-1, heads, C_) * a.unsqueeze(-1)).sum(0)
    return out.view(n, D) + state[p + "bias"]


def gat_trimap_dims(state: Dict[str, torch.Tensor]):
    D = int(state["input_proj.0.weight"].shape[0])
    n = sum(1 for k in state if k.startswith("convs.") and k.endswith(".att"))
    heads = int(state["convs.0.att"].shape[1])
    return D, n, heads


@torch.no_grad()
def gat_trimap_forward(state: Dict[str, torch.Tensor], x: torch.Tensor, edge_index: torch.Tensor,
                       edge_attr: torch.Tensor, batch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GATTrimapNet.forward in eval mode (model.py:384-404): logits (N, 3)."""
    D, n, heads = gat_trimap_dims(state)
    x, edge_attr = x.float(), edge_attr.float()
    N = x.size(0)
    h = F.gelu(F.layer_norm(F.linear(_bn_eval(x, state, "in_norm.norm"), state["input_proj.0.weight"],
                                     state["input_proj.0.bias"]), (D,), state["input_proj.1.weight"],
                            state["input_proj.1.bias"], 1e-5))
    skip = F.linear(h, state["skip_proj.weight"])
    for i in range(n):
        c = _gatv2_conv(h, edge_index, edge_attr, state, f"convs.{i}.", heads)
        c = F.gelu(F.layer_norm(c, (D,), state[f"lns.{i}.weight"], state[f"lns.{i}.bias"], 1e-5))
        p = f"edge_gates.{i}."
        e = torch.sigmoid(F.linear(F.relu(F.linear(edge_attr, state[p + "proj.0.weight"], state[p + "proj.0.bias"])),
                                   state[p + "proj.2.weight"], state[p + "proj.2.bias"]))
        h = c * _scatter_mean(e, edge_index[1], N)
    h = h + skip
    w = _graph_softmax(F.linear(h, state["ctx.attn.weight"], state["ctx.attn.bias"]), batch)   # model.py:176-188
    if batch is None:
        g = (w * h).sum(0, keepdim=True).expand(N, -1)
    else:
        g = torch.zeros(int(batch.max()) + 1, D).index_add_(0, batch, w * h)[batch]
    g = torch.sigmoid(F.linear(F.relu(F.linear(g, state["ctx.compress.weight"], state["ctx.compress.bias"])),
                               state["ctx.expand.weight"], state["ctx.expand.bias"]))
    h = h * g
    t = F.gelu(F.linear(h, state["head.0.weight"], state["head.0.bias"]))
    return F.linear(t, state["head.3.weight"], state["head.3.bias"])


def random_gat_trimap_state(hidden: int = 128, n_heads: int = 8, n_layers: int = 5, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded random GATTrimapNet state-dict with the reference's keys."""
    g = torch.Generator().manual_seed(seed)
    D = hidden

    def lin(o, i):
        return torch.randn(o, i, generator=g) * (2.0 / i) ** 0.5

    def vec(n, scale=0.1, shift=0.0):
        return torch.randn(n, generator=g) * scale + shift

    s: Dict[str, torch.Tensor] = {}
    s["in_norm.norm.weight"], s["in_norm.norm.bias"] = vec(19, 0.1, 1.0), vec(19)
    s["in_norm.norm.running_mean"], s["in_norm.norm.running_var"] = vec(19, 0.2, 0.1), vec(19, 0.1, 0.6).abs() + 0.05
    s["in_norm.norm.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    s["input_proj.0.weight"], s["input_proj.0.bias"] = lin(D, 19), vec(D)
    s["input_proj.1.weight"], s["input_proj.1.bias"] = vec(D, 0.1, 1.0), vec(D)
    for i in range(n_layers):
        p = f"convs.{i}."
        s[p + "att"] = torch.randn(1, n_heads, D // n_heads, generator=g) * 0.5
        s[p + "bias"] = vec(D)
        s[p + "lin_l.weight"], s[p + "lin_l.bias"] = lin(D, D) * 0.7, vec(D)
        s[p + "lin_r.weight"], s[p + "lin_r.bias"] = lin(D, D) * 0.7, vec(D)
        s[p + "lin_edge.weight"] = lin(D, 5)
        s[f"lns.{i}.weight"], s[f"lns.{i}.bias"] = vec(D, 0.1, 1.0), vec(D)
        q = f"edge_gates.{i}.proj."
        s[q + "0.weight"], s[q + "0.bias"] = lin(D, 5), vec(D)
        s[q + "2.weight"], s[q + "2.bias"] = lin(D, D), vec(D)
    s["skip_proj.weight"] = lin(D, D) * 0.5
    s["ctx.attn.weight"], s["ctx.attn.bias"] = lin(1, D), vec(1)
    s["ctx.compress.weight"], s["ctx.compress.bias"] = lin(D // 2, D), vec(D // 2)
    s["ctx.expand.weight"], s["ctx.expand.bias"] = lin(D, D // 2), vec(D)
    s["head.0.weight"], s["head.0.bias"] = lin(D, D), vec(D)
    s["head.3.weight"], s["head.3.bias"] = lin(3, D), vec(3)
    return {k: v.contiguous() for k, v in s.items()}
