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
