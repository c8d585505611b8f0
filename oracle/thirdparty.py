"""
Restatements of the third-party arithmetic the reference calls on the trimap path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  None of these packages is vendored
under /root/reference nor installed in this image; the reference leaves their
versions unpinned (pyproject.toml:10-23: bare ``scikit-image``, ``torch-geometric``).
The formulas below restate the published algorithms of

* scikit-image >= 0.19  ``skimage.color.rgb2lab`` (= ``xyz2lab(rgb2xyz(rgb))``),
  ``skimage.color.rgb2hsv``, ``skimage.segmentation.find_boundaries``
  -- call sites: graph_builder.py:148, :149, :211
* PyTorch Geometric >= 2.0  ``GCNConv`` (add_self_loops, normalize, bias),
  ``SAGEConv`` (aggr="mean", root_weight, bias), ``Data``, ``Batch``
  -- call sites: model.py:480, :483, :524, :530; pipeline.py:304; trainer.py:279

Parity against real scikit-image / PyG builds: UNPINNED (not installable here).
"""

from __future__ import annotations

import math
from typing import Optional

import numpy as np

# ----------------------------------------------------------------------------
# scikit-image colour conversions (float64 arithmetic on a uint8 image / 255)
# ----------------------------------------------------------------------------

XYZ_FROM_RGB = np.array(
    [[0.412453, 0.357580, 0.180423],
     [0.212671, 0.715160, 0.072169],
     [0.019334, 0.119193, 0.950227]], dtype=np.float64)

# CIE XYZ tristimulus values of D65, 2-degree observer (skimage default)
D65_WHITE = np.array([0.95047, 1.0, 1.08883], dtype=np.float64)


def img_as_float(image: np.ndarray) -> np.ndarray:
    """skimage.util.img_as_float for the dtypes the path sees."""
    image = np.asarray(image)
    if image.dtype == np.uint8:
        return image.astype(np.float64) / 255.0
    if image.dtype.kind == "f":
        return image
    raise TypeError(f"img_as_float restatement: unsupported dtype {image.dtype}")


def rgb2xyz(rgb: np.ndarray) -> np.ndarray:
    arr = np.array(img_as_float(rgb), dtype=np.float64, copy=True)
    mask = arr > 0.04045
    arr[mask] = np.power((arr[mask] + 0.055) / 1.055, 2.4)
    arr[~mask] /= 12.92
    return arr @ XYZ_FROM_RGB.T


def xyz2lab(xyz: np.ndarray) -> np.ndarray:
    arr = np.asarray(xyz, dtype=np.float64) / D65_WHITE
    mask = arr > 0.008856
    arr[mask] = np.cbrt(arr[mask])
    arr[~mask] = 7.787 * arr[~mask] + 16.0 / 116.0
    x, y, z = arr[..., 0], arr[..., 1], arr[..., 2]
    L = (116.0 * y) - 16.0
    a = 500.0 * (x - y)
    b = 200.0 * (y - z)
    return np.concatenate([c[..., np.newaxis] for c in (L, a, b)], axis=-1)


def rgb2lab(rgb: np.ndarray) -> np.ndarray:
    """sRGB uint8 (H,W,3) -> CIELAB float64, D65 / 2 deg."""
    return xyz2lab(rgb2xyz(rgb))


def rgb2hsv(rgb: np.ndarray) -> np.ndarray:
    """sRGB uint8 (H,W,3) -> HSV float64 with h,s,v in [0,1]."""
    arr = np.asarray(img_as_float(rgb), dtype=np.float64)
    out = np.empty_like(arr)

    out_v = arr.max(-1)
    delta = np.ptp(arr, axis=-1)

    old = np.seterr(invalid="ignore", divide="ignore")
    try:
        out_s = delta / out_v
        out_s[delta == 0.0] = 0.0

        # red is max, then green is max, then blue is max: later masks overwrite
        idx = arr[..., 0] == out_v
        out[idx, 0] = (arr[idx, 1] - arr[idx, 2]) / delta[idx]
        idx = arr[..., 1] == out_v
        out[idx, 0] = 2.0 + (arr[idx, 2] - arr[idx, 0]) / delta[idx]
        idx = arr[..., 2] == out_v
        out[idx, 0] = 4.0 + (arr[idx, 0] - arr[idx, 1]) / delta[idx]
        out_h = (out[..., 0] / 6.0) % 1.0
        out_h[delta == 0.0] = 0.0
    finally:
        np.seterr(**old)

    out[..., 0] = out_h
    out[..., 1] = out_s
    out[..., 2] = out_v
    out[np.isnan(out)] = 0
    return out


def find_boundaries(label_img: np.ndarray, connectivity: int = 1,
                    mode: str = "thick", background: int = 0) -> np.ndarray:
    """
    skimage.segmentation.find_boundaries: grey dilation != grey erosion under the
    (2*ndim)-neighbourhood cross; ``mode="inner"`` keeps only non-background pixels.
    The morphology pads by reflection, so the image frame itself is not a boundary.
    """
    from scipy import ndimage as ndi
    if mode not in ("thick", "inner"):
        raise NotImplementedError(f"find_boundaries restatement: mode={mode!r}")
    label_img = np.asarray(label_img)
    footprint = ndi.generate_binary_structure(label_img.ndim, connectivity)
    dil = ndi.grey_dilation(label_img, footprint=footprint, mode="reflect")
    ero = ndi.grey_erosion(label_img, footprint=footprint, mode="reflect")
    boundaries = dil != ero
    if mode == "inner":
        boundaries &= label_img != background
    return boundaries


def mark_boundaries(image, label_img, color=(1, 1, 0), **_):
    """Visualisation helper (graph_builder.py:353); not on the path."""
    out = img_as_float(image).copy()
    out[find_boundaries(label_img, mode="thick")] = color
    return out


# ----------------------------------------------------------------------------
# PyTorch Geometric layers / containers
# ----------------------------------------------------------------------------

try:
    import torch
    import torch.nn as nn
    _TORCH = True
except ImportError:                                  # pragma: no cover
    _TORCH = False


if _TORCH:

    class PyGLinear(nn.Module):
        """
        torch_geometric.nn.dense.linear.Linear.  Deliberately NOT a subclass of
        ``nn.Linear``: the reference's ``_init_weights`` (model.py:501-506) only
        re-initialises ``nn.Linear`` modules, so PyG layers keep their own defaults.
        """

        def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                     weight_initializer: Optional[str] = None):
            super().__init__()
            self.in_channels, self.out_channels = in_channels, out_channels
            self.weight_initializer = weight_initializer
            self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
            self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
            self.reset_parameters()

        def reset_parameters(self):
            if self.weight_initializer == "glorot":
                a = math.sqrt(6.0 / (self.in_channels + self.out_channels))
                nn.init.uniform_(self.weight, -a, a)
            else:                                   # PyG default: kaiming_uniform(a=sqrt(5))
                nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            if self.bias is not None:
                bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
                nn.init.uniform_(self.bias, -bound, bound)

        def forward(self, x):
            return torch.nn.functional.linear(x, self.weight, self.bias)

    class GCNConv(nn.Module):
        """
        PyG GCNConv with defaults (improved=False, add_self_loops=True, normalize=True,
        bias=True):  x' = lin(x) (no bias);  self loops removed and re-added once per
        node;  deg_i = #edges with dst i (incl. loop);  w_e = deg[src]^-1/2 deg[dst]^-1/2;
        out_i = sum_{e: dst=i} w_e x'_src + bias.
        """

        def __init__(self, in_channels: int, out_channels: int):
            super().__init__()
            self.lin = PyGLinear(in_channels, out_channels, bias=False,
                                 weight_initializer="glorot")
            self.bias = nn.Parameter(torch.zeros(out_channels))

        def forward(self, x, edge_index):
            n = x.size(0)
            src, dst = edge_index[0], edge_index[1]
            keep = src != dst
            loop = torch.arange(n, device=x.device, dtype=src.dtype)
            src = torch.cat([src[keep], loop])
            dst = torch.cat([dst[keep], loop])
            w = torch.ones(src.numel(), device=x.device, dtype=x.dtype)
            deg = torch.zeros(n, device=x.device, dtype=x.dtype).scatter_add_(0, dst, w)
            dis = deg.pow(-0.5)
            dis = dis.masked_fill(dis == float("inf"), 0.0)
            w = dis[src] * w * dis[dst]
            h = self.lin(x)
            out = torch.zeros(n, h.size(1), device=x.device, dtype=h.dtype)
            out.index_add_(0, dst, h[src] * w.unsqueeze(1))
            return out + self.bias

    class SAGEConv(nn.Module):
        """
        PyG SAGEConv with defaults (aggr="mean", root_weight=True, bias=True):
        out_i = lin_l(mean_{j->i} x_j) + lin_r(x_i); lin_l carries the bias, lin_r has
        none; no self loops; a node with no in-edges aggregates to zero.
        """

        def __init__(self, in_channels: int, out_channels: int):
            super().__init__()
            self.lin_l = PyGLinear(in_channels, out_channels, bias=True)
            self.lin_r = PyGLinear(in_channels, out_channels, bias=False)

        def forward(self, x, edge_index):
            n = x.size(0)
            src, dst = edge_index[0], edge_index[1]
            agg = torch.zeros(n, x.size(1), device=x.device, dtype=x.dtype)
            agg.index_add_(0, dst, x[src])
            cnt = torch.zeros(n, device=x.device, dtype=x.dtype).scatter_add_(
                0, dst, torch.ones(dst.numel(), device=x.device, dtype=x.dtype))
            agg = agg / cnt.clamp(min=1).unsqueeze(1)
            return self.lin_l(agg) + self.lin_r(x)

    class GATv2Conv(nn.Module):
        """
        PyG GATv2Conv (Brody et al., "How attentive are graph attention networks?") as the reference
        constructs it (model.py:360-366): heads=H, concat=True, edge_dim given, share_weights=False,
        negative_slope=0.2, add_self_loops=True, fill_value="mean", bias=True, no residual.
        Parameters: lin_l / lin_r (PyG Linear with bias, glorot), lin_edge (no bias), att [1,H,C],
        bias [H*C].  forward: existing self loops are removed and one loop per node is appended
        whose edge attribute is the mean of the attributes of the node's incoming edges (0 if none);
        for an edge j -> i:  m = lin_l(x)_j + lin_r(x)_i + lin_edge(a_ji);
        score_h = sum_c att[h,c] * leaky_relu(m[h,c], 0.2);  alpha = softmax over the edges entering i
        (exp(s - max) / (sum + 1e-16));  out_i = sum_j alpha_ji * lin_l(x)_j, heads concatenated, + bias.
        Attention dropout is inactive in eval mode.
        """

        def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                     negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                     edge_dim: Optional[int] = None, fill_value="mean", bias: bool = True,
                     share_weights: bool = False, **kwargs):
            super().__init__()
            if not concat or share_weights or edge_dim is None or not add_self_loops or fill_value != "mean":
                raise NotImplementedError("only the configuration the reference uses is restated")
            self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
            self.negative_slope, self.dropout = negative_slope, dropout
            self.lin_l = PyGLinear(in_channels, heads * out_channels, bias=True, weight_initializer="glorot")
            self.lin_r = PyGLinear(in_channels, heads * out_channels, bias=True, weight_initializer="glorot")
            self.lin_edge = PyGLinear(edge_dim, heads * out_channels, bias=False, weight_initializer="glorot")
            self.att = nn.Parameter(torch.empty(1, heads, out_channels))
            self.bias = nn.Parameter(torch.zeros(heads * out_channels))
            a = math.sqrt(6.0 / (heads + out_channels))
            nn.init.uniform_(self.att, -a, a)

        def forward(self, x, edge_index, edge_attr):
            n, H, Cc = x.size(0), self.heads, self.out_channels
            x_l = self.lin_l(x).view(n, H, Cc)
            x_r = self.lin_r(x).view(n, H, Cc)
            src, dst = edge_index[0], edge_index[1]
            keep = src != dst
            src, dst, ea = src[keep], dst[keep], edge_attr[keep]
            loop_attr = torch.zeros(n, ea.size(1), device=x.device, dtype=ea.dtype).index_add_(0, dst, ea)
            cnt = torch.zeros(n, device=x.device, dtype=ea.dtype).scatter_add_(
                0, dst, torch.ones(dst.numel(), device=x.device, dtype=ea.dtype))
            loop_attr = loop_attr / cnt.clamp(min=1).unsqueeze(1)
            loop = torch.arange(n, device=x.device, dtype=src.dtype)
            src, dst, ea = torch.cat([src, loop]), torch.cat([dst, loop]), torch.cat([ea, loop_attr])
            m = x_l[src] + x_r[dst] + self.lin_edge(ea).view(-1, H, Cc)
            m = torch.nn.functional.leaky_relu(m, self.negative_slope)
            score = (m * self.att).sum(-1)                                           # [E', H]
            peak = torch.full((n, H), float("-inf"), device=x.device, dtype=score.dtype)
            peak = peak.scatter_reduce(0, dst.unsqueeze(1).expand(-1, H), score, "amax", include_self=True)
            ex = torch.exp(score - peak[dst])
            tot = torch.zeros(n, H, device=x.device, dtype=score.dtype).index_add_(0, dst, ex)
            alpha = ex / (tot[dst] + 1e-16)
            out = torch.zeros(n, H, Cc, device=x.device, dtype=x.dtype).index_add_(0, dst, x_l[src] * alpha.unsqueeze(-1))
            return out.view(n, H * Cc) + self.bias

    class Data:
        """torch_geometric.data.Data: attribute bag with .to(); absent attrs read as None."""

        def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
            self.x, self.edge_index, self.edge_attr, self.y = x, edge_index, edge_attr, y
            for k, v in kwargs.items():
                setattr(self, k, v)

        @property
        def num_nodes(self):
            return None if self.x is None else self.x.size(0)

        def keys(self):
            return [k for k, v in self.__dict__.items() if v is not None]

        def to(self, device, *args, **kwargs):
            for k, v in list(self.__dict__.items()):
                if torch.is_tensor(v):
                    setattr(self, k, v.to(device, *args, **kwargs))
            return self

        def cuda(self):
            return self.to("cuda")

        def cpu(self):
            return self.to("cpu")

    class Batch(Data):
        """torch_geometric.data.Batch.from_data_list: concat + edge offset + batch vector."""

        @classmethod
        def from_data_list(cls, data_list):
            out = cls()
            offset, xs, eis, batch = 0, [], [], []
            for g, d in enumerate(data_list):
                n = d.x.size(0)
                xs.append(d.x)
                eis.append(d.edge_index + offset)
                batch.append(torch.full((n,), g, dtype=torch.long, device=d.x.device))
                offset += n
            out.x = torch.cat(xs, 0)
            out.edge_index = torch.cat(eis, 1)
            out.batch = torch.cat(batch, 0)
            extra = set()
            for d in data_list:
                extra.update(k for k in d.keys() if k not in ("x", "edge_index"))
            for k in sorted(extra):
                vals = [getattr(d, k, None) for d in data_list]
                if all(torch.is_tensor(v) for v in vals):
                    setattr(out, k, torch.cat(vals, 0))
            out.num_graphs = len(data_list)
            return out
