"""
ctypes binding of libgcn_grabcut_b200.so (include/gcn_grabcut_b200.h).

PyTorch is used for device memory and streams only; every computation of the trimap path
happens in the CUDA library.  There is no CPU fallback: if the shared library is missing or
no sm_100 device is present, calls raise ``NativeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GG_LIB") or os.path.join(_PKG, "libgcn_grabcut_b200.so")   # GG_LIB: A/B builds

GG_OK, GG_ERR_INVALID, GG_ERR_CUDA, GG_ERR_CAPACITY, GG_ERR_STATE = 0, -1, -2, -3, -4


class NativeError(RuntimeError):
    """A call into libgcn_grabcut_b200.so failed (message from gg_last_error)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"[gg status {status}] {message}")
        self.status = status


# ----------------------------------------------------------------------------- structs
class GraphConfig(C.Structure):
    _fields_ = [("connectivity", C.c_int32), ("n_nonlocal", C.c_int32),
                ("node_cap", C.c_int32), ("pair_cap", C.c_int32)]


class GraphOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "n_nodes", "n_edges", "node_off", "edge_off", "x", "edge_index", "edge_attr",
        "centroids", "areas", "csr_rowptr", "csr_src", "csr_eid", "n_adj_pairs", "n_nl_pairs",
        "shared_cnt")]


_WEIGHT_PTRS = (
    "jk_logits", "in_norm_weight", "in_norm_bias", "in_norm_mean", "in_norm_var",
    "input_proj_0_weight", "input_proj_0_bias", "input_proj_1_weight", "input_proj_1_bias",
    "prior_booster_0_weight", "prior_booster_0_bias", "prior_booster_2_weight", "prior_booster_2_bias",
    "edge_enc_0_weight", "edge_enc_0_bias", "edge_enc_2_weight", "edge_enc_2_bias",
    "edge_gate_0_weight", "edge_gate_0_bias", "edge_gate_1_weight", "edge_gate_1_bias",
    "gcn_lin_weight", "gcn_bias", "norm_weight", "norm_bias",
    "sage_lin_l_weight", "sage_lin_l_bias", "sage_lin_r_weight", "sage_norm_weight", "sage_norm_bias",
    "ctx_attn_weight", "ctx_attn_bias", "ctx_compress_weight", "ctx_compress_bias",
    "ctx_expand_weight", "ctx_expand_bias", "fuse_0_weight", "fuse_0_bias",
    "fuse_1_weight", "fuse_1_bias", "head_weight", "head_bias")


class ResGCNWeights(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("n_layers", C.c_int32)] + [(n, C.c_void_p) for n in _WEIGHT_PTRS]


class VariantWeights(C.Structure):
    _fields_ = [("variant", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32),
                ("n_tensors", C.c_int32), ("reserved", C.c_int32), ("tensors", C.c_void_p), ("numel", C.c_void_p)]


class PathConfig(C.Structure):
    _fields_ = [("graph", GraphConfig), ("radius", C.c_int32), ("eps", C.c_float),
                ("thr_fg", C.c_float), ("thr_bg", C.c_float), ("edge_aware", C.c_int32),
                ("chunk", C.c_int32), ("label_bytes", C.c_int32), ("reserved", C.c_int32),
                ("seed_frac", C.c_double), ("slic_segments", C.c_int32), ("slic_iters", C.c_int32),
                ("slic_compactness", C.c_float), ("slic_sigma", C.c_float)]


# state-dict key -> struct field (single tensors)
_KEY_TO_FIELD = {
    "jk_logits": "jk_logits",
    "in_norm.norm.weight": "in_norm_weight", "in_norm.norm.bias": "in_norm_bias",
    "in_norm.norm.running_mean": "in_norm_mean", "in_norm.norm.running_var": "in_norm_var",
    "input_proj.0.weight": "input_proj_0_weight", "input_proj.0.bias": "input_proj_0_bias",
    "input_proj.1.weight": "input_proj_1_weight", "input_proj.1.bias": "input_proj_1_bias",
    "prior_booster.0.weight": "prior_booster_0_weight", "prior_booster.0.bias": "prior_booster_0_bias",
    "prior_booster.2.weight": "prior_booster_2_weight", "prior_booster.2.bias": "prior_booster_2_bias",
    "edge_ctx.encode.0.weight": "edge_enc_0_weight", "edge_ctx.encode.0.bias": "edge_enc_0_bias",
    "edge_ctx.encode.2.weight": "edge_enc_2_weight", "edge_ctx.encode.2.bias": "edge_enc_2_bias",
    "edge_ctx.to_gate.0.weight": "edge_gate_0_weight", "edge_ctx.to_gate.0.bias": "edge_gate_0_bias",
    "edge_ctx.to_gate.1.weight": "edge_gate_1_weight", "edge_ctx.to_gate.1.bias": "edge_gate_1_bias",
    "sage.lin_l.weight": "sage_lin_l_weight", "sage.lin_l.bias": "sage_lin_l_bias",
    "sage.lin_r.weight": "sage_lin_r_weight",
    "sage_norm.weight": "sage_norm_weight", "sage_norm.bias": "sage_norm_bias",
    "ctx.attn.weight": "ctx_attn_weight", "ctx.attn.bias": "ctx_attn_bias",
    "ctx.compress.weight": "ctx_compress_weight", "ctx.compress.bias": "ctx_compress_bias",
    "ctx.expand.weight": "ctx_expand_weight", "ctx.expand.bias": "ctx_expand_bias",
    "fuse.0.weight": "fuse_0_weight", "fuse.0.bias": "fuse_0_bias",
    "fuse.1.weight": "fuse_1_weight", "fuse.1.bias": "fuse_1_bias",
    "head.weight": "head_weight", "head.bias": "head_bias",
}

EXPORTED_SYMBOLS = (
    "gg_abi_version", "gg_last_error", "gg_create", "gg_destroy", "gg_set_option",
    "gg_check_device_status", "gg_build_graphs", "gg_pixel_planes", "gg_load_weights",
    "gg_coo_to_csr", "gg_resgcn_forward", "gg_variant_load_weights", "gg_variant_forward", "gg_refine_trimap", "gg_project_trimap",
    "gg_guided_filter", "gg_region_labels", "gg_seed_from_prior", "gg_grabcut_guards", "gg_clean_masks", "gg_auto_prior", "gg_slic", "gg_trimap_path_host", "gg_trimap_path_host_submit", "gg_trimap_path_host_wait",
    "gg_trimap_path_device", "gg_kernel_launch_count",
    "gg_profile_enable", "gg_profile_report", "gg_selftest_math")

_lib = None
_lib_lock = threading.Lock()


def lib() -> C.CDLL:
    """The loaded shared library (fails loudly when it has not been built)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeError(GG_ERR_STATE,
                                  f"{LIB_PATH} not found: build it with `python -m gcn_grabcut_b200.build` "
                                  "(nvcc, sm_100a). There is no CPU fallback.")
            L = C.CDLL(LIB_PATH)
            L.gg_last_error.restype = C.c_char_p
            L.gg_kernel_launch_count.restype = C.c_int64
            L.gg_kernel_launch_count.argtypes = [C.c_void_p]
            L.gg_destroy.restype = None
            L.gg_destroy.argtypes = [C.c_void_p]
            L.gg_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
            L.gg_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
            L.gg_check_device_status.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
            L.gg_build_graphs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(GraphConfig), C.POINTER(GraphOut), C.c_void_p]
            L.gg_pixel_planes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.gg_load_weights.argtypes = [C.c_void_p, C.POINTER(ResGCNWeights)]
            L.gg_coo_to_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
            L.gg_resgcn_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
            L.gg_variant_load_weights.argtypes = [C.c_void_p, C.POINTER(VariantWeights)]
            L.gg_variant_forward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
            L.gg_refine_trimap.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                                           C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.gg_project_trimap.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
            L.gg_guided_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                           C.c_float, C.c_void_p, C.c_void_p]
            L.gg_region_labels.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                           C.c_void_p]
            L.gg_seed_from_prior.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_int, C.c_int, C.c_int64, C.c_double, C.c_void_p]
            L.gg_slic.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            L.gg_auto_prior.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
            L.gg_grabcut_guards.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
            L.gg_clean_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                         C.c_int, C.c_void_p]
            L.gg_trimap_path_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                              C.c_int, C.POINTER(PathConfig), C.c_void_p, C.c_void_p,
                                              C.c_void_p]
            L.gg_trimap_path_host_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                     C.c_int, C.POINTER(PathConfig), C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.POINTER(C.c_int)]
            L.gg_trimap_path_host_wait.argtypes = [C.c_void_p, C.c_int]
            L.gg_trimap_path_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                                C.c_int, C.POINTER(PathConfig), C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_void_p]
            L.gg_selftest_math.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
            L.gg_profile_enable.argtypes = [C.c_void_p, C.c_int]
            L.gg_profile_report.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
            _lib = L
    return _lib


def check(status: int) -> None:
    if status != GG_OK:
        raise NativeError(status, lib().gg_last_error().decode("utf-8", "replace"))


# ----------------------------------------------------------------------------- handles
class Handle:
    """One gg_context per CUDA device."""

    def __init__(self, device: int):
        self.device = int(device)
        self._h = C.c_void_p()
        check(lib().gg_create(C.byref(self._h), self.device))
        self.weights_token = None          # identity of the state-dict currently loaded

    @property
    def ptr(self) -> C.c_void_p:
        return self._h

    def launches(self) -> int:
        return int(lib().gg_kernel_launch_count(self._h))

    def set_option(self, key: str, value: int) -> None:
        check(lib().gg_set_option(self._h, key.encode(), int(value)))

    def check_status(self, stream: int = 0) -> None:
        bits = C.c_int(0)
        check(lib().gg_check_device_status(self._h, C.c_void_p(stream), C.byref(bits)))

    def profile(self, enable: bool) -> None:
        check(lib().gg_profile_enable(self._h, int(bool(enable))))

    def profile_report(self):
        """[(kernel, launches, total_ms)] sorted by total time, from the CUDA-event session."""
        buf = C.create_string_buffer(1 << 16)
        check(lib().gg_profile_report(self._h, buf, len(buf)))
        rows = []
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(",", 2)
            rows.append((name, int(n), float(ms)))
        return rows

    def close(self):
        if self._h:
            lib().gg_destroy(self._h)
            self._h = C.c_void_p()


_handles: Dict[int, Handle] = {}


def handle(device: Optional[int] = None) -> Handle:
    import torch
    if not torch.cuda.is_available():
        raise NativeError(GG_ERR_CUDA, "no CUDA device visible; gcn_grabcut_b200 has no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    device = int(device)
    if device not in _handles:
        _handles[device] = Handle(device)
    return _handles[device]


def device_index(device) -> int:
    import torch
    d = torch.device(device) if not isinstance(device, torch.device) else device
    if d.type != "cuda":
        raise NativeError(GG_ERR_CUDA, f"device {d}: the trimap path runs on CUDA only (no CPU fallback)")
    if not torch.cuda.is_available():
        raise NativeError(GG_ERR_CUDA, "no CUDA device visible; gcn_grabcut_b200 has no CPU fallback")
    return torch.cuda.current_device() if d.index is None else d.index


def ptr(t, dtype=None) -> C.c_void_p:
    """Device (or pinned host) pointer of a contiguous torch tensor; None -> NULL."""
    if t is None:
        return C.c_void_p(0)
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return C.c_void_p(t.data_ptr())


def current_stream(device: int) -> int:
    import torch
    return int(torch.cuda.current_stream(device).cuda_stream)


# ----------------------------------------------------------------------------- weights
def load_state_dict(h: Handle, state: dict) -> None:
    """gg_load_weights from a reference-layout state-dict (torch tensors or numpy arrays)."""
    def arr(v):
        if hasattr(v, "detach"):
            v = v.detach().to("cpu")
            v = v.float().numpy()
        return np.ascontiguousarray(np.asarray(v, dtype=np.float32))

    missing = [k for k in _KEY_TO_FIELD if k not in state]
    if missing:
        raise KeyError(f"state_dict is missing keys: {missing[:6]}{'…' if len(missing) > 6 else ''}")
    D = int(state["input_proj.0.weight"].shape[0])
    n = sum(1 for k in state if k.startswith("gcn_layers.") and k.endswith(".bias"))
    keep = []                                   # keep the numpy arrays alive during the call
    w = ResGCNWeights()
    w.hidden, w.n_layers = D, n
    for key, fld in _KEY_TO_FIELD.items():
        a = arr(state[key])
        keep.append(a)
        setattr(w, fld, a.ctypes.data_as(C.c_void_p))
    for fld, pat in (("gcn_lin_weight", "gcn_layers.{}.lin.weight"), ("gcn_bias", "gcn_layers.{}.bias"),
                     ("norm_weight", "norms.{}.weight"), ("norm_bias", "norms.{}.bias")):
        arrs = [arr(state[pat.format(i)]) for i in range(n)]
        keep.extend(arrs)
        tbl = (C.c_void_p * max(n, 1))(*[a.ctypes.data_as(C.c_void_p) for a in arrs])
        keep.append(tbl)
        setattr(w, fld, C.cast(tbl, C.c_void_p))
    check(lib().gg_load_weights(h.ptr, C.byref(w)))
    del keep


def load_variant_state_dict(h: Handle, variant: int, hidden: int, n_layers: int, n_heads: int, state: dict,
                            keys) -> None:
    """gg_variant_load_weights from a reference-layout state-dict; ``keys`` lists the tensors in the
    order include/gcn_grabcut_b200.h documents for the variant."""
    missing = [k for k in keys if k not in state]
    if missing:
        raise KeyError(f"state_dict is missing keys: {missing[:6]}{'…' if len(missing) > 6 else ''}")
    arrs = []
    for k in keys:
        v = state[k]
        if hasattr(v, "detach"):
            v = v.detach().to("cpu").float().numpy()
        arrs.append(np.ascontiguousarray(np.asarray(v, dtype=np.float32)))
    tbl = (C.c_void_p * len(arrs))(*[a.ctypes.data_as(C.c_void_p) for a in arrs])
    numel = (C.c_int64 * len(arrs))(*[int(a.size) for a in arrs])
    w = VariantWeights(int(variant), int(hidden), int(n_layers), int(n_heads), len(arrs), 0,
                       C.cast(tbl, C.c_void_p), C.cast(numel, C.c_void_p))
    check(lib().gg_variant_load_weights(h.ptr, C.byref(w)))
    del arrs
