"""
CPU: the C-ABI shared library loads, exports every symbol include/gcn_grabcut_b200.h declares,
the ctypes structs mirror the header field for field, and -- with no GPU -- every entry point
fails loudly instead of falling back to a CPU path.
"""
import ctypes as C
import os
import re

import pytest

from gcn_grabcut_b200 import _native as nat

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(REPO, "include", "gcn_grabcut_b200.h")).read()


def _declared_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(gg_[a-z0-9_]+)\s*\(", body)))


def _struct_fields(name):
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), body, flags=re.S)
    assert m, name
    fields = []
    for stmt in m.group(1).split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        # "const float *a, *b" / "int32_t x" / "const float* const* p" / "gg_graph_config graph"
        names = re.findall(r"[\*\s]([A-Za-z_][A-Za-z0-9_]*)\s*(?:,|$)", stmt)
        fields.extend(names)
    return fields


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(nat.LIB_PATH):
        from gcn_grabcut_b200.build import build
        build()
    return nat.lib()


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_functions()
    assert len(declared) >= 15
    assert sorted(nat.EXPORTED_SYMBOLS) == declared, set(declared) ^ set(nat.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert lib.gg_abi_version() == 4


@pytest.mark.parametrize("cname,ctype", [("gg_graph_config", nat.GraphConfig), ("gg_graph_out", nat.GraphOut),
                                         ("gg_resgcn_weights", nat.ResGCNWeights),
                                         ("gg_path_config", nat.PathConfig)])
def test_ctypes_structs_mirror_header(cname, ctype):
    assert [f[0] for f in ctype._fields_] == _struct_fields(cname)


def test_state_dict_keys_cover_the_reference_checkpoint():
    """Every key of a reference ResGCNNet checkpoint maps to a weight pointer."""
    from oracle.model_port import random_state_dict
    state = random_state_dict(64, 3)
    mapped = set(nat._KEY_TO_FIELD)
    per_layer = {k for k in state if k.startswith(("gcn_layers.", "norms."))}
    assert set(state) - mapped - per_layer == {"in_norm.norm.num_batches_tracked"}
    fields = {f[0] for f in nat.ResGCNWeights._fields_} - {"hidden", "n_layers"}
    assert set(nat._KEY_TO_FIELD.values()) | {"gcn_lin_weight", "gcn_bias", "norm_weight", "norm_bias"} == fields


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product path must raise, never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import gcn_grabcut_b200 as gg
    h = C.c_void_p()
    assert lib.gg_create(C.byref(h), 0) == nat.GG_ERR_CUDA
    assert b"no CPU fallback" in lib.gg_last_error()
    img = np.zeros((8, 8, 3), np.uint8)
    seg = np.zeros((8, 8), np.int32)
    with pytest.raises(nat.NativeError):
        gg.GraphBuilder(img, segments=seg).build()
    with pytest.raises(nat.NativeError):
        gg.refine_trimap(np.ones((1, 3), np.float32) / 3, seg, img)
    with pytest.raises(nat.NativeError):
        gg.guided_filter(np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32))
    net = gg.ResGCNNet(hidden_channels=32, n_layers=1)
    with pytest.raises(nat.NativeError):
        net.predict_probs(gg.Data(x=torch.zeros(2, 19), edge_index=torch.zeros(2, 0, dtype=torch.long)))
    with pytest.raises(nat.NativeError):
        gg.TrimapPath(net)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "gcn_grabcut_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
