"""
Import the UNMODIFIED reference files over the oracle shims.

TEST INFRASTRUCTURE.  The files come from /root/reference (authoring container) or, where
that tree does not exist (the GPU box), from ``oracle/_ref/`` -- the verbatim, git-ignored
copy that ``oracle/make_ref.py`` writes and gpurun ships.  Used by tests/golden/make_golden*.py
(fixture generation), by the drop-in test that runs the reference's own ``segment()`` over this
repository's classes, and by the ``reference`` CPU arm of bench.py.

The package ``gcn_grabcut/__init__.py`` imports every submodule (incl. matplotlib-based
ones, absent here), so the package object is pre-seeded as an empty namespace and the
needed submodules are imported directly; no reference file is modified.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIMS = os.path.join(_HERE, "shims")


def _find_root() -> str:
    cands = [os.environ.get("GG_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "src", "gcn_grabcut", "graph_builder.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_root()
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "gcn_grabcut", "graph_builder.py"))


def load():
    """Returns a namespace with the reference modules graph_builder, model, pipeline, grabcut."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (_REPO, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    pkg_dir = os.path.join(REFERENCE_ROOT, "src", "gcn_grabcut")
    name = "gcn_grabcut"
    if name not in sys.modules or getattr(sys.modules[name], "__gg_stub__", False) is False:
        pkg = types.ModuleType(name)
        pkg.__path__ = [pkg_dir]
        pkg.__gg_stub__ = True
        sys.modules[name] = pkg
    ns = types.SimpleNamespace()
    for sub in ("graph_builder", "grabcut", "metrics", "model", "pipeline"):
        setattr(ns, sub, importlib.import_module(f"{name}.{sub}"))
    # the orphan benchmark-image module at the reference root
    spec = importlib.util.spec_from_file_location(
        "parametric_geom_dataset", os.path.join(REFERENCE_ROOT, "parametric_geom_dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ns.parametric_geom_dataset = mod
    return ns
