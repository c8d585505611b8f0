"""
Importable stand-in for PyTorch Geometric, built from oracle/thirdparty.py.
TEST INFRASTRUCTURE: lets the unmodified reference model.py / pipeline.py import in the
authoring container (torch_geometric is not installed; no network).  Never on the
product path.
"""
__version__ = "0.0-oracle-shim"
from . import nn, data  # noqa: F401
