"""torch_geometric.nn shim -> oracle.thirdparty (model.py:48)."""
from oracle.thirdparty import GCNConv, SAGEConv, GATv2Conv  # noqa: F401
