"""torch_geometric.data shim -> oracle.thirdparty (model.py:49, pipeline.py:293)."""
from oracle.thirdparty import Data, Batch  # noqa: F401
