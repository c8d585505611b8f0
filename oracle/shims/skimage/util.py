"""skimage.util shim -> oracle.thirdparty (parametric_geom_dataset.py:5)."""
from oracle.thirdparty import img_as_float  # noqa: F401
