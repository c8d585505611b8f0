"""skimage.color shim -> oracle.thirdparty (graph_builder.py:52)."""
from oracle.thirdparty import rgb2lab, rgb2hsv, rgb2xyz, xyz2lab  # noqa: F401
