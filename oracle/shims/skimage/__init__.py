"""
Importable stand-in for scikit-image, built from oracle/thirdparty.py.
TEST INFRASTRUCTURE: exists only so the unmodified reference files under
/root/reference can be imported in the authoring container (scikit-image is not
installed and cannot be installed: no network).  Never on the product path.
"""
__version__ = "0.0-oracle-shim"
from . import color, segmentation, util  # noqa: F401
