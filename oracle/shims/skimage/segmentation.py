"""
skimage.segmentation shim (graph_builder.py:51).

``slic`` is the INPUT PRODUCER of the trimap path (graph_builder.py:177-188), not part
of it.  The stand-in returns the label map registered with ``set_next_labels`` (so the
oracle, the reference and the CUDA path all consume the SAME label map) or, if none is
registered, the deterministic SLIC-like partition of gcn_grabcut_b200.synthetic.
"""
import numpy as np
from oracle.thirdparty import find_boundaries, mark_boundaries  # noqa: F401

_NEXT_LABELS = []


def set_next_labels(labels):
    """Queue a label map to be returned by the next ``slic`` call."""
    _NEXT_LABELS.append(np.asarray(labels))


def slic(image, n_segments=100, compactness=10.0, sigma=0, start_label=1,
         channel_axis=-1, **_):
    if _NEXT_LABELS:
        lab = _NEXT_LABELS.pop(0)
        assert lab.shape == image.shape[:2], (lab.shape, image.shape)
        return lab.astype(np.int64) - int(lab.min()) + int(start_label)
    from gcn_grabcut_b200.synthetic import slic_like_labels
    h, w = image.shape[:2]
    return slic_like_labels(h, w, int(n_segments), seed=0).astype(np.int64) + int(start_label)
