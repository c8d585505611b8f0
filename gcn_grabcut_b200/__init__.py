"""
B200-native trimap path of GCN-GrabCut:
label map + BGR image -> attributed region graph -> ResGCNNet posterior -> guided-filter trimap.

The public names mirror the reference package ``gcn_grabcut`` for this path
(src/gcn_grabcut/__init__.py:10-49); everything is executed by hand-written sm_100a CUDA
kernels in libgcn_grabcut_b200.so (see include/gcn_grabcut_b200.h).  No CPU fallback.
"""
__version__ = "0.1.0"

from .graph_builder import (  # noqa: F401
    GraphBuilder, SuperpixelGraph, SuperpixelGraphConfig, BatchedRegionGraphs, build_graph_batch, slic_labels,
    compute_auto_prior, encode_user_hints, N_NODE_FEATS, N_EDGE_FEATS, N_PRIOR_FEATS, N_IMAGE_FEATS,
)

try:  # torch-dependent names (same guard as the reference's __init__)
    from .model import (  # noqa: F401
        ResGCNNet, GCNTrimapNet, GATTrimapNet, build_model, Data, _probs_to_trimap, probs_to_node_trimap, project_to_pixels,
        TRIMAP_BG, TRIMAP_FG, TRIMAP_PROB_BG, TRIMAP_PROB_FG, CLASS_BG, CLASS_UNK, CLASS_FG,
    )
    from .pipeline import (guided_filter, refine_trimap, seed_from_prior, TrimapPath, PendingTrimaps,  # noqa: F401
                           shard_range, grabcut_guards, clean_mask)  # noqa: F401
    from .dataset import derive_trimap_labels, prepare_sample, prepare_samples, prepare_dataset, cache_key  # noqa: F401
    _MODELS_AVAILABLE = True
except ImportError:  # pragma: no cover
    _MODELS_AVAILABLE = False
