"""B200-native trimap path of GCN-GrabCut (label map -> region graph -> ResGCNNet -> trimap)."""
__version__ = "0.1.0"
