"""Shared helpers for the test-suite: golden fixtures and seeded inputs."""
from __future__ import annotations

import glob
import hashlib
import os

import numpy as np

from gcn_grabcut_b200.synthetic import geometric_sample, slic_like_labels

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_inputs(g, name):
    """Rebuild (image, label map) of a golden case from its seeds and check the digests."""
    H, W, seed = int(g["H"]), int(g["W"]), int(g["seed"])
    if name.startswith("geom"):
        img = geometric_sample(H, W, seed)[0]
    else:
        img = np.random.RandomState(seed).randint(20, 220, (H, W, 3), dtype=np.uint8)
    seg = slic_like_labels(H, W, int(g["n_segments"]), seed)
    assert digest(img) == str(g["image_sha1"]), "input image drifted from the fixture"
    assert digest(seg) == str(g["seg_sha1"]), "label map drifted from the fixture"
    return img, seg
