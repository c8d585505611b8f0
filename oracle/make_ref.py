"""
Recipe for ``oracle/_ref/``: a verbatim, git-ignored copy of the reference's own Python files
for the trimap path, so that the UNMODIFIED reference can run where /root/reference does not
exist (the GPU box: gpurun ships the working tree, ``oracle/_ref/`` included, but not
/root/reference).

    python -m oracle.make_ref            # copy /root/reference/... -> oracle/_ref/...

TEST INFRASTRUCTURE.  Nothing is modified; the files keep their relative paths
(``src/gcn_grabcut/*.py``, ``parametric_geom_dataset.py``), so ``oracle/ref_loader.py`` loads
them exactly as it loads /root/reference.  ``oracle/_ref/`` is listed in .gitignore (reference
sources never enter this repository's history) and NOT in .gpurunignore.

Used by
  * tests/test_gpu_parity.py::test_reference_segment_through_the_drop_in  (the reference's own
    GCNGrabCutPipeline.segment with the names of INTEGRATION.md §2 swapped in);
  * bench.py --impl reference / cpu_baseline (kind "reference": the reference's GraphBuilder,
    ResGCNNet.predict_probs and refine_trimap, timed on the host cores).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("GG_REFERENCE_SRC", "/root/reference")

FILES = [
    "src/gcn_grabcut/graph_builder.py",
    "src/gcn_grabcut/grabcut.py",
    "src/gcn_grabcut/metrics.py",
    "src/gcn_grabcut/model.py",
    "src/gcn_grabcut/pipeline.py",
    "src/gcn_grabcut/dataset.py",
    "src/gcn_grabcut/losses.py",
    "parametric_geom_dataset.py",
]


def make(verbose: bool = True) -> bool:
    """Copy the files; returns False (and does nothing) when the reference tree is absent."""
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        if verbose:
            print(f"[make_ref] {SRC} not present: nothing copied (using what oracle/_ref already holds)")
        return False
    lines = []
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            lines.append(f"{hashlib.sha1(f.read()).hexdigest()}  {rel}")
    with open(os.path.join(DEST, "MANIFEST.sha1"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print(f"[make_ref] copied {len(FILES)} reference files to {DEST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
