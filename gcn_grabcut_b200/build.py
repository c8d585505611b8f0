"""
Build libgcn_grabcut_b200.so in-tree with nvcc for sm_100a (B200).

    python -m gcn_grabcut_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU
box with the gpurun snapshot.  Graph-construction and trimap kernels are compiled with
-fmad=false (their float32 epilogues must round like numpy / OpenCV: separate multiply
and add); the network kernels keep FMA contraction.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
REPO = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libgcn_grabcut_b200.so")
OBJ_DIR = os.path.join(PKG, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "--expt-extended-lambda"]
SOURCES = {
    "api.cu": [],
    "graph_build.cu": ["-fmad=false"],
    "trimap.cu": ["-fmad=false"],
    "resgcn.cu": [],
    "gemm_tc.cu": [],
    "gcn_fused.cu": [],
    "slic.cu": [],
    "variants.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest() -> str:
    h = hashlib.sha1()
    for root in (CSRC, os.path.join(REPO, "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode())
                    h.update(f.read())
    h.update(repr((ARCH, COMMON, sorted(SOURCES.items()))).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile (if sources changed) and return the path of the shared library."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp.sha1")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"[build] {src} FAILED\n{out}\n")
        elif verbose and out:
            print(out)
    if failed:
        raise RuntimeError("nvcc failed (see above)")
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    subprocess.run(cmd, check=True)
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
