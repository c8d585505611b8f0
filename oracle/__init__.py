"""
oracle/ -- CPU restatement of the reference trimap path.  TEST INFRASTRUCTURE ONLY.

Nothing under this directory is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker (or as the CPU baseline
being timed), never as the thing shipped.  The product package
``gcn_grabcut_b200`` never imports ``oracle`` and has no CPU fallback.

Contents
--------
thirdparty.py   restatements of the third-party functions the reference calls but
                that are absent from /root/reference and from this image
                (scikit-image ``rgb2lab`` / ``rgb2hsv`` / ``find_boundaries``,
                PyTorch-Geometric ``GCNConv`` / ``SAGEConv`` / ``Data`` / ``Batch``;
                versions are unpinned by the reference: pyproject.toml:10-23).
shims/          importable ``skimage`` / ``torch_geometric`` packages built from
                thirdparty.py so that the UNMODIFIED reference files can be
                imported in the authoring container (ref_loader.py).
graph_port.py   numpy/cv2 restatement of graph_builder.py (file:line cited per function)
model_port.py   torch restatement of ResGCNNet + helpers (model.py)
trimap_port.py  numpy/cv2 restatement of guided_filter / refine_trimap (pipeline.py)
synthetic.py    seeded restatement of parametric_geom_dataset.GeometricDataset.sample
                and the deterministic SLIC-like label-map generator used as the
                common input of oracle and CUDA path.

Pinning
-------
The reference ships no golden vectors for this path (tests/test.py holds shape /
range assertions only).  The port is pinned by running the reference's own files
(imported from /root/reference over the shims) on seeded inputs and committing the
outputs as fixtures under tests/golden/ (tests/golden/make_golden.py).  Parity with
*real* scikit-image / PyG builds is UNPINNED: neither is installable here (no
network); the shim formulas follow their published algorithms.
"""
