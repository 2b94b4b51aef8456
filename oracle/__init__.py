"""CPU oracle for the descriptor-match-and-solve hot path of vniclas/VFM-Registration.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vfm_registration_b200/`` imports, links or
executes anything from this package; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.

Two layers:

* ``oracle.match / oracle.ransac / oracle.project / oracle.metrics / oracle.vit`` --
  float64 NumPy (torch-CPU for the ViT) restatements that follow the reference's own
  semantics, each function citing the reference file:line it follows.
* ``oracle/c/*.c`` (built into ``oracle/_build/liboracle.so`` by ``oracle/Makefile``,
  loaded through ``oracle.cref``) -- a dependency-free C/OpenMP restatement in the
  *canonical arithmetic order* documented in DESIGN.md, so the CUDA path can be compared
  bit-for-bit (indices, similarities, inlier counts, masks, transforms).  It is also the
  timed CPU baseline.

PARITY PINNING STATUS
---------------------
* a3/a4/a5/a7/a11/a12 (projection, gather/dedup, mutual filter, metrics, transform_pcl):
  PINNED -- the NumPy restatements are checked against outputs of the reference's own
  Python functions, executed in the build container from ``/root/reference`` by
  ``oracle/gen_golden.py`` and committed under ``tests/golden/``.
* Kabsch reflection handling: PINNED against the reference's in-tree
  ``pointdsc/common.py:rigid_transform_3d`` (same golden mechanism).
* a6 (faiss ``IndexFlatIP`` + ``fvec_renorm_L2``) and a10 (Open3D 0.18
  ``registration_ransac_based_on_correspondence``): **parity unpinned** -- the arithmetic
  lives in third-party packages that are neither vendored in ``/root/reference`` nor
  installable offline, and the reference ships no golden vectors for them.  The oracle
  restates their published algorithms, anchored on the reference's call sites
  (``VoxelHashMap.cpp:461-626``, ``registration_node.py:319-327``).
* a1/a2 (DINOv2 via torch.hub FeatUp): **parity unpinned** with respect to the hub
  weights (no network); architecture-level parity against ``transformers.Dinov2Model``
  with seeded random weights.
"""
