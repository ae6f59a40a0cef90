"""CPU oracle for the DIMO deform -> raster -> loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``dimo_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Parity status ("Oracle"):

* PINNED against the reference's own in-tree Python (run in the build container,
  fixtures committed under ``tests/golden/`` by ``tests/golden/make_golden.py``):
  positional encoding, TimeNet, LBS skinning block, quaternion helpers,
  activations, SH evaluation, camera matrices, SSIM / L1.
* PARITY UNPINNED for the rasteriser interior, KNN_CUDA, simple-knn and
  fused-ssim: their CUDA sources are third-party submodules that are absent from
  ``/root/reference`` (empty directories) and the reference has
  no tests or golden images (F2).  ``oracle/raster.py`` restates the published
  3DGS tile-rasterisation algorithm; every constant it relies on is a named
  module-level constant so a later correction is a one-line change.
"""
