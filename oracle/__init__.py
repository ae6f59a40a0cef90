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
* PINNED the same way (``tests/golden/make_golden_model.py`` executes the reference's whole
  ``renderer/latent_gs_renderer.py`` and ``utils/deform_utils.py`` on the CPU): the ARAP connectivity and energy
  (``oracle/points.py``); the GaussianModel life cycle has no oracle of its own -- the product's host logic is
  compared with the reference classes' record directly (``tests/test_model_cpu.py``).
* PARITY UNPINNED for FPS / ball query (pytorch3d) and chamfer (chamferdist): pip packages absent from the reference
  tree and this image; restated in ``oracle/points.py`` with known-answer tests.
* PARITY UNPINNED for the rasteriser interior, KNN_CUDA, simple-knn and
  fused-ssim: their CUDA sources are third-party submodules that are absent from
  ``/root/reference`` (empty directories) and the reference has
  no tests or golden images (F2).  ``oracle/raster.py`` restates the published
  3DGS tile-rasterisation algorithm; every constant it relies on is a named
  module-level constant so a later correction is a one-line change.
"""
