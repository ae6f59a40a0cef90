"""GPU diagnostic: per-tensor error table of the full step (max-norm, L2, entry outliers) with and without the
forced activation pattern, and the run-to-run spread of the CUDA gradients.  gpurun -- python tools/parity_diag.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_parity as gp

for reg in (False, True):
    for force in (True, False):
        lc, lo, gc, go, st = gp.run_step_pair(regularisers=reg, force_masks=force)
        print(f"== regularisers={reg} force_masks={force} loss {lc:.6f} vs {lo:.6f} rel {abs(lc-lo)/abs(lo):.2e} {st}")
        for k in gc:
            print(f"   {k:12s} max {gp.rel_err(gc[k], go[k]):.2e}  l2 {gp.l2_err(gc[k], go[k]):.2e}  "
                  f"out {gp.entry_outlier_frac(gc[k], go[k]):.2e}  scale {go[k].abs().max():.2e}")
a = gp.run_step_pair()[2]
b = gp.run_step_pair()[2]
print("== run-to-run (atomics order)")
for k in a:
    print(f"   {k:12s} max {gp.rel_err(a[k], b[k]):.2e}")

import time
for (N, W, H, bwd, tag) in ((100000, 512, 512, True, "c3"), (500000, 800, 800, False, "c5 fwd")):
    t0 = time.time()
    o, c = gp.run_raster_pair(N, W, H, view=3, nviews=8, use_dn=False, backward=bwd)
    print(f"== raster pair {tag}: {time.time()-t0:.1f} s, R={c['R']}")
    ints, flo, gr = gp.compare_raster(o, c)
    for k in ("image", "depth", "normal", "alpha"):
        print(f"   {k:10s} outlier_frac(1e-4) {gp.outlier_frac(c[k], o[k], 1e-4):.2e} l2 {gp.l2_err(c[k], o[k]):.2e}")
    for k in gr:
        print(f"   grad {k:10s} max {gr[k]:.2e} l2 {gp.l2_err(c['grads'][k], o['grads'][k]):.2e} "
              f"out(1e-4*scale) {gp.outlier_frac(c['grads'][k], o['grads'][k], 1e-4):.2e}")
