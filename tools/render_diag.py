import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gpu_parity as gp, render_scenario
from dimo_b200 import knn as dknn
from dimo_b200.camera import MiniCam, orbit_camera
from dimo_b200.renderer import Renderer
gold = np.load(os.path.join(ROOT, "tests", "golden", "render.npz"))
fovy = np.deg2rad(33.9)
mk = lambda v: MiniCam(orbit_camera(-10.0 + 7 * v, 40.0 * v, 2.0), render_scenario.W, render_scenario.H, fovy, fovy, 0.01, 100)
r = render_scenario.build(Renderer, "cuda")
rec = render_scenario.run(r, mk, lambda c, x: dknn.knn(c, x, 4), "cuda")
print(sorted(set(rec) ^ set(gold.files)))
for k in gold.files:
    w, g = gold[k], rec[k]
    if w.shape != g.shape: print("SHAPE", k, w.shape, g.shape); continue
    if w.dtype.kind in "biu": print(f"{k:28s} int mismatches {(w != g).sum()} {w.dtype} {g.dtype}")
    elif w.size: print(f"{k:28s} rel {gp.rel_err(torch.from_numpy(g), torch.from_numpy(w)):.2e} out(1e-4) {gp.outlier_frac(torch.from_numpy(g), torch.from_numpy(w), 1e-4):.2e} scale {np.abs(w).max():.3g}")
    else: print(k, "empty", g.shape)
