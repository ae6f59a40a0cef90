#!/usr/bin/env python
"""How many tile instances of the c3 scene could an exact, opacity-aware tile test drop?  

The published binning admits a Gaussian to every tile its 3-sigma bounding SQUARE touches.  An instance contributes to
no pixel of a tile when  max over the tile's pixels of  opacity * exp(-q(d) / 2)  < 1/255,  q the conic's quadratic form
-- those pixels are skipped one by one by the blend loop anyway, so dropping the instance changes neither pixels nor
gradients, only the integer side (R, keys, n_contrib).  The maximum over the pixel box is found exactly: q is convex, so
its minimum over the rectangle is at the centre if inside, else on an edge (1-D minimisation per edge, clamped).

CPU only (oracle preprocess, vectorised); prints the instance counts for one frame per view."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dimo_b200 import synthetic  # noqa: E402
from oracle import camera as ocam, raster as orast  # noqa: E402


def min_quadratic_over_box(a, b, c, x0, x1, y0, y1, cx, cy):
    """min over (x,y) in [x0,x1]x[y0,y1] of a dx^2 + 2 b dx dy + c dy^2, (dx,dy) = (x-cx, y-cy); all arrays."""
    inside = (cx >= x0) & (cx <= x1) & (cy >= y0) & (cy <= y1)
    best = np.full_like(a, np.inf)
    for xe in (x0, x1):                                   # vertical edges: x fixed, minimise over y
        dx = xe - cx
        dy = np.clip(-b * dx / c, y0 - cy, y1 - cy)
        best = np.minimum(best, a * dx * dx + 2 * b * dx * dy + c * dy * dy)
    for ye in (y0, y1):                                   # horizontal edges
        dy = ye - cy
        dx = np.clip(-b * dy / a, x0 - cx, x1 - cx)
        best = np.minimum(best, a * dx * dx + 2 * b * dx * dy + c * dy * dy)
    return np.where(inside, 0.0, best)


def main(n=100000, W=512, H=512, views=(0, 3)):
    sc = synthetic.make_scene(n, n_ctrl=512, n_motions=1, seed=0)
    xyz, scales = sc["_xyz"], torch.exp(sc["_scaling"])
    rot = torch.nn.functional.normalize(sc["_rotation"])
    op = torch.sigmoid(sc["_opacity"]).reshape(-1)
    for v in views:
        cam = ocam.orbit_cam(v, 8, W, H)
        pre = orast.preprocess(xyz, scales, rot, op, cam.world_view_transform, cam.full_proj_transform,
                               cam.camera_center, cam.tanfovx, cam.tanfovy, W, H, colors_precomp=torch.zeros(n, 3))
        rect = pre["rect"].numpy().astype(np.int64)
        vis = np.nonzero(pre["tiles_touched"].numpy() > 0)[0]
        R = int(pre["tiles_touched"].sum())
        kept = 0
        conic, xy, o = pre["conic"].numpy().astype(np.float64), pre["xy"].numpy().astype(np.float64), op.numpy().astype(np.float64)
        # enumerate instances tile-column by tile-column to stay vectorised
        gi, tx, ty = [], [], []
        w = rect[vis, 2] - rect[vis, 0]
        h = rect[vis, 3] - rect[vis, 1]
        for dxi in range(int(w.max())):
            for dyi in range(int(h.max())):
                m = (w > dxi) & (h > dyi)
                gi.append(vis[m]); tx.append(rect[vis[m], 0] + dxi); ty.append(rect[vis[m], 1] + dyi)
        gi, tx, ty = np.concatenate(gi), np.concatenate(tx), np.concatenate(ty)
        assert gi.shape[0] == R
        x0, y0 = tx * 16.0, ty * 16.0
        x1, y1 = np.minimum(x0 + 15, W - 1), np.minimum(y0 + 15, H - 1)      # pixel centres are integers
        q = min_quadratic_over_box(conic[gi, 0], conic[gi, 1], conic[gi, 2], x0, x1, y0, y1, xy[gi, 0], xy[gi, 1])
        alpha_max = o[gi] * np.exp(-0.5 * q)
        kept = int((alpha_max >= 1.0 / 255.0).sum())
        print(f"view {v}: visible {len(vis)} of {n}, instances R = {R} ({R / len(vis):.1f} per Gaussian), "
              f"with a contribution somewhere in their tile: {kept} ({100.0 * kept / R:.1f} %)")


if __name__ == "__main__":
    main()
