#!/usr/bin/env python
"""Generates tests/golden/render.npz by EXECUTING THE REFERENCE'S OWN `Renderer.render`
(renderer/latent_gs_renderer.py:1096-1293, incl. TimeNet, the LBS block, get_covariance, MiniCam) on the CPU, with the
oracle rasteriser (oracle/raster.py, autograd backward) standing in for the absent `diff_gauss` extension.  Build
container only (needs /root/reference):  python tests/golden/make_golden_render.py
The scenario is tests/render_scenario.py, shared with the GPU test."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from make_golden_model import reference_modules  # noqa: E402
from oracle import knn as oknn, raster as oraster  # noqa: E402
import render_scenario  # noqa: E402


def oracle_diff_gauss():
    """a `diff_gauss` module whose GaussianRasterizer is the CPU oracle (same 12-field settings, same 6-tuple)"""
    from dimo_b200.shims._raster_common import GaussianRasterizationSettings, cov3d_to_scale_rotation

    class GaussianRasterizer(torch.nn.Module):
        def __init__(self, raster_settings):
            super().__init__()
            self.rs = raster_settings

        def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                    cov3Ds_precomp=None, extra_attrs=None):
            rs = self.rs
            mod = rs.scale_modifier
            if cov3Ds_precomp is not None:
                scales, rotations = cov3d_to_scale_rotation(cov3Ds_precomp)
                mod = 1.0
            o = oraster.rasterize(means3D, scales, rotations, opacities, rs.viewmatrix, rs.projmatrix, rs.campos,
                                  rs.tanfovx, rs.tanfovy, rs.image_width, rs.image_height, rs.bg, scale_modifier=mod,
                                  shs=shs, sh_degree=rs.sh_degree, colors_precomp=colors_precomp, means2D=means2D)
            extra = torch.zeros(0, rs.image_height, rs.image_width)
            return o["image"], o["depth"], o["normal"], o["alpha"], o["radii"], extra

    m = types.ModuleType("diff_gauss")
    m.GaussianRasterizationSettings = GaussianRasterizationSettings
    m.GaussianRasterizer = GaussianRasterizer
    return m


def main():
    sys.modules["diff_gauss"] = oracle_diff_gauss()
    renderer, _ = reference_modules()
    # the module was exec'd with the shim's diff_gauss names bound at import time: rebind to the oracle stand-in
    renderer.GaussianRasterizationSettingsNormal = sys.modules["diff_gauss"].GaussianRasterizationSettings
    renderer.GaussianRasterizerNormal = sys.modules["diff_gauss"].GaussianRasterizer
    from utils.cam_utils import orbit_camera
    fovy = np.deg2rad(33.9)

    def make_cam(view):
        return renderer.MiniCam(orbit_camera(-10.0 + 7 * view, 40.0 * view, 2.0), render_scenario.W, render_scenario.H,
                                fovy, fovy, 0.01, 100)

    r = render_scenario.build(renderer.Renderer, "cpu")
    rec = render_scenario.run(r, make_cam, lambda c, x: oknn.knn(c, x, 4), "cpu")     # (the reference reaches distCUDA2
    # through its simple_knn import, which reference_modules() serves with the oracle)
    np.savez_compressed(os.path.join(HERE, "render.npz"), **rec)
    print("render.npz:", len(rec), "arrays", {k: float(np.abs(v).max()) for k, v in rec.items() if k.endswith("/image")})


if __name__ == "__main__":
    main()
