"""Shared implementation of the two reference rasteriser front-ends (B=1 of dimo_b200.raster)."""
from typing import NamedTuple

import torch
import torch.nn as nn

from dimo_b200 import raster as _raster


class GaussianRasterizationSettings(NamedTuple):
    # field order and names: renderer/latent_gs_renderer.py:1133-1146 (diff_gauss) and :1149-1162;
    # src/helpers.py:41-53 constructs it without `debug`, hence the default.
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool = False


def render_one(rs, means3D, means2D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp):
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    if cov3D_precomp is not None:
        raise NotImplementedError("dimo_b200: precomputed 3D covariance is not on the DIMO hot path "
                                  "(renderer passes scales/rotations, latent_gs_renderer.py:1185-1189)")
    cams = _raster.pack_cameras(rs.viewmatrix, rs.projmatrix, rs.campos, rs.tanfovx, rs.tanfovy, rs.bg)
    color, depth, normal, alpha, radii = _raster.rasterize_batch(
        cams, means3D, scales, rotations, opacities, rs.image_width, rs.image_height, shs=shs,
        colors_precomp=colors_precomp, sh_degree=rs.sh_degree, scale_modifier=rs.scale_modifier, means2D=means2D)
    if rs.debug:
        torch.cuda.synchronize()
    return color[0], depth[0], normal[0], alpha[0], radii[0]


class _Base(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """frustum test only (unused by DIMO): p_view.z > 0.2"""
        V = self.raster_settings.viewmatrix
        with torch.no_grad():
            z = positions @ V[:3, 2] + V[3, 2]
            return z > 0.2
