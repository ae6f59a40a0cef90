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


def cov3d_to_scale_rotation(cov6):
    """Precomputed 3-D covariances [N,6] = (xx, xy, xz, yy, yz, zz) (strip_symmetric layout,
    renderer/latent_gs_renderer.py:61-66, 251-255) -> (scales [N,3], quaternions [N,4] (r,x,y,z)) with
    R diag(s^2) R^T = Sigma, by a symmetric eigen-decomposition (differentiable: gradients reach the covariances through
    torch.linalg.eigh).  The rasteriser kernels take (scale, rotation); this is the adapter for the reference's
    compute_cov3D_python path (:1184-1185), which its drivers never enable."""
    c = cov6
    S = torch.stack([torch.stack([c[:, 0], c[:, 1], c[:, 2]], -1), torch.stack([c[:, 1], c[:, 3], c[:, 4]], -1),
                     torch.stack([c[:, 2], c[:, 4], c[:, 5]], -1)], -2)
    evals, evecs = torch.linalg.eigh(S)                                  # ascending eigenvalues, orthonormal columns
    det = torch.linalg.det(evecs)
    evecs = torch.cat([evecs[..., :2], evecs[..., 2:] * det[:, None, None]], dim=-1)    # proper rotation
    scales = torch.sqrt(evals.clamp_min(0.0))
    m = evecs
    m00, m01, m02 = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    m10, m11, m12 = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    m20, m21, m22 = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    cand = torch.stack([
        torch.stack([1 + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, 1 + m00 - m11 - m22, m01 + m10, m02 + m20], -1),
        torch.stack([m02 - m20, m01 + m10, 1 - m00 + m11 - m22, m12 + m21], -1),
        torch.stack([m10 - m01, m02 + m20, m12 + m21, 1 - m00 - m11 + m22], -1)], dim=1)     # [N,4 candidates,4]
    best = torch.stack([m00 + m11 + m22, m00, m11, m22], -1).argmax(dim=-1)
    q = cand[torch.arange(c.shape[0], device=c.device), best]
    return scales, torch.nn.functional.normalize(q, dim=-1)


def render_one(rs, means3D, means2D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp):
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
    scale_modifier = rs.scale_modifier
    if cov3D_precomp is not None:
        # the reference applies scaling_modifier while building the covariance (get_covariance(scaling_modifier))
        scales, rotations = cov3d_to_scale_rotation(cov3D_precomp)
        scale_modifier = 1.0
    cams = _raster.pack_cameras(rs.viewmatrix, rs.projmatrix, rs.campos, rs.tanfovx, rs.tanfovy, rs.bg)
    color, depth, normal, alpha, radii = _raster.rasterize_batch(
        cams, means3D, scales, rotations, opacities, rs.image_width, rs.image_height, shs=shs,
        colors_precomp=colors_precomp, sh_degree=rs.sh_degree, scale_modifier=scale_modifier, means2D=means2D)
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
