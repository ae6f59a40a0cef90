"""Drop-in for `diff_gauss` (slothfulxtx/diff-gaussian-rasterization @ 726449a8), the live rasteriser
of the reference (renderer/latent_gs_renderer.py:13-16, 1133-1147, 1256-1266): 12-field settings,
call kwargs (means3D, means2D, shs, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
extra_attrs) -> 6-tuple (image, depth, normal, alpha, radii, extra)."""
import torch

from dimo_b200.shims._raster_common import GaussianRasterizationSettings, _Base, render_one

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]


class GaussianRasterizer(_Base):
    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3Ds_precomp=None, extra_attrs=None):
        if extra_attrs is not None:
            raise NotImplementedError("dimo_b200: extra_attrs is always None on the DIMO path "
                                      "(latent_gs_renderer.py:1265)")
        color, depth, normal, alpha, radii = render_one(self.raster_settings, means3D, means2D, opacities, shs,
                                                        colors_precomp, scales, rotations, cov3Ds_precomp)
        extra = torch.zeros(0, color.shape[1], color.shape[2], dtype=color.dtype, device=color.device)
        return color, depth, normal, alpha, radii, extra
