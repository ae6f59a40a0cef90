"""Drop-in for `diff_gaussian_rasterization` (ashawkey fork @ 8829d14f): imported by
renderer/latent_gs_renderer.py:9-12 and src/helpers.py:6; call kwargs (means3D, means2D, shs,
colors_precomp, opacities, scales, rotations, cov3D_precomp) -> 4-tuple (image, radii, depth, alpha)
(:1268-1277).  Same kernels as diff_gauss; the normal map is simply not returned."""
from dimo_b200.shims._raster_common import GaussianRasterizationSettings, _Base, render_one

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer"]


class GaussianRasterizer(_Base):
    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        color, depth, _normal, alpha, radii = render_one(self.raster_settings, means3D, means2D, opacities, shs,
                                                         colors_precomp, scales, rotations, cov3D_precomp)
        return color, radii, depth, alpha
