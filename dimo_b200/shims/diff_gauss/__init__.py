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
        color, depth, normal, alpha, radii = render_one(self.raster_settings, means3D, means2D, opacities, shs,
                                                        colors_precomp, scales, rotations, cov3Ds_precomp)
        if extra_attrs is None:                    # always the case on the DIMO path (latent_gs_renderer.py:1265)
            extra = torch.zeros(0, color.shape[1], color.shape[2], dtype=color.dtype, device=color.device)
            return color, depth, normal, alpha, radii, extra
        # extra per-Gaussian attributes [N, E]: alpha-blended like colours, three channels per pass, black background
        rs0 = self.raster_settings._replace(bg=torch.zeros_like(self.raster_settings.bg))
        E = extra_attrs.shape[1]
        pad = (-E) % 3
        attrs = torch.nn.functional.pad(extra_attrs, (0, pad)) if pad else extra_attrs
        planes = [render_one(rs0, means3D, None, opacities, None, attrs[:, k:k + 3].contiguous(), scales, rotations,
                             cov3Ds_precomp)[0] for k in range(0, E + pad, 3)]
        return color, depth, normal, alpha, radii, torch.cat(planes, dim=0)[:E]
