"""Host glue of Renderer.render()'s non-default flags on a machine without a GPU: the rasteriser entry point is
replaced by a recorder and TimeNet's launch set by the oracle, so only the Python in dimo_b200/renderer.py and
dimo_b200/flags.py runs (their arithmetic is pinned by tests/test_flags_cpu.py; the kernels by the GPU tests)."""
import pytest
import torch

import dimo_b200.deform as dd
import dimo_b200.raster as dr
from dimo_b200 import synthetic
from dimo_b200.camera import orbit_minicam
from dimo_b200.renderer import Renderer
from oracle import deform as od
from oracle import knn as oknn


@pytest.fixture()
def recorder(monkeypatch):
    calls = {}

    def forward_batched(self, pts, times, latents):
        ps = self.flat_params()
        params = list(zip(ps[0::2], ps[1::2]))
        outs = [od.timenet_forward(params, pts, float(t), l) for t, l in zip(times, latents)]
        return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])

    def rasterize_batch(cams, means3D, scales, rotations, opacities, W, H, shs=None, colors_precomp=None, sh_degree=0,
                        scale_modifier=1.0, means2D=None, **kw):
        calls.update(shs=shs, colors=colors_precomp, rot=rotations, means=means3D, mod=scale_modifier, deg=sh_degree)
        z = lambda c: torch.zeros(1, c, H, W)
        return z(3), z(1), z(3), z(1), torch.ones(1, scales.shape[0], dtype=torch.int32)

    monkeypatch.setattr(dd.TimeNet, "forward_batched", forward_batched)
    monkeypatch.setattr(dr, "rasterize_batch", rasterize_batch)
    return calls


def _renderer():
    r = Renderer(sh_degree=2, device="cpu", num_latent_code=2)
    g = r.gaussians
    g.load_state(synthetic.make_scene(300, n_ctrl=16, n_motions=2, sh_coeffs=9, seed=1))
    g.active_sh_degree = 2
    g.neighbor_dists, g.neighbor_indices = oknn.knn(g._c_xyz.detach(), g._xyz.detach(), 4)
    return r, g


def test_flags_reach_the_rasteriser_as_the_reference_would_hand_them(recorder):
    r, g = _renderer()
    cam = orbit_minicam(1, 8, 32, 32, device="cpu")
    out = r.render(cam, time=0.3, stage="s2", latent_index=1, local_frame=False)
    assert set(out) == {"image", "depth", "normal", "alpha", "viewspace_points", "visibility_filter", "radii", "pts_t",
                        "cpts_t"}                                         # result keys, latent_gs_renderer.py:1283-1293
    assert out["pts_t"].shape == (300, 3) and recorder["shs"] is not None and recorder["colors"] is None
    assert float(recorder["rot"].detach().norm(dim=1).sub(1).abs().max()) < 1e-6
    r.render(cam, time=0.3, stage="s1", latent_index=0, convert_SHs_python=True)
    assert recorder["shs"] is None and recorder["colors"].shape == (300, 3) and float(recorder["colors"].min()) >= 0
    r.render(cam, time=0.3, stage="s1", latent_index=0, compute_cov3D_python=True, scaling_modifier=1.5)
    assert recorder["mod"] == 1.5
    assert torch.allclose(recorder["rot"], torch.nn.functional.normalize(g._rotation))      # canonical rotations
    with pytest.raises(ValueError):
        r.render(cam, time=0.3, stage="s2", compute_cov3D_python=True)
    with pytest.raises(ValueError):
        r.render(cam, time=0.3, stage="s0")
    override = torch.rand(300, 3)
    r.render(cam, time=0.0, stage="s1", override_color=override, xyz_detach=True)
    assert recorder["colors"] is override and recorder["shs"] is None and not recorder["means"].requires_grad
