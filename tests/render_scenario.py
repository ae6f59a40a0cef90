"""One scripted set of `Renderer.render` calls written against the REFERENCE class surface
(renderer/latent_gs_renderer.py:973-1293) so the same function drives

  * the reference's own Renderer on the CPU, with the oracle rasteriser standing in for `diff_gauss`
    (tests/golden/make_golden_render.py, build container only)                        -> tests/golden/render.npz
  * dimo_b200.renderer.Renderer on the GPU (tests/test_zz_reference_flow_gpu.py)

and the two records are compared key by key: stage s1, stage s2 (local frame), stage s2 with local_frame=False,
override_color.  Every call also back-propagates a fixed linear functional of
(image, depth, normal, alpha) and records parameter gradients, including viewspace_points.grad."""
import numpy as np
import torch

W = H = 48
CASES = [
    ("s1", dict(stage="s1", time=0.3, latent_index=1)),
    ("s2", dict(stage="s2", time=0.6, latent_index=0)),
    ("s2_global", dict(stage="s2", time=0.45, latent_index=1, local_frame=False)),
    ("s2_override", dict(stage="s2", time=0.1, latent_index=0, override_color="colors")),
    # compute_cov3D_python=True is not listed: the reference itself fails on it in every stage (rotations stays None and
    # :1219 normalises it -> AttributeError); the shim's cov3Ds_precomp input is tested on its own (test_raster_gpu.py)
]
GRAD_PARAMS = ("_xyz", "_features_dc", "_opacity", "_scaling", "_rotation", "_c_xyz", "_c_radius", "_latent_codes")


def _np(t):
    return t.detach().cpu().numpy().copy()


def build(renderer_cls, device, init_kwargs=None):
    np.random.seed(21)
    torch.manual_seed(21)
    r = renderer_cls(sh_degree=0, white_background=True, num_latent_code=2, latent_code_dim=32, add_normal=True)
    r.initialize(num_pts=300, num_cpts=24, radius=0.5, radius2=0.5, **(init_kwargs or {}))
    g = r.gaussians
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():       # a live deformation (the reference init is the identity) and a non-trivial appearance
        g._timenet.pts_layers[-1].weight.copy_(0.02 * torch.randn(3, 256, generator=gen))
        g._timenet.rot_layers[-1].weight.copy_(0.02 * torch.randn(4, 256, generator=gen))
        g._opacity.copy_(torch.randn(g._opacity.shape, generator=gen))
        g._scaling.add_((0.9 + 0.2 * torch.randn(g._scaling.shape, generator=gen)).to(g._scaling.device))
        g._features_dc.copy_(torch.randn(g._features_dc.shape, generator=gen))
        g._rotation.copy_(torch.randn(g._rotation.shape, generator=gen))
        g._latent_codes.copy_(torch.randn(g._latent_codes.shape, generator=gen))
    return r


def run(r, make_cam, knn_fn, device):
    """make_cam(view) -> MiniCam of the class family under test; knn_fn(c_xyz, xyz) -> (dist [N,4], idx [N,4] int64)."""
    g = r.gaussians
    gen = torch.Generator().manual_seed(9)
    wi = torch.rand(3, H, W, generator=gen).to(device); wd = (0.3 * torch.rand(1, H, W, generator=gen)).to(device)
    wn = (torch.rand(3, H, W, generator=gen) - 0.5).to(device); wa = torch.rand(1, H, W, generator=gen).to(device)
    colors = torch.rand(g._xyz.shape[0], 3, generator=gen).to(device)
    dist, idx = knn_fn(g._c_xyz.detach(), g._xyz.detach())
    g.neighbor_dists, g.neighbor_indices = dist, idx
    rec = {}
    for ci, (tag, kw) in enumerate(CASES):
        kw = dict(kw)
        if kw.get("override_color") == "colors":
            kw["override_color"] = colors
        for name in GRAD_PARAMS:
            getattr(g, name).grad = None
        for p in g._timenet.parameters():
            p.grad = None
        out = r.render(make_cam(ci), **kw)
        loss = (out["image"] * wi).sum() + (out["depth"] * wd).sum() + (out["normal"] * wn).sum() + (out["alpha"] * wa).sum()
        loss.backward()
        for k in ("image", "depth", "normal", "alpha", "radii", "pts_t", "cpts_t"):
            rec[f"{tag}/{k}"] = _np(out[k])
        rec[f"{tag}/visibility_filter"] = _np(out["visibility_filter"])
        rec[f"{tag}/viewspace_grad"] = _np(out["viewspace_points"].grad)
        for name in GRAD_PARAMS:
            gr = getattr(g, name).grad
            rec[f"{tag}/grad{name}"] = _np(gr) if gr is not None else np.zeros(0, dtype=np.float32)
        rec[f"{tag}/grad_timenet_w0"] = _np(g._timenet.deformnet[0].weight.grad)
        rec[f"{tag}/grad_timenet_pts2"] = _np(g._timenet.pts_layers[2].weight.grad)
    return rec
