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
GRAD_PARAMS = ("_xyz", "_features_dc", "_opacity", "_scaling", "_rotation", "_c_xyz", "_c_radius", "_latent_codes", "_r")


def _np(t):
    return t.detach().cpu().numpy().copy()


def _load_timenet(g):
    # TimeNet weights from the seeded generator of oracle.deform.timenet_init (the two class families consume the global
    # RNG differently while initialising the MLP); heads scaled so that the deformation is live
    from oracle import deform as od
    names = [f"deformnet.{i}" for i in range(8)] + ["pts_layers.0", "pts_layers.2", "rot_layers.0", "rot_layers.2"]
    params = od.timenet_init(32, seed=7, final_scale=0.05)
    dev0 = g._xyz.device
    g._timenet.load_state_dict({f"{n}.{k}": v.to(dev0) for n, (Wt, b) in zip(names, params)
                                for k, v in (("weight", Wt), ("bias", b))})


def _appearance(g, gen, scaling=True):
    dev0 = g._xyz.device
    rnd = lambda t, s=1.0: (s * torch.randn(t.shape, generator=gen)).to(dev0)
    with torch.no_grad():
        g._opacity.copy_(rnd(g._opacity))
        g._features_dc.copy_(rnd(g._features_dc))
        g._rotation.copy_(rnd(g._rotation))
        if scaling:
            g._scaling.add_(-1.0 + rnd(g._scaling, 0.2))


def build(renderer_cls, device, init_kwargs=None):
    """Stage-s1 state of the reference flow: the Gaussians ARE the key points (GUI.__init__ initialises num_pts =
    num_cpts points, main_train_dimo.py:137-146) and share one learnable radius `_r` (get_scaling, :340-350)."""
    np.random.seed(21)
    torch.manual_seed(21)
    r = renderer_cls(sh_degree=0, white_background=True, num_latent_code=2, latent_code_dim=32, add_normal=True)
    r.initialize(num_pts=24, num_cpts=24, radius=0.5, radius2=0.5, **(init_kwargs or {}))
    g = r.gaussians
    _load_timenet(g)
    gen = torch.Generator().manual_seed(5)
    _appearance(g, gen, scaling=False)
    with torch.no_grad():
        g._r.add_(0.3)                                                   # visible key-point splats
        g._latent_codes.copy_(torch.randn(g._latent_codes.shape, generator=gen).to(g._latent_codes.device))
    return r


def to_stage_s2(r, init_kwargs=None):
    """GUI.prepare_train_s2 (main_train_dimo.py:471-500) with init_type "ag": key points -> control points, 12 Gaussians
    spawned around each, the shared radius retired."""
    g = r.gaussians
    with torch.no_grad():
        g._c_xyz.copy_(g._xyz)
        g._scaling.copy_(g._r.expand_as(g._xyz))
        g._c_radius.copy_(g._r.expand_as(g._c_radius))
    r.initialize_ag(g._c_xyz, g.get_c_radius(stage="s2"), num_cpts=g._c_xyz.shape[0], num_pts_per_cpt=12, init_ratio=1.0,
                    **(init_kwargs or {}))
    g._r = torch.tensor([], device=g._xyz.device)
    _appearance(g, torch.Generator().manual_seed(6))


def run(r, make_cam, knn_fn, device, init_kwargs=None):
    """make_cam(view) -> MiniCam of the class family under test; knn_fn(c_xyz, xyz) -> (dist [N,4], idx [N,4] int64)."""
    g = r.gaussians
    gen = torch.Generator().manual_seed(9)
    wi = torch.rand(3, H, W, generator=gen).to(device); wd = (0.3 * torch.rand(1, H, W, generator=gen)).to(device)
    wn = (torch.rand(3, H, W, generator=gen) - 0.5).to(device); wa = torch.rand(1, H, W, generator=gen).to(device)
    rec = {}
    for ci, (tag, kw) in enumerate(CASES):
        kw = dict(kw)
        if kw["stage"] == "s2" and len(g._r) > 0:                      # first s2 case: the s1 -> s2 transition
            to_stage_s2(r, init_kwargs)
            dist, idx = knn_fn(g._c_xyz.detach(), g._xyz.detach())      # GUI.find_knn (main_train_dimo.py:502-509)
            g.neighbor_dists, g.neighbor_indices = dist, idx
        if kw.get("override_color") == "colors":
            kw["override_color"] = torch.rand(g._xyz.shape[0], 3, generator=gen).to(device)
        for name in GRAD_PARAMS:
            if isinstance(getattr(g, name), torch.nn.Parameter):
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
            gr = getattr(getattr(g, name), "grad", None)
            rec[f"{tag}/grad{name}"] = _np(gr) if gr is not None else np.zeros(0, dtype=np.float32)
        rec[f"{tag}/grad_timenet_w0"] = _np(g._timenet.deformnet[0].weight.grad)
        rec[f"{tag}/grad_timenet_pts2"] = _np(g._timenet.pts_layers[2].weight.grad)
    return rec
