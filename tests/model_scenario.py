"""One scripted life of a GaussianModel on the CPU -- initialise, set up training, a few Adam steps, densify / prune,
opacity reset, schedules, the FPS-style index prune, adaptive re-initialisation -- written against the REFERENCE class
surface (renderer/latent_gs_renderer.py) so the same function drives

  * the reference's own classes (tests/golden/make_golden_model.py, build container only)  -> tests/golden/model.npz
  * dimo_b200.renderer.Renderer / GaussianModel with optimizer="torch"                      -> tests/test_model_cpu.py

and the two records are compared array by array.  Everything random comes from seeded generators.
"""
import types

import numpy as np
import torch

# configs/train_config.yaml:70-106 (the values the schedules and the optimizer groups read)
TRAIN_ARGS = dict(
    percent_dense=0.01, position_lr_init=0.01, position_lr_final=0.0002, position_lr_delay_mult=0.02,
    position_lr_max_steps=1000, feature_lr=0.01, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.005,
    c_radius_lr=0.005, latent_code_lr_init=0.005, latent_code_lr_final=0.0002, latent_code_lr_delay_mult=0.02,
    latent_code_lr_max_steps=1000, deform_lr_init=0.0002, deform_lr_final=0.000002, deform_learn_start=0,
    deformation_lr_delay_mult=0.01, c_position_lr_init=0.000002, c_position_lr_final=0.000002,
    c_position_lr_delay_mult=0.02, r_lr=0.01)


def train_args(**over):
    d = dict(TRAIN_ARGS)
    d.update(over)
    return types.SimpleNamespace(**d)


_ROWS = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_scaling", "_rotation")
_GROUP_OF = {"_xyz": "xyz", "_features_dc": "f_dc", "_features_rest": "f_rest", "_opacity": "opacity",
             "_scaling": "scaling", "_rotation": "rotation", "_c_xyz": "c_xyz", "_c_radius": "c_radius", "_r": "r"}


def _np(t):
    return t.detach().cpu().numpy().copy()


def snapshot(g, rec, tag, moments=True):
    for name in _ROWS + ("_c_xyz", "_c_radius", "_r"):
        t = getattr(g, name)
        rec[f"{tag}/{name}"] = _np(t)
        if moments and g.optimizer is not None and isinstance(t, torch.nn.Parameter) and t.numel() > 0:
            st = g.optimizer.state.get(t, None)
            if st:
                rec[f"{tag}/{name}/exp_avg"] = _np(st["exp_avg"])
                rec[f"{tag}/{name}/exp_avg_sq"] = _np(st["exp_avg_sq"])
    for name in ("max_radii2D", "xyz_gradient_accum", "denom"):
        rec[f"{tag}/{name}"] = _np(getattr(g, name))
    if g.optimizer is not None:
        rec[f"{tag}/lrs"] = np.array([grp["lr"] for grp in g.optimizer.param_groups], dtype=np.float64)
        rec[f"{tag}/group_sizes"] = np.array([sum(p.numel() for p in grp["params"]) for grp in g.optimizer.param_groups])


def fake_backward(g, gen, scale=1.0):
    """Deterministic stand-in for loss.backward(): every parameter of every group gets a seeded gradient."""
    for grp in g.optimizer.param_groups:
        for p in grp["params"]:
            p.grad = scale * torch.randn(p.shape, generator=gen, dtype=p.dtype)


def run(renderer_cls, init_kwargs=None, setup_kwargs=None, schedule_fn=None):
    """renderer_cls(sh_degree=..., num_latent_code=..., latent_code_dim=...) -> object with .gaussians.
    init_kwargs / setup_kwargs: extra keyword arguments for initialize()/initialize_ag() and training_setup() (the
    dimo_b200 classes take the 3-NN distance function and the optimizer kind there; the reference gets the former
    through its simple_knn import and has only one optimizer)."""
    init_kwargs = init_kwargs or {}
    setup_kwargs = setup_kwargs or {}
    rec = {}
    np.random.seed(11)
    torch.manual_seed(11)
    r = renderer_cls(sh_degree=0, white_background=True, num_latent_code=3, latent_code_dim=32)
    g = r.gaussians

    # ---- initialise: 400 Gaussians, 24 control points (Renderer.initialize, :995-1036) ----
    r.initialize(num_pts=400, num_cpts=24, radius=0.5, radius2=0.5, **init_kwargs)
    snapshot(g, rec, "init", moments=False)
    rec["init/scaling_act"] = _np(g.get_scaling)
    rec["init/c_radius_s1"] = _np(g.get_c_radius("s1"))

    # The reference's `_c_radius` is a Parameter over a VIEW of the log-scale tensor (latent_gs_renderer.py:445-448),
    # i.e. it aliases `_scaling[:, 0]`.  Unobservable in a DIMO run (c_radius has lr 0 in s1 and is overwritten at the
    # start of s2, main_train_dimo.py:466-470, 481); broken up here so both implementations start from equal, independent
    # tensors.
    g._c_radius = torch.nn.Parameter(g._c_radius.detach().clone())

    # ---- training_setup + schedules (:453-515) ----
    opt = train_args()
    g.training_setup(opt, **setup_kwargs)
    g.active_sh_degree = g.max_sh_degree
    for grp in g.optimizer.param_groups:              # prepare_train_s1, main_train_dimo.py:466-470
        if grp["name"] in ("c_radius", "c_xyz"):
            grp["lr"] = 0.0
    rec["setup/names"] = np.array([grp["name"] for grp in g.optimizer.param_groups])
    snapshot(g, rec, "setup")
    steps = [0, 1, 10, 250, 999, 1000, 5000]
    for name in ("xyz_scheduler_args", "c_xyz_scheduler_args", "latent_code_scheduler_args", "deform_scheduler_args"):
        rec[f"sched/{name}"] = np.array([getattr(g, name)(s) for s in steps], dtype=np.float64)
    for stage in ("s1", "s2"):
        g.update_learning_rate(321, stage)
        rec[f"sched/lrs_{stage}"] = np.array([grp["lr"] for grp in g.optimizer.param_groups], dtype=np.float64)

    # ---- three optimizer steps with seeded gradients ----
    gen = torch.Generator().manual_seed(5)
    for it in range(3):
        fake_backward(g, gen, scale=0.01)
        g.optimizer.step()
        g.optimizer.zero_grad()
    snapshot(g, rec, "steps")

    # ---- densification statistics + densify_and_prune (:826-924) ----
    n = g._xyz.shape[0]
    with torch.no_grad():
        # spread the scales so both the clone (small) and the split (large) branch fire
        # the shared radius _r is alive in stage s1 (get_scaling reads it), so the spread goes there
        g._r.data.fill_(float(np.log(0.05)))
    for it in range(4):
        vs = types.SimpleNamespace(grad=0.02 * torch.randn(n, 3, generator=gen))
        vis = torch.rand(n, generator=gen) > 0.3
        radii = torch.randint(0, 4, (n,), generator=gen).float()
        g.max_radii2D[vis] = torch.max(g.max_radii2D[vis], radii[vis])
        g.add_densification_stats(vs, vis)
    snapshot(g, rec, "stats")
    torch.manual_seed(23)
    g.densify_and_prune(0.01, min_opacity=0.01, extent=4, max_screen_size=2)
    snapshot(g, rec, "densified")

    # ---- stage-s2-like state: per-Gaussian scales (shared radius dropped), then clone AND split both fire ----
    g2 = _second_stage(renderer_cls, init_kwargs, setup_kwargs, gen, rec)

    # ---- plain prune (:892-901) and the index-tensor prune GUI.FPS performs (main_train_dimo.py:511-515) ----
    fake_backward(g2, gen, scale=0.01)
    g2.optimizer.step()
    g2.optimizer.zero_grad()
    with torch.no_grad():
        g2._opacity.data[::7] = -6.0
    g2.prune(min_opacity=0.01, extent=4, max_screen_size=None)
    snapshot(g2, rec, "pruned")
    idx = torch.randperm(g2._xyz.shape[0], generator=gen)[:40]
    g2.prune_points(idx)
    snapshot(g2, rec, "fps_pruned")

    # ---- opacity reset (:571-574) after another step ----
    fake_backward(g2, gen, scale=0.01)
    g2.optimizer.step()
    g2.optimizer.zero_grad()
    g2.reset_opacity()
    snapshot(g2, rec, "reset")

    # ---- geometry helpers (:385-410) ----
    cam = types.SimpleNamespace(camera_center=torch.tensor([0.3, -0.2, 2.0]))
    with torch.no_grad():
        g2._rotation.data = torch.randn(g2._rotation.shape, generator=gen)
        g2._scaling.data = g2._scaling.data + 0.3 * torch.randn(g2._scaling.shape, generator=gen)
    rec["geom/covariance"] = _np(g2.get_covariance(1.3))
    rec["geom/smallest_axis"] = _np(g2.get_smallest_axis())
    rec["geom/normal"] = _np(g2.get_normal(cam))
    rec["geom/rotmat"] = _np(g2.get_rotation_matrix())

    # ---- end of stage s1 (main_train_dimo.py:199-201): key points pruned together with their Gaussians ----
    np.random.seed(13)
    torch.manual_seed(13)
    r3 = renderer_cls(sh_degree=0, white_background=True, num_latent_code=2, latent_code_dim=32)
    g3 = r3.gaussians
    r3.initialize(num_pts=48, num_cpts=48, radius=0.5, radius2=0.5, **init_kwargs)
    g3._c_radius = torch.nn.Parameter(g3._c_radius.detach().clone())
    g3.training_setup(train_args(), **setup_kwargs)
    for it in range(2):
        fake_backward(g3, gen, scale=0.01)
        g3.optimizer.step()
        g3.optimizer.zero_grad()
    with torch.no_grad():
        g3._opacity.data[::4] = -7.0
    g3.prune_s1_end(min_opacity=0.01, extent=4, max_screen_size=1)
    snapshot(g3, rec, "s1_end")
    return rec


def _second_stage(renderer_cls, init_kwargs, setup_kwargs, gen, rec):
    """Adaptive initialisation around control points (Renderer.initialize_ag, :1038-1058) + a densify round with
    per-Gaussian scales, where both branches select rows."""
    np.random.seed(12)
    torch.manual_seed(12)
    r = renderer_cls(sh_degree=0, white_background=True, num_latent_code=2, latent_code_dim=32)
    g = r.gaussians
    r.initialize(num_pts=30, num_cpts=30, radius=0.5, radius2=0.5, **init_kwargs)
    c_xyz = g._c_xyz.detach().clone()
    c_radius = torch.exp(g._c_radius.detach())
    r.initialize_ag(c_xyz, c_radius, num_cpts=30, num_pts_per_cpt=12, init_ratio=1, **init_kwargs)
    g._r = torch.tensor([])                      # prepare_train_s2 (main_train_dimo.py:486)
    snapshot(g, rec, "ag", moments=False)
    g.training_setup(train_args(), **setup_kwargs)
    n = g._xyz.shape[0]
    with torch.no_grad():
        g._scaling.data = torch.log(torch.rand(n, 3, generator=gen) * 0.08 + 0.005)
        g._rotation.data = torch.randn(n, 4, generator=gen)
        g._opacity.data = torch.randn(n, 1, generator=gen) * 2
    for it in range(2):
        fake_backward(g, gen, scale=0.01)
        g.optimizer.step()
        g.optimizer.zero_grad()
    for it in range(3):
        vs = types.SimpleNamespace(grad=0.03 * torch.randn(n, 3, generator=gen))
        vis = torch.rand(n, generator=gen) > 0.2
        radii = torch.randint(0, 3, (n,), generator=gen).float()
        g.max_radii2D[vis] = torch.max(g.max_radii2D[vis], radii[vis])
        g.add_densification_stats(vs, vis)
    torch.manual_seed(29)
    g.densify_and_prune(0.02, min_opacity=0.02, extent=4, max_screen_size=1)
    snapshot(g, rec, "densified2")
    return g
