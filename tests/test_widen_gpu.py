"""-m gpu: the rows widened from SURVEY.md 8f -- point-set kernels (FPS, ball query, chamfer) against the CPU oracle
through the C ABI, the GaussianModel life cycle on the fused optimizer (flat re-layout after densify / prune) against
the same life on torch.optim.Adam, a training step that survives a change of N, and the pytorch3d / chamferdist
shims at the reference's call sites."""
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


# ------------------------------------------------------------------------------------------------------------------
# csrc/points.cu
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,K", [(1, 1), (37, 37), (5000, 64), (20000, 512)])
def test_fps_bit_exact(cuda, N, K):
    from dimo_b200 import points
    from oracle import points as opoints
    g = torch.Generator().manual_seed(N)
    pts = torch.rand(N, 3, generator=g) - 0.5
    if N > 100:
        pts[N // 2] = pts[3]                                   # a duplicate: tie on the min-distance
    sel, idx = points.sample_farthest_points(pts.cuda()[None], K)
    want = opoints.fps(pts, K)
    assert idx.shape == (1, K) and idx.dtype == torch.int64
    assert torch.equal(idx[0].cpu(), want)
    assert torch.equal(sel[0].cpu(), pts[want])


def test_fps_batched_and_c5_size(cuda):
    from dimo_b200 import points
    from oracle import points as opoints
    g = torch.Generator().manual_seed(1)
    pts = torch.rand(3, 700, 3, generator=g)
    _sel, idx = points.sample_farthest_points(pts.cuda(), 50)
    for b in range(3):
        assert torch.equal(idx[b].cpu(), opoints.fps(pts[b], 50))
    big = torch.rand(1, 500_000, 3, generator=g).cuda()        # c5: 500k Gaussians -> 512 key points
    _s, bi = points.sample_farthest_points(big, 512)
    bi = bi[0]
    assert int(bi[0]) == 0 and len(set(bi.tolist())) == 512
    # farthest-point property: every pick is at least as far from the earlier picks as any later pick is
    p = big[0, bi]
    d = torch.cdist(p.double(), p.double())
    gap = torch.stack([d[k, :k].min() for k in range(1, 512)])
    assert bool((gap[1:] <= gap[:-1] * (1 + 1e-5)).all())


@pytest.mark.parametrize("B,P,K,radius", [(1, 5, 3, 0.1), (3, 300, 11, 0.1), (8, 512, 11, 0.1), (2, 2500, 7, 0.05)])
def test_ball_query_bit_exact(cuda, B, P, K, radius):
    from dimo_b200 import points
    from oracle import points as opoints
    g = torch.Generator().manual_seed(B * 1000 + P)
    p = (torch.rand(B, P, 3, generator=g) - 0.5) * (0.6 if P < 1000 else 1.0)
    d, idx, nn = points.ball_query(p.cuda(), p.cuda(), K=K, radius=radius)
    d0, i0, n0 = opoints.ball_query(p, p, K=K, radius=radius)
    assert torch.equal(idx.cpu(), i0)
    assert torch.equal(d.cpu(), d0)
    assert torch.equal(nn.cpu(), n0)
    if P >= 300:
        assert int((i0 >= 0).sum()) > B * P                   # the case has real neighbours, not only self hits


def test_chamfer_forward_backward(cuda):
    from dimo_b200 import points
    from oracle import points as opoints
    g = torch.Generator().manual_seed(2)
    for n, m in ((512, 512), (1, 3), (3000, 700)):
        a = torch.randn(n, 3, generator=g) * 0.3
        b = torch.randn(m, 3, generator=g) * 0.3
        a0, b0 = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        v0 = opoints.chamfer_forward(a0, b0)
        (v0 * 1.7).backward()
        a1, b1 = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        v1 = points.chamfer_forward(a1[None], b1[None])
        (v1 * 1.7).backward()
        assert abs(v1.item() - v0.item()) <= 1e-5 * abs(v0.item())
        assert (a1.grad.cpu() - a0.grad).abs().max() <= 1e-5 * a0.grad.abs().max()
        assert (b1.grad.cpu() - b0.grad).abs().max() <= 1e-5 * b0.grad.abs().max()
        # the reference's call: target detached (main_train_dimo.py:297-299)
        a2 = a.cuda().requires_grad_(True)
        points.ChamferDistance()(a2[None], b.cuda()[None]).backward()
        assert (a2.grad.cpu() * 1.7 - a0.grad).abs().max() <= 1e-5 * a0.grad.abs().max()


# ------------------------------------------------------------------------------------------------------------------
# GaussianModel on the fused optimizer
# ------------------------------------------------------------------------------------------------------------------
def _life(kind, seed=0):
    """A scripted life on a CUDA model: steps, statistics, densify_and_prune, prune, index prune, opacity reset, steps.
    Yields (tag, model) after every stage.  Everything random is seeded, so optimizer kinds can be compared."""
    import model_scenario as ms
    from dimo_b200 import synthetic
    from dimo_b200.renderer import Renderer
    torch.manual_seed(7)                   # TimeNet's xavier init: both optimizer kinds start from the same weights
    r = Renderer(sh_degree=0, device="cuda", num_latent_code=2)
    g = r.gaussians
    g.load_state(synthetic.make_scene(1500, n_ctrl=32, n_motions=2, seed=4))
    g.spatial_lr_scale = 1
    g.training_setup(ms.train_args(), optimizer=kind)
    gen = torch.Generator(device="cuda").manual_seed(seed)

    def step(k=1):
        for _ in range(k):
            for grp in g.optimizer.param_groups:
                for p in grp["params"]:
                    gr = 0.01 * torch.randn(p.shape, generator=gen, device="cuda")
                    if p.grad is None:
                        p.grad = gr
                    else:
                        p.grad.copy_(gr)                  # fused: a view of the flat gradient buffer
            g.optimizer.step()
            g.optimizer.zero_grad()

    step(3)
    yield "steps", g
    n = g._xyz.shape[0]
    for _ in range(3):
        vs = types.SimpleNamespace(grad=0.03 * torch.randn(n, 3, generator=gen, device="cuda"))
        vis = torch.rand(n, generator=gen, device="cuda") > 0.25
        radii = torch.randint(0, 3, (n,), generator=gen, device="cuda").float()
        g.max_radii2D[vis] = torch.max(g.max_radii2D[vis], radii[vis])
        g.add_densification_stats(vs, vis)
    with torch.no_grad():
        g._scaling.data.copy_(torch.log(torch.rand(n, 3, generator=gen, device="cuda") * 0.08 + 0.005))
    torch.manual_seed(31)
    g.densify_and_prune(0.02, min_opacity=0.1, extent=4, max_screen_size=1)
    yield "densified", g
    step(2)
    yield "steps2", g
    with torch.no_grad():
        g._opacity.data[::5] = -7.0
    g.prune(min_opacity=0.01, extent=4)
    yield "pruned", g
    idx = torch.randperm(g._xyz.shape[0], generator=gen, device="cuda")[:200]
    g.prune_points(idx)
    yield "index_pruned", g
    step(1)
    g.reset_opacity()
    yield "reset", g
    step(2)
    yield "steps3", g


def _state(g, kind):
    out = {}
    for name in ("_xyz", "_features_dc", "_opacity", "_scaling", "_rotation", "_c_xyz", "_c_radius", "_latent_codes"):
        p = getattr(g, name)
        out[name] = p.detach().cpu().clone()
        if kind == "fused":
            m, v = g.optimizer.moments(p)
        else:
            st = g.optimizer.state[p]
            m, v = st["exp_avg"], st["exp_avg_sq"]
        out[name + "/m"], out[name + "/v"] = m.detach().cpu().clone(), v.detach().cpu().clone()
    w = g._timenet.deformnet[3].weight
    out["timenet"] = w.detach().cpu().clone()
    for name in ("max_radii2D", "xyz_gradient_accum", "denom"):
        out[name] = getattr(g, name).detach().cpu().clone()
    return out


def test_lifecycle_fused_equals_torch_adam(cuda):
    """Every densify / prune / reset re-lays the fused optimizer's flat buffers out; parameters AND Adam moments must
    follow the life the reference's per-group state surgery produces (optimizer="torch" is that surgery on
    torch.optim.Adam, pinned against the reference's own class by tests/test_model_cpu.py)."""
    sizes = []
    for (tag_f, gf), (tag_t, gt) in zip(_life("fused"), _life("torch")):
        assert tag_f == tag_t
        a, b = _state(gf, "fused"), _state(gt, "torch")
        sizes.append(a["_xyz"].shape[0])
        for k in a:
            assert a[k].shape == b[k].shape, (tag_f, k, a[k].shape, b[k].shape)
            if a[k].numel() == 0:
                continue
            scale = float(b[k].abs().max()) + 1e-12
            assert float((a[k] - b[k]).abs().max()) <= 2e-5 * scale + 1e-9, (tag_f, k)
        # flat layout invariants
        flat, gflat = gf.optimizer.flat, gf.reducer.flat
        for p in gf.optimizer.reducer.params:
            assert flat.data_ptr() <= p.data_ptr() < flat.data_ptr() + flat.numel() * 4
            assert gflat.data_ptr() <= p.grad.data_ptr() < gflat.data_ptr() + gflat.numel() * 4
        assert float(gflat.abs().max()) == 0.0
    assert sizes[0] == 1500 and sizes[1] != 1500 and sizes[4] == 200 and len(set(sizes)) >= 4
    assert int(gf.optimizer.state[0]) == 8


def test_create_from_pcd_uses_dist3nn_kernel(cuda):
    from dimo_b200.renderer import Renderer
    from oracle import knn as oknn
    np.random.seed(3)
    r = Renderer(sh_degree=0, device="cuda", num_latent_code=2)
    r.initialize(num_pts=3000, num_cpts=64)
    g = r.gaussians
    assert g._xyz.shape == (3000, 3) and g._c_xyz.shape == (64, 3) and g._r.shape == (1, 1)
    d2 = oknn.dist3nn(g._xyz.detach().cpu()).clamp_min(1e-7)
    want = torch.log(torch.sqrt(d2))[:, None].repeat(1, 3)
    assert torch.allclose(g._scaling.detach().cpu(), want, rtol=0, atol=2e-6)
    assert torch.allclose(g._opacity.detach().cpu(), torch.full((3000, 1), math.log(0.05 / 0.95)), atol=1e-6)


def test_train_step_survives_densify_and_prune(cuda):
    """Steps -> densification statistics from the step's own means2D gradients -> densify_and_prune (N changes, flat
    buffers re-laid out, moments carried) -> more steps, eager and graph mode; FPS key-point pruning as in GUI.FPS."""
    from dimo_b200 import synthetic
    from dimo_b200.camera import orbit_minicam
    from dimo_b200.renderer import Renderer
    from dimo_b200.trainstep import TrainStep
    import model_scenario as ms
    W = H = 96
    r = Renderer(sh_degree=0, device="cuda", num_latent_code=2)
    g = r.gaussians
    g.load_state({k: v for k, v in synthetic.make_scene(3000, n_ctrl=64, n_motions=2, seed=1).items()})
    g.spatial_lr_scale = 1
    g.training_setup(ms.train_args(), optimizer="fused")
    g.active_sh_degree = 0
    ts = TrainStep(r, stage="s2", graph=False)
    assert ts.opt is g.optimizer and ts.reducer is g.reducer
    cams = [orbit_minicam(v, 8, W, H) for v in (0, 3, 0, 3)]
    times, lat = [0.1, 0.1, 0.6, 0.6], [0, 0, 1, 1]
    gen = torch.Generator().manual_seed(0)
    gt = torch.rand(4, 3, H, W, generator=gen).cuda()
    mask = torch.rand(4, 1, H, W, generator=gen).cuda()
    l0 = [ts.run(cams, times, lat, gt, mask, n_motions=2).item() for _ in range(3)]
    assert all(math.isfinite(x) for x in l0)
    m_before, _ = g.optimizer.moments(g._c_xyz)
    m_before = m_before.clone()
    n0 = g._xyz.shape[0]
    # one reference-signature render for the densification statistics (viewspace_points.grad)
    g.find_knn(4)
    out = r.render(cams[0], time=0.1, stage="s2", latent_index=0)
    (out["image"].sum() + out["alpha"].sum()).backward()
    vis, radii = out["visibility_filter"], out["radii"]
    g.max_radii2D[vis] = torch.max(g.max_radii2D[vis], radii[vis].float())
    g.add_densification_stats(out["viewspace_points"], vis)
    g.optimizer.zero_grad()
    g.reducer.zero()
    assert float(g.denom.sum()) == float(vis.sum())
    thr = float((g.xyz_gradient_accum / g.denom.clamp_min(1)).flatten().quantile(0.8))
    g.densify_and_prune(thr, min_opacity=0.06, extent=4, max_screen_size=None)
    n1 = g._xyz.shape[0]
    assert n1 != n0 and ts.opt is g.optimizer and ts.reducer is g.reducer
    assert g.reducer.flat.numel() >= 14 * n1
    for p in g.parameters():
        if p.numel():
            assert p.grad is not None and p.grad.data_ptr() >= g.reducer.flat.data_ptr()
            assert p.data_ptr() >= g.optimizer.flat.data_ptr()
    m_after, _ = g.optimizer.moments(g._c_xyz)
    assert torch.equal(m_after, m_before), "moments of untouched groups must survive the re-layout"
    assert int(g.optimizer.state[0]) == 3
    l1 = [ts.run(cams, times, lat, gt, mask, n_motions=2).item() for _ in range(2)]
    assert all(math.isfinite(x) for x in l1)
    assert int(g.optimizer.state[0]) == 5
    # graph mode picks the new layout up (re-probe + re-capture).  Nothing of an earlier backward may be alive at
    # capture time: live autograd graphs pin the parameters' AccumulateGrad nodes to the stream they were created on.
    del out, vis, radii
    tg = TrainStep(r, stage="s2", graph=True, probe_steps=2)
    lg = [tg.run(cams, times, lat, gt, mask, n_motions=2).item() for _ in range(4)]
    assert tg.graph_error is None, tg.graph_error
    assert tg.graph is not None
    g.reset_opacity()
    assert tg.graph is None, "a re-layout must invalidate the captured graph"
    assert float(torch.sigmoid(g._opacity).max()) <= 0.0100001
    lg2 = [tg.run(cams, times, lat, gt, mask, n_motions=2).item() for _ in range(4)]
    assert tg.graph_error is None, tg.graph_error
    assert tg.graph is not None and all(math.isfinite(x) for x in lg + lg2)
    # GUI.FPS (main_train_dimo.py:511-515) through the pytorch3d shim
    import dimo_b200
    dimo_b200.install_shims()
    import pytorch3d.ops as ops
    _, idxs = ops.sample_farthest_points(points=g._xyz.unsqueeze(0), K=64)
    xyz_before = g._xyz.detach().clone()
    g.prune_points(idxs[0])
    assert g._xyz.shape[0] == 64
    assert torch.equal(g._xyz.detach(), xyz_before[-idxs[0] - 1])          # the reference's `~idx` row selection


def test_arap_loss_v2_against_oracle(cuda):
    from dimo_b200 import synthetic
    from dimo_b200.renderer import Renderer
    from oracle import deform as odeform
    from oracle import points as opoints
    r = Renderer(sh_degree=0, device="cuda", num_latent_code=2)
    g = r.gaussians
    sc = synthetic.make_scene(2000, n_ctrl=256, n_motions=2, seed=2)
    sc["_c_xyz"] = sc["_c_xyz"] * 0.45                       # denser key points: real neighbourhoods inside r = 0.1
    g.load_state(sc)
    with torch.no_grad():                                      # a live deformation (the reference init is the identity)
        g._timenet.pts_layers[-1].weight.normal_(0, 0.004)
    torch.manual_seed(5)
    import gpu_parity as gp
    from dimo_b200 import deform as ddeform
    ddeform.DEBUG_CAPTURE = []
    try:
        err, (ii, jj, nn, nbr) = r.arap_loss_v2(stage="s2", latent_index=1)      # fused kernels (CUDA tensors)
        cmasks = gp.cuda_relu_masks(ddeform.DEBUG_CAPTURE[0])                     # rows ordered [time sample][key point]
    finally:
        ddeform.DEBUG_CAPTURE = None
    err.backward()
    assert len(ii) > 50 and math.isfinite(err.item()) and err.item() > 0
    # oracle: same time samples, TimeNet + connectivity + energy on the CPU
    torch.manual_seed(5)
    q = torch.rand(8).to("cuda").cpu()
    params = [(l.weight.detach().cpu(), l.bias.detach().cpu()) for l in list(g._timenet.deformnet) +
              [g._timenet.pts_layers[0], g._timenet.pts_layers[2], g._timenet.rot_layers[0], g._timenet.rot_layers[2]]]
    c = g._c_xyz.detach().cpu().clone().requires_grad_(True)
    lat = g._latent_codes.detach().cpu()[1]
    # the oracle's TimeNet runs on the activation pattern the CUDA forward chose (gpu_parity.run_step_pair explains why)
    M = c.shape[0]
    frames = [c.detach() + odeform.timenet_forward(params, c, float(t), lat,
                                                   masks_in=[m[k * M:(k + 1) * M] for m in cmasks])[0]
              for k, t in enumerate(q)]
    nodes = torch.stack(frames)
    oi, oj, on = opoints.arap_connectivity_v2(nodes.detach())
    # the GPU and CPU node positions differ by ~1e-7, so a pair sitting on the ball's surface may flip: allow 2 edges
    diff = set(zip(ii.tolist(), jj.tolist())) ^ set(zip(oi.tolist(), oj.tolist()))
    assert len(diff) <= 2, diff
    # energy and gradient on the product's own edge list (exact comparison of the arithmetic)
    want = opoints.arap_error(nodes, ii.cpu(), jj.cpu(), nn.cpu())
    assert abs(err.item() - want.item()) <= 1e-4 * want.item()
    want.backward()
    e = (g._c_xyz.grad.cpu() - c.grad).abs().flatten()
    scale = float(c.grad.abs().max())
    assert float(e.max()) <= 2e-4 * scale, float(e.max()) / scale        # (ARAP's fp64 Jacobi SVD vs torch.linalg.svd: 2e-4)


def test_keypoint_trajectory_loss(cuda):
    from dimo_b200 import regularisers
    g = torch.Generator().manual_seed(9)
    a = (torch.randn(512, 3, generator=g) * 0.2).cuda().requires_grad_(True)
    b = (a.detach() + 0.01 * torch.randn(512, 3, generator=g).cuda())
    v = regularisers.keypoint_trajectory_loss(a, b, chamfer=True)
    d2 = torch.cdist(a.detach(), b) ** 2
    assert abs(v.item() - 10.0 * d2.min(dim=1).values.sum().item()) <= 1e-4 * v.item()
    v.backward()
    assert a.grad.abs().sum() > 0
    w = regularisers.keypoint_trajectory_loss(a, b, chamfer=False)
    assert abs(w.item() - 10000.0 * (a.detach() - b).abs().mean().item()) <= 1e-5 * w.item()


# ------------------------------------------------------------------------------------------------------------------
# ground truth resident in HBM
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_gt_cache_fetch_matches_interpolate(cuda, dtype):
    from dimo_b200.data import GroundTruthCache
    from oracle import data as odata
    gen = torch.Generator().manual_seed(6)
    cache = GroundTruthCache(n_motions=2, n_views=3, n_frames=4, size=96, dtype=dtype)
    host = torch.zeros(24, 4, 96, 96, dtype=dtype)
    for m in range(2):
        for v in range(3):
            for f in range(4):
                if dtype == torch.uint8:
                    frame = torch.randint(0, 256, (4, 96, 96), generator=gen, dtype=torch.uint8)
                    img, msk = frame[:3].float() / 255.0, frame[3:].float() / 255.0     # what the reference's loader returns
                else:
                    frame = torch.rand(4, 96, 96, generator=gen)
                    img, msk = frame[:3], frame[3:]
                host[cache.slot(m, v, f)] = frame
                cache.put(m, v, f, img[None].cuda(), msk[None].cuda())
    assert torch.equal(cache.store.cpu(), host)
    triples = [(1, 2, 3), (0, 0, 0), (1, 0, 2), (0, 2, 1), (1, 2, 3)]
    slots = [cache.slot(*t) for t in triples]
    for res in (96, 48, 32, 128, 77):
        rgb, mask = cache.fetch(triples, res)
        want_rgb, want_mask = odata.fetch(host, slots, res)
        assert rgb.shape == (5, 3, res, res) and mask.shape == (5, 1, res, res)
        if res == 96:
            assert torch.equal(rgb.cpu(), want_rgb) and torch.equal(mask.cpu(), want_mask)   # same size: exact copy
        else:
            # the source coordinate scale * (dst + 0.5) - 0.5 is an fp32 number of magnitude <= 96: its fraction (the
            # tap weight) carries ~4e-6 of rounding, fused or not -- the same spread exists between torch's own CPU
            # and CUDA kernels
            assert float((rgb.cpu() - want_rgb).abs().max()) <= 2e-5
            assert float((mask.cpu() - want_mask).abs().max()) <= 2e-5
    with pytest.raises(IndexError):
        cache.fetch([(2, 0, 0)])
    if dtype == torch.uint8:
        with pytest.raises(ValueError):
            cache.put(0, 0, 0, torch.full((1, 3, 96, 96), 0.123).cuda(), torch.zeros(1, 1, 96, 96).cuda())


# ------------------------------------------------------------------------------------------------------------------
# fused ARAP kernels vs the torch formulation (same device) and the host build of the same source
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,M,spread", [(8, 512, 0.45), (4, 96, 0.22), (2, 7, 0.05), (1, 30, 0.2), (3, 1500, 0.6)])
def test_arap_kernels_match_torch_formulation(cuda, T, M, spread):
    from dimo_b200 import regularisers as reg
    g = torch.Generator().manual_seed(T * 1000 + M)
    base = (torch.rand(M, 3, generator=g) - 0.5) * 2 * spread
    nodes = torch.stack([base + 0.004 * t * torch.randn(M, 3, generator=g) for t in range(T)])
    if T > 2:
        nodes[-1] = nodes[0]                                   # identical frame: the "unchanged vertex" rule, R = I
    a = nodes.cuda().requires_grad_(True)
    b = nodes.cuda().requires_grad_(True)
    np.random.seed(5)
    e_f, conn = reg.arap_loss_points(a, fused=True)
    if T == 1:                                                   # no target frame: zero energy, zero gradient
        assert e_f.item() == 0.0
        e_f.backward()
        assert float(a.grad.abs().max()) == 0.0
        return
    np.random.seed(5)
    e_t, (ii, jj, nn, nbr) = reg.arap_loss_points(b, fused=False)
    assert torch.equal(conn.nbr, nbr), "neighbour tables differ"
    assert torch.equal(conn.count.long(), (nbr >= 0).sum(1))
    fi, fj, fn, _ = conn                                         # unpacks like the reference's tuple
    assert torch.equal(fi, ii) and torch.equal(fj, jj) and torch.equal(fn, nn)
    scale = max(e_t.item(), 1e-12)
    assert abs(e_f.item() - e_t.item()) <= 3e-5 * scale, (e_f.item(), e_t.item())
    if T > 1 and e_t.item() > 0:
        (2.5 * e_f).backward()
        (2.5 * e_t).backward()
        gs = float(b.grad.abs().max())
        assert float((a.grad - b.grad).abs().max()) <= 2e-4 * gs
    if M > 512:
        assert int(conn.count.sum()) > 0                       # the sampled branch (np.random.choice) was exercised
