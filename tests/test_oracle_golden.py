"""CPU: the oracle restatement vs fixtures produced by executing the reference's own code
(tests/golden/make_golden.py).  This is what pins oracle/{deform,sh,camera,loss}.py."""
import os

import numpy as np
import pytest
import torch

from oracle import camera as ocam, deform as od, loss as ol, sh as osh

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return {k: v for k, v in np.load(os.path.join(G, name)).items()}


def T(a):
    return torch.tensor(np.asarray(a))


def test_posenc_bit_exact():
    d = load("posenc.npz")
    assert tuple(d["dims"]) == (60, 12)
    assert torch.equal(od.posenc(T(d["x"]), 10), T(d["ex"]))
    assert torch.equal(od.posenc(T(d["t"]), 6), T(d["et"]))


def test_timenet_matches_reference_module():
    d = load("timenet.npz")
    params = od.timenet_init(32, seed=int(d["seed"]), final_scale=float(d["final_scale"]))
    dx, dq = od.timenet_forward(params, T(d["pts"]), float(d["t"]), T(d["lat"]))
    assert torch.allclose(dx, T(d["dxyz"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(dq, T(d["dquat"]), rtol=1e-5, atol=1e-6)
    # batched t_apply form (arap_loss_v2): pts [1,M,3], t [T,M,1]
    qt = T(d["qt"])
    Tn, M = qt.shape[0], qt.shape[1]
    pts = T(d["pts"])[None].expand(Tn, M, 3).reshape(-1, 3)
    dxb, dqb = od.timenet_forward(params, pts, qt.reshape(-1, 1), T(d["lat"]))
    assert torch.allclose(dxb.reshape(Tn, M, 3), T(d["dxyz_b"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(dqb.reshape(Tn, M, 4), T(d["dquat_b"]), rtol=1e-5, atol=1e-6)


def test_timenet_identity_init():
    d = load("timenet.npz")
    params = od.timenet_init(32, seed=0, final_scale=None)
    dx, dq = od.timenet_forward(params, torch.rand(5, 3), 0.25, torch.randn(32))
    assert torch.equal(dx, T(d["ident_p"])) and torch.equal(dq, T(d["ident_r"]))
    assert float(dx.abs().max()) == 0 and torch.equal(dq, torch.tensor([1., 0, 0, 0]).repeat(5, 1))


def test_timenet_param_count():
    assert sum(o * i + o for o, i in od.timenet_layer_shapes(32)) == 647431


def test_lbs_matches_reference_block():
    d = load("lbs.npz")
    m, r = od.lbs_deform(T(d["xyz"]), T(d["rot"]), T(d["c_xyz"]), torch.exp(T(d["c_radius_raw"])), T(d["dxyz"]),
                         T(d["dquat"]), T(d["idx"]), T(d["dist"]))
    assert torch.allclose(m, T(d["means3D"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(r, T(d["rotations"]), rtol=1e-5, atol=1e-6)
    w = od.lbs_weights(T(d["dist"]), torch.exp(T(d["c_radius_raw"])), T(d["idx"]))
    assert torch.allclose(w, T(d["w"]), rtol=1e-6, atol=1e-7)


def test_sh_matches_reference():
    d = load("sh.npz")
    coef = T(d["coef"])                     # reference layout [N, C, K] -> ours [N, K, C]
    ours = coef.permute(0, 2, 1).contiguous()
    for deg in range(4):
        out = osh.eval_sh(deg, ours, T(d["dirs"]))
        assert torch.allclose(out, T(d[f"deg{deg}"]), rtol=1e-5, atol=1e-6), deg
    assert abs(osh.C0 - float(d["C0"])) == 0
    assert torch.allclose(osh.RGB2SH(torch.tensor([0.2, 0.9])), T(d["rgb2sh"]))


def test_camera_matches_reference():
    d = load("camera.npz")
    for i in range(3):
        el, az, W, H = d[f"args{i}"]
        pose = ocam.orbit_camera(el, az, 2.0)
        assert np.array_equal(pose, d[f"pose{i}"])
        fovy = np.deg2rad(33.9)
        fovx = 2 * np.arctan(np.tan(fovy / 2) * W / H)
        c = ocam.Camera(pose, int(W), int(H), fovy, fovx, 0.01, 100)
        assert torch.equal(c.world_view_transform, T(d[f"view{i}"]))
        assert torch.equal(c.projection_matrix, T(d[f"proj{i}"]))
        assert torch.equal(c.full_proj_transform, T(d[f"full{i}"]))
        assert torch.equal(c.camera_center, T(d[f"center{i}"]))


def test_product_camera_matches_reference():
    """dimo_b200.camera (host NumPy, product side) against the same fixtures, on CPU tensors"""
    from dimo_b200.camera import MiniCam, orbit_camera
    d = load("camera.npz")
    for i in range(3):
        el, az, W, H = d[f"args{i}"]
        pose = orbit_camera(el, az, 2.0)
        assert np.array_equal(pose, d[f"pose{i}"])
        fovy = np.deg2rad(33.9)
        fovx = 2 * np.arctan(np.tan(fovy / 2) * W / H)
        c = MiniCam(pose, int(W), int(H), fovy, fovx, 0.01, 100, device="cpu")
        assert torch.equal(c.world_view_transform, T(d[f"view{i}"]))
        assert torch.equal(c.full_proj_transform, T(d[f"full{i}"]))
        assert torch.equal(c.camera_center, T(d[f"center{i}"]))


def test_ssim_l1_match_reference():
    d = load("loss.npz")
    a = T(d["a"]).requires_grad_(True)
    b = T(d["b"])
    s = ol.ssim(a, b)
    s.backward()
    assert abs(s.item() - float(d["ssim"])) < 1e-6
    assert torch.allclose(a.grad, T(d["dssim_da"]), rtol=1e-4, atol=1e-9)
    assert abs(ol.l1_loss(a, b).item() - float(d["l1"])) < 1e-7
    assert abs(ol.ssim(a.detach(), a.detach()).item() - float(d["ssim_same"])) < 1e-6
    w1 = ol.gaussian_window_1d()
    assert torch.equal(w1[:, None].mm(w1[None, :]), T(d["window"]))


def test_smoothness_regularisers_match_reference():
    """edge-aware depth / bilateral normal smoothness (src/loss.py:64-107): values and gradients"""
    d = load("smooth.npz")
    depth = T(d["depth"]).requires_grad_(True); normal = T(d["normal"]).requires_grad_(True)
    rgb = T(d["rgb"]).requires_grad_(True)
    ld = ol.edge_aware_smoothness(depth, rgb)
    ln = ol.bilateral_normal_smoothness(normal, rgb)
    assert abs(ld.item() - float(d["loss_depth"])) < 1e-6
    assert abs(ln.item() - float(d["loss_normal"])) < 1e-6
    gd = torch.autograd.grad(ld, [depth, rgb], retain_graph=True)
    gn = torch.autograd.grad(ln, [normal, rgb])
    assert torch.allclose(gd[0], T(d["ddepth"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(gd[1], T(d["drgb_depth"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(gn[0], T(d["dnormal"]), rtol=1e-5, atol=1e-9)
    assert torch.allclose(gn[1], T(d["drgb_normal"]), rtol=1e-5, atol=1e-9)
