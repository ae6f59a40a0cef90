"""BASELINE.json configs[0] -- "1k synthetic Gaussians, 1 motion, 1 frame, 64x64: deform MLP + L1 on PyTorch CPU", the
case the reference runs by itself without a GPU: the oracle against tests/golden/c1.npz, which
tests/golden/make_golden_c1.py produced by executing the reference's TimeNet / l1_loss / ssim (forward and autograd
backward).  The CUDA path is checked against the same fixture in tests/test_zz_reference_flow_gpu.py."""
import os

import numpy as np
import torch

from dimo_b200 import synthetic
from oracle import deform as od
from oracle import loss as ol

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1.npz")


def c1_inputs():
    d = np.load(GOLD)
    scene = synthetic.make_scene(1000, n_ctrl=512, n_motions=1, seed=0)
    params = od.timenet_init(32, seed=int(d["seed"]), final_scale=float(d["final_scale"]))
    return d, scene, params


def test_c1_deform_and_l1_match_reference():
    d, scene, params = c1_inputs()
    xyz = scene["_xyz"].clone().requires_grad_(True)
    lat = scene["_latent_codes"][0].clone().requires_grad_(True)
    ps = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in params]
    dxyz, dquat = od.timenet_forward(ps, xyz, float(d["t"]), lat)
    assert xyz.shape == (1000, 3) and dxyz.shape == (1000, 3) and dquat.shape == (1000, 4)
    assert np.abs(dxyz.detach().numpy() - d["dxyz"]).max() <= 1e-5 * np.abs(d["dxyz"]).max()
    assert np.abs(dquat.detach().numpy() - d["dquat"]).max() <= 1e-5 * np.abs(d["dquat"]).max()
    loss = ol.l1_loss(xyz + dxyz, torch.from_numpy(d["target"]))
    assert abs(loss.item() - float(d["l1_points"])) <= 1e-6 * float(d["l1_points"])
    loss.backward()
    for got, key in ((xyz.grad, "d_xyz"), (lat.grad, "d_latent"), (ps[0][0].grad, "d_w0"), (ps[9][0].grad, "d_wp")):
        want = d[key]
        assert np.abs(got.numpy() - want).max() <= 1e-4 * np.abs(want).max(), key


def test_c1_image_losses_match_reference():
    d = np.load(GOLD)
    a, b = torch.from_numpy(d["img_a"]), torch.from_numpy(d["img_b"])
    assert a.shape == (1, 3, 64, 64)
    assert abs(ol.l1_loss(a, b).item() - float(d["l1_images"])) <= 1e-6
    assert abs(ol.ssim(a, b).item() - float(d["ssim_images"])) <= 1e-5
