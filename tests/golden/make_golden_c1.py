#!/usr/bin/env python
"""Generates tests/golden/c1.npz -- BASELINE.json configs[0] ("1k synthetic Gaussians, 1 motion, 1 frame, 64x64: deform
MLP + L1 on PyTorch CPU"), the one case of the path the reference can run by itself without a GPU -- by EXECUTING THE
REFERENCE'S OWN CODE (build container only: needs /root/reference):

  * TimeNet + its initialisers: exec'd from renderer/latent_gs_renderer.py (device strings mapped to 'cpu'), weights
    loaded from oracle.deform.timenet_init(seed) through its state_dict (not stored: the tests regenerate them);
  * src/pos_enc.py, src/loss.py (l1_loss, ssim) imported as modules.

Recorded: stage-s1 deformation of the 1000 Gaussians at t = 0 (Renderer.render :1174-1176, 1211-1212), L1 between the
deformed centres and seeded targets and between two seeded 64x64 images, SSIM of the images, and the gradients of the
L1 term w.r.t. the Gaussian centres, the latent code and two weight tensors (autograd through the reference module).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import cpuify, extract, load_module, ref_source  # noqa: E402
from dimo_b200 import synthetic  # noqa: E402
from oracle import deform as od  # noqa: E402

SEED, FINAL_SCALE = 5, 0.05


def main():
    src = ref_source()
    pe = load_module("ref_pos_enc", "src/pos_enc.py")
    ls = load_module("ref_loss", "src/loss.py")
    ns = {"torch": torch, "nn": torch.nn, "init": torch.nn.init, "F": torch.nn.functional,
          "get_embedder": pe.get_embedder, "np": np}
    code = extract(src, "def build_rotation_3d(r):", "class BasicPointCloud", include_end=False)
    code += extract(src, "def initialize_weights(m):", "class GaussianModel:", include_end=False)
    exec(cpuify(code), ns)
    net = ns["TimeNet"](latent_code_dim=32, device="cpu")
    params = od.timenet_init(32, seed=SEED, final_scale=FINAL_SCALE)
    names = [f"deformnet.{i}" for i in range(8)] + ["pts_layers.0", "pts_layers.2", "rot_layers.0", "rot_layers.2"]
    net.load_state_dict({f"{n}.{k}": v for n, (W, b) in zip(names, params) for k, v in (("weight", W), ("bias", b))})

    scene = synthetic.make_scene(1000, n_ctrl=512, n_motions=1, seed=0)
    g = torch.Generator().manual_seed(17)
    xyz = scene["_xyz"].clone().requires_grad_(True)
    lat = scene["_latent_codes"][0].clone().requires_grad_(True)
    t = 0.0                                                    # 1 frame: source_time = [0 / 1]
    dxyz, dquat = net(xyz, t, lat)
    means = xyz + dxyz                                         # stage s1 (:1211-1212)
    target = means.detach() + 0.01 * torch.randn(1000, 3, generator=g)
    loss = ls.l1_loss(means, target)
    loss.backward()
    a = torch.rand(1, 3, 64, 64, generator=g)
    b = (a + 0.1 * torch.randn(1, 3, 64, 64, generator=g)).clamp(0, 1)
    np.savez_compressed(
        os.path.join(HERE, "c1.npz"), seed=SEED, final_scale=FINAL_SCALE, t=t, target=target.numpy(),
        dxyz=dxyz.detach().numpy(), dquat=dquat.detach().numpy(), l1_points=loss.item(),
        d_xyz=xyz.grad.numpy(), d_latent=lat.grad.numpy(),
        d_w0=net.deformnet[0].weight.grad.numpy(), d_wp=net.pts_layers[2].weight.grad.numpy(),
        img_a=a.numpy(), img_b=b.numpy(), l1_images=ls.l1_loss(a, b).item(), ssim_images=ls.ssim(a, b).item())
    print("c1.npz", os.path.getsize(os.path.join(HERE, "c1.npz")), "l1", loss.item(), "|dxyz|max", float(dxyz.abs().max()))


if __name__ == "__main__":
    main()
