#!/usr/bin/env python
"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE in the build container.

Run here only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

What is executed from /root/reference (nothing is copied into the repo -- the fixtures hold only
inputs and outputs):
  * src/pos_enc.py            imported as a module            -> posenc.npz
  * src/loss.py               imported (ssim, l1_loss)        -> loss.npz
  * utils/sh_utils.py         imported (eval_sh, RGB2SH)      -> sh.npz
  * utils/cam_utils.py        imported (orbit_camera)         -> camera.npz
  * renderer/latent_gs_renderer.py cannot be imported (plyfile, pytorch3d, diff_gauss ... are absent,
    SURVEY.md 8c), so the needed definitions are exec'd from its source text at generation time with the
    hard-coded 'cuda' device strings mapped to 'cpu':
      - build_rotation_3d, quat_mul, TimeNet (+ its initialisers)        -> timenet.npz
      - the stage-s2 LBS block of Renderer.render ("eps = 1e-7" .. "rotations = quat_mul(...)")
        followed by the rotation activation                               -> lbs.npz
      - getProjectionMatrix + MiniCam                                     -> camera.npz
TimeNet weights are not stored (2.6 MB): they come from oracle.deform.timenet_init(seed), are loaded into
the reference module through its state_dict, and the test regenerates them from the same seed.
"""
import importlib.util
import os
import re
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
from oracle import deform as od  # noqa: E402


def load_module(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def ref_source():
    return open(os.path.join(REF, "renderer/latent_gs_renderer.py")).read()


def extract(src, start_pat, end_pat, include_end=True):
    s = src.index(start_pat)
    e = src.index(end_pat, s)
    if include_end:
        e = src.index("\n", e) + 1
    return src[s:e]


def cpuify(code):
    code = code.replace("device='cuda'", "device='cpu'").replace('device="cuda"', 'device="cpu"')
    code = re.sub(r"\.cuda\(\)", "", code)
    return code


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    g = torch.Generator().manual_seed(42)
    src = ref_source()

    # ---------------- positional encoding ----------------
    pe = load_module("ref_pos_enc", "src/pos_enc.py")
    x = (torch.rand(9, 3, generator=g) - 0.5) * 1.2
    t = torch.rand(9, 1, generator=g)
    emb_x, dim_x = pe.get_embedder(10, 3)
    emb_t, dim_t = pe.get_embedder(6, 1)
    np.savez(os.path.join(HERE, "posenc.npz"), x=x.numpy(), t=t.numpy(), ex=emb_x(x).numpy(), et=emb_t(t).numpy(),
             dims=np.array([dim_x, dim_t]))

    # ---------------- TimeNet ----------------
    ns = {"torch": torch, "nn": torch.nn, "init": torch.nn.init, "F": torch.nn.functional,
          "get_embedder": pe.get_embedder, "np": np}
    code = extract(src, "def build_rotation_3d(r):", "class BasicPointCloud", include_end=False)
    code += extract(src, "def initialize_weights(m):", "class GaussianModel:", include_end=False)
    exec(cpuify(code), ns)
    TimeNet = ns["TimeNet"]
    net = TimeNet(latent_code_dim=32, device="cpu")
    # identity-init property of the reference constructor
    p0, r0 = net(torch.rand(5, 3, generator=g), 0.25, torch.randn(32, generator=g))
    ident = dict(p0=p0.detach().numpy(), r0=r0.detach().numpy())
    seed = 3
    params = od.timenet_init(32, seed=seed, final_scale=0.05)
    names = [f"deformnet.{i}" for i in range(8)] + ["pts_layers.0", "pts_layers.2", "rot_layers.0", "rot_layers.2"]
    sd = {}
    for n, (W, b) in zip(names, params):
        sd[n + ".weight"] = W
        sd[n + ".bias"] = b
    net.load_state_dict(sd)
    pts = (torch.rand(24, 3, generator=g) - 0.5)
    lat = torch.randn(32, generator=g)
    tt = 0.37
    dx, dq = net(pts, tt, lat)
    # t_apply form used by arap_loss_v2
    qt = torch.rand(3, generator=g)[:, None, None].repeat(1, 24, 1)
    dx_b, dq_b = net(pts[None], qt, lat, t_apply=True)
    np.savez(os.path.join(HERE, "timenet.npz"), seed=seed, final_scale=0.05, pts=pts.numpy(), lat=lat.numpy(), t=tt,
             dxyz=dx.detach().numpy(), dquat=dq.detach().numpy(), qt=qt.numpy(), dxyz_b=dx_b.detach().numpy(),
             dquat_b=dq_b.detach().numpy(), ident_p=ident["p0"], ident_r=ident["r0"])

    # ---------------- LBS block of Renderer.render ----------------
    block = extract(src, "eps = 1e-7\n", "rotations = quat_mul(rots3D, rotations)")
    block = textwrap.dedent("            " + block)
    N, M, K = 200, 16, 4
    xyz = torch.rand(N, 3, generator=g) - 0.5
    rot = torch.randn(N, 4, generator=g)
    c_xyz = xyz[torch.randperm(N, generator=g)[:M]].clone()
    c_radius_raw = torch.log(torch.full((M, 1), 0.1)) + 0.2 * torch.randn(M, 1, generator=g)
    dxyz = 0.05 * torch.randn(M, 3, generator=g)
    dquat = torch.tensor([1.0, 0, 0, 0]) + 0.3 * torch.randn(M, 4, generator=g)
    d2 = ((xyz[:, None] - c_xyz[None]) ** 2).sum(-1)
    dist, idx = torch.sort(d2, dim=1)
    dist, idx = torch.sqrt(dist[:, :K]), idx[:, :K]

    class G:  # the attributes the block reads from self.gaussians
        neighbor_dists = dist
        neighbor_indices = idx

        @staticmethod
        def get_c_radius(stage):
            return torch.exp(c_radius_raw)

        rotation_activation = staticmethod(torch.nn.functional.normalize)

    class Self:
        gaussians = G

    loc = dict(self=Self, stage="s2", c_means3D=c_xyz, means3D=xyz, means3D_deform=dxyz, rots_deform=dquat,
               rotations=rot, local_frame=True, torch=torch, F=torch.nn.functional,
               build_rotation_3d=ns["build_rotation_3d"], quat_mul=ns["quat_mul"])
    exec(cpuify(block), loc)
    means_out = loc["means3D"]
    rot_out = G.rotation_activation(loc["rotations"])       # :1219
    np.savez(os.path.join(HERE, "lbs.npz"), xyz=xyz.numpy(), rot=rot.numpy(), c_xyz=c_xyz.numpy(),
             c_radius_raw=c_radius_raw.numpy(), dxyz=dxyz.numpy(), dquat=dquat.numpy(), dist=dist.numpy(),
             idx=idx.numpy(), means3D=means_out.numpy(), rotations=rot_out.numpy(), w=loc["w"].numpy())

    # ---------------- SH ----------------
    sh = load_module("ref_sh", "utils/sh_utils.py")
    coef = torch.randn(11, 3, 16, generator=g)          # reference layout [..., C, (deg+1)^2]
    dirs = torch.nn.functional.normalize(torch.randn(11, 3, generator=g))
    outs = {f"deg{d}": sh.eval_sh(d, coef, dirs).numpy() for d in range(4)}
    np.savez(os.path.join(HERE, "sh.npz"), coef=coef.numpy(), dirs=dirs.numpy(), rgb2sh=sh.RGB2SH(torch.tensor([0.2, 0.9])).numpy(),
             C0=sh.C0, **outs)

    # ---------------- cameras ----------------
    cu = load_module("ref_cam_utils", "utils/cam_utils.py")
    code = extract(src, "def getProjectionMatrix(", "class Renderer:", include_end=False)
    cns = {"torch": torch, "np": np, "math": __import__("math")}
    exec(cpuify(code), cns)
    cams = {}
    for i, (el, az, W, H) in enumerate([(0, 45.0, 512, 512), (-20, 200.0, 128, 96), (0, 0.0, 64, 64)]):
        pose = cu.orbit_camera(el, az, 2.0)
        fovy = np.deg2rad(33.9)
        fovx = 2 * np.arctan(np.tan(fovy / 2) * W / H)
        mc = cns["MiniCam"](pose, W, H, fovy, fovx, 0.01, 100)
        cams[f"pose{i}"] = pose
        cams[f"view{i}"] = mc.world_view_transform.numpy()
        cams[f"proj{i}"] = mc.projection_matrix.numpy()
        cams[f"full{i}"] = mc.full_proj_transform.numpy()
        cams[f"center{i}"] = mc.camera_center.numpy()
        cams[f"args{i}"] = np.array([el, az, W, H], dtype=np.float64)
    np.savez(os.path.join(HERE, "camera.npz"), **cams)

    # ---------------- SSIM / L1 ----------------
    ls = load_module("ref_loss", "src/loss.py")
    a = torch.rand(2, 3, 40, 52, generator=g)
    b = (a + 0.15 * torch.randn(2, 3, 40, 52, generator=g)).clamp(0, 1)
    ag = a.clone().requires_grad_(True)
    s = ls.ssim(ag, b)
    s.backward()
    np.savez(os.path.join(HERE, "loss.npz"), a=a.numpy(), b=b.numpy(), ssim=s.item(), dssim_da=ag.grad.numpy(),
             l1=ls.l1_loss(a, b).item(), ssim_same=ls.ssim(a, a).item(),
             window=ls.create_window(11, 1)[0, 0].numpy())
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
