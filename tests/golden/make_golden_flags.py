#!/usr/bin/env python
"""tests/golden/flags.npz: the non-default branches of the reference's Renderer.render, produced by EXECUTING the
reference's source (build container only):
  * the stage-s2 skinning block ("eps = 1e-7" .. "rotations = quat_mul(...)", renderer/latent_gs_renderer.py:1192-1209)
    with local_frame=False, followed by the rotation activation (:1219);
  * the convert_SHs_python colour block (:1228-1238) with utils/sh_utils.eval_sh, degrees 0..3."""
import os
import sys
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from make_golden import cpuify, extract, load_module, ref_source  # noqa: E402


def main():
    src = ref_source()
    g = torch.Generator().manual_seed(21)
    ns = {"torch": torch, "np": np}
    exec(cpuify(extract(src, "def build_rotation_3d(r):", "class BasicPointCloud", include_end=False)), ns)
    block = textwrap.dedent("            " + extract(src, "eps = 1e-7\n", "rotations = quat_mul(rots3D, rotations)"))
    N, M, K = 150, 12, 4
    xyz = torch.rand(N, 3, generator=g) - 0.5
    rot = torch.randn(N, 4, generator=g)
    c_xyz = xyz[torch.randperm(N, generator=g)[:M]].clone()
    c_radius = torch.exp(torch.log(torch.full((M, 1), 0.1)) + 0.2 * torch.randn(M, 1, generator=g))
    dxyz = 0.05 * torch.randn(M, 3, generator=g)
    dquat = torch.tensor([1.0, 0, 0, 0]) + 0.3 * torch.randn(M, 4, generator=g)
    d2 = ((xyz[:, None] - c_xyz[None]) ** 2).sum(-1)
    dist, idx = torch.sort(d2, dim=1)
    dist, idx = torch.sqrt(dist[:, :K]), idx[:, :K]

    class G:
        neighbor_dists, neighbor_indices = dist, idx
        get_c_radius = staticmethod(lambda stage: c_radius)
        rotation_activation = staticmethod(torch.nn.functional.normalize)

    loc = dict(self=type("S", (), {"gaussians": G}), stage="s2", c_means3D=c_xyz, means3D=xyz, means3D_deform=dxyz,
               rots_deform=dquat, rotations=rot, local_frame=False, torch=torch, F=torch.nn.functional,
               build_rotation_3d=ns["build_rotation_3d"], quat_mul=ns["quat_mul"])
    exec(cpuify(block), loc)
    out = dict(xyz=xyz.numpy(), rot=rot.numpy(), c_radius=c_radius.numpy(), dxyz=dxyz.numpy(), dquat=dquat.numpy(),
               dist=dist.numpy(), idx=idx.numpy(), means3D=loc["means3D"].numpy(),
               rotations=G.rotation_activation(loc["rotations"]).numpy())

    # convert_SHs_python branch
    sh = load_module("ref_sh", "utils/sh_utils.py")
    cblock = textwrap.dedent("                " + extract(src, "shs_view = self.gaussians.get_features.transpose(1, 2).view(",
                                                          "colors_precomp = torch.clamp_min(sh2rgb + 0.5, 0.0)"))
    campos = torch.tensor([0.4, -0.3, 2.0])
    for deg in range(4):
        feats = 0.4 * torch.randn(N, (deg + 1) ** 2, 3, generator=g)

        class GG:
            get_features, get_xyz, max_sh_degree, active_sh_degree = feats, xyz, deg, deg

        cl = dict(self=type("S", (), {"gaussians": GG}), viewpoint_camera=type("C", (), {"camera_center": campos}),
                  eval_sh=sh.eval_sh, torch=torch)
        exec(cblock, cl)
        out[f"feats{deg}"], out[f"colors{deg}"] = feats.numpy(), cl["colors_precomp"].numpy()
    out["campos"] = campos.numpy()
    np.savez_compressed(os.path.join(HERE, "flags.npz"), **out)
    print("flags.npz", os.path.getsize(os.path.join(HERE, "flags.npz")))


if __name__ == "__main__":
    main()
