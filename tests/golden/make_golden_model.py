#!/usr/bin/env python
"""Generates tests/golden/model.npz and tests/golden/arap.npz by EXECUTING THE REFERENCE'S OWN CLASSES on the CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_model.py

What is executed from /root/reference (nothing is copied into the repo; the fixtures hold inputs and outputs only):
  * renderer/latent_gs_renderer.py -- the WHOLE module (GaussianModel, Renderer.initialize / initialize_ag,
    get_expon_lr_func, ...), exec'd from its source text with the hard-coded 'cuda' device strings mapped to 'cpu'.
    Its imports that do not exist in this image are served as follows:
        plyfile, pytorch3d, diff_gauss, diff_gaussian_rasterization   dimo_b200/shims (import only; not exercised here,
                                                                      except quaternion_to_matrix, which is replaced by
                                                                      the reference's in-tree copy, deform_utils.py:17-35)
        simple_knn._C.distCUDA2                                       oracle.knn.dist3nn (CPU)
        src.helpers (needs open3d, matplotlib)                        stub module (o3d_knn is never called)
  * utils/deform_utils.py -- cal_connectivity_from_points_v2 + cal_arap_error (the ARAP term of the step), same
    treatment; pytorch3d.ops.ball_query is served by oracle.points.ball_query.

The scenario itself is tests/model_scenario.py, shared with tests/test_model_cpu.py.
"""
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dimo_b200  # noqa: E402
from oracle import knn as oknn  # noqa: E402
from oracle import points as opoints  # noqa: E402
import model_scenario  # noqa: E402


def cpuify(code):
    code = code.replace("device='cuda'", "device='cpu'").replace('device="cuda"', 'device="cpu"')
    code = code.replace('.to("cuda")', '.to("cpu")').replace('map_location="cuda"', 'map_location="cpu"')
    code = re.sub(r"\.cuda\(\)", "", code)
    return code


def load_patched(name, rel):
    src = cpuify(open(os.path.join(REF, rel)).read())
    mod = types.ModuleType(name)
    mod.__file__ = os.path.join(REF, rel)
    sys.modules[name] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


def reference_modules():
    dimo_b200.install_shims()
    sys.path.append(REF)                                   # utils.sh_utils, src.pos_enc import as they are
    helpers = types.ModuleType("src.helpers")
    helpers.o3d_knn = None
    sys.modules["src.helpers"] = helpers
    import simple_knn._C as sk                             # the shim; give it a CPU implementation for this script
    sk.distCUDA2 = lambda pts: oknn.dist3nn(pts.float())
    import pytorch3d.ops
    pytorch3d.ops.ball_query = opoints.ball_query
    deform_utils = load_patched("utils.deform_utils", "utils/deform_utils.py")
    import pytorch3d.transforms
    pytorch3d.transforms.quaternion_to_matrix = deform_utils.quaternion_to_matrix
    renderer = load_patched("renderer.latent_gs_renderer", "renderer/latent_gs_renderer.py")
    return renderer, deform_utils


def arap_fixture(deform_utils):
    g = torch.Generator().manual_seed(77)
    out = {}
    for tag, (T, M, spread, amp) in {"a": (4, 96, 0.22, 0.01), "b": (8, 64, 0.16, 0.02)}.items():
        base = (torch.rand(M, 3, generator=g) - 0.5) * 2 * spread
        # a smooth time-dependent deformation so that neighbourhoods mostly persist, plus noise
        frames = []
        for t in range(T):
            ang = 0.15 * t
            rot = torch.tensor([[np.cos(ang), -np.sin(ang), 0.0], [np.sin(ang), np.cos(ang), 0.0], [0.0, 0.0, 1.0]],
                               dtype=torch.float32)
            frames.append(base @ rot.T * (1 + 0.03 * t) + amp * torch.randn(M, 3, generator=g))
        nodes = torch.stack(frames).requires_grad_(True)
        ii, jj, nn, _ = deform_utils.cal_connectivity_from_points_v2(nodes.detach(), K=10)
        err = deform_utils.cal_arap_error(nodes, ii, jj, nn)
        (grad,) = torch.autograd.grad(err, nodes)
        out[f"{tag}/nodes"] = nodes.detach().numpy()
        out[f"{tag}/ii"], out[f"{tag}/jj"], out[f"{tag}/nn"] = ii.numpy(), jj.numpy(), nn.numpy()
        out[f"{tag}/error"] = np.float64(err.item())
        out[f"{tag}/grad"] = grad.numpy()
        print(f"arap {tag}: T={T} M={M} edges={len(ii)} error={err.item():.6f}")
    return out


def main():
    renderer, deform_utils = reference_modules()
    rec = model_scenario.run(renderer.Renderer)
    np.savez_compressed(os.path.join(HERE, "model.npz"), **rec)
    print("model.npz:", len(rec), "arrays;",
          "N after densify:", rec["densified/_xyz"].shape[0], rec["densified2/_xyz"].shape[0],
          "after prune:", rec["pruned/_xyz"].shape[0], "after fps prune:", rec["fps_pruned/_xyz"].shape[0])
    np.savez_compressed(os.path.join(HERE, "arap.npz"), **arap_fixture(deform_utils))


if __name__ == "__main__":
    main()
