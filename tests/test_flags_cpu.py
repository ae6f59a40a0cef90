"""The non-default render() flags (dimo_b200/flags.py: plain tensor expressions) against tests/golden/flags.npz,
produced by executing the corresponding blocks of the reference's Renderer.render."""
import os

import numpy as np
import torch

from dimo_b200 import flags
from dimo_b200.gaussian_model import covariance_from_scaling_rotation, quat_to_rotmat

D = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flags.npz"))
T = lambda k: torch.from_numpy(D[k])


def test_global_frame_skinning_matches_reference_block():
    means, rots = flags.lbs_global_frame(T("xyz"), T("rot"), T("c_radius"), T("dxyz"), T("dquat"), T("idx"), T("dist"))
    assert torch.allclose(means, T("means3D"), rtol=1e-5, atol=1e-6)
    assert torch.allclose(rots, T("rotations"), rtol=1e-5, atol=1e-6)


def test_python_sh_colours_match_reference_block():
    for deg in range(4):
        got = flags.sh_colors(deg, T(f"feats{deg}"), T("xyz"), T("campos"))
        assert torch.allclose(got, T(f"colors{deg}"), rtol=1e-5, atol=1e-6), deg
        assert float(got.min()) >= 0.0


def test_covariance_equivalence_behind_compute_cov3d_python():
    """render(compute_cov3D_python=True) hands the rasteriser (scaling_modifier * scales, canonical rotations) instead of
    the reference's precomputed covariance: the covariance the kernel builds from them is the same matrix."""
    g = torch.Generator().manual_seed(0)
    s, q, mod = torch.rand(50, 3, generator=g) * 0.1 + 0.01, torch.randn(50, 4, generator=g), 1.7
    cov6 = covariance_from_scaling_rotation(s, mod, q)                       # what get_covariance(mod) returns
    R = quat_to_rotmat(q)
    full = R @ torch.diag_embed((mod * s) ** 2) @ R.transpose(1, 2)          # what the kernel computes: R S^2 R^T
    want = torch.stack([full[:, 0, 0], full[:, 0, 1], full[:, 0, 2], full[:, 1, 1], full[:, 1, 2], full[:, 2, 2]], -1)
    assert torch.allclose(cov6, want, rtol=1e-5, atol=1e-9)
