"""Seeded synthetic inputs of the reference's shapes.  Used by bench.py and the tests;
everything is generated on the host (NumPy/SciPy) so the CPU oracle and the CUDA path see identical bits."""
import math

import numpy as np
import torch

C0 = 0.28209479177387814


def _ball(rng, n, radius):
    # renderer/latent_gs_renderer.py:999-1007
    phis = rng.random(n) * 2 * np.pi
    costheta = rng.random(n) * 2 - 1
    thetas = np.arccos(costheta)
    r = radius * np.cbrt(rng.random(n))
    return np.stack([r * np.sin(thetas) * np.cos(phis), r * np.sin(thetas) * np.sin(phis), r * np.cos(thetas)], 1)


def _mean_sq_dist_3nn(pts):
    if len(pts) < 4:
        return np.full((len(pts),), 1e-3)
    from scipy.spatial import cKDTree
    d, _ = cKDTree(pts).query(pts, k=4)
    return (d[:, 1:] ** 2).mean(axis=1)


def make_scene(n_gaussians, n_ctrl=512, n_motions=1, latent_dim=32, sh_coeffs=1, seed=0):
    """Returns a dict of fp32 CPU tensors with the reference's parameter names/shapes."""
    rng = np.random.default_rng(seed)
    xyz = _ball(rng, n_gaussians, 0.5)
    d2 = np.maximum(_mean_sq_dist_3nn(xyz), 1e-7)
    scaling = np.log(np.sqrt(d2))[:, None].repeat(3, 1) + rng.normal(0, 0.1, (n_gaussians, 3))
    rotation = rng.normal(0, 1, (n_gaussians, 4))
    opacity = rng.uniform(0.05, 0.95, (n_gaussians, 1))
    opacity = np.log(opacity / (1 - opacity))
    f_dc = (rng.random((n_gaussians, 1, 3)) - 0.5) / C0
    f_rest = rng.normal(0, 0.05, (n_gaussians, sh_coeffs - 1, 3))
    perm = rng.permutation(n_gaussians)[:n_ctrl]
    c_xyz = xyz[perm]
    m = min(n_ctrl, n_gaussians)
    if m >= 4:
        c_d2 = _mean_sq_dist_3nn(c_xyz)
    else:
        c_d2 = np.full((m,), 0.01)
    c_radius = np.log(np.sqrt(np.maximum(c_d2, 1e-7)))[:, None]
    latents = rng.normal(0, 1, (n_motions, latent_dim))
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32)
    return dict(_xyz=t(xyz), _scaling=t(scaling), _rotation=t(rotation), _opacity=t(opacity),
                _features_dc=t(f_dc), _features_rest=t(f_rest), _c_xyz=t(c_xyz), _c_radius=t(c_radius),
                _latent_codes=t(latents))
