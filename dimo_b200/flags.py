"""The non-default flags of `Renderer.render` (renderer/latent_gs_renderer.py:1096-1111) as plain tensor expressions.
Neither driver sets them (main_train_dimo.py, main_test_dimo.py call render() with time / stage / latent_index only),
so they stay off the measured path: a few torch launches in front of the same rasteriser kernels.

  convert_SHs_python=True    colours = max(SH(active degree, normalize(xyz_canonical - camera)) + 0.5, 0) handed to the
                             rasteriser as colors_precomp (:1227-1238; note: CANONICAL centres, not the deformed ones)
  local_frame=False          skinning without the per-control-point rotation of the offset: x' = x + sum_k w_k dx_jk
                             (:1205-1206); the rotation blend is unchanged
  compute_cov3D_python=True  the reference precomputes the 3-D covariance from the CANONICAL rotations (:1184-1185,
                             get_covariance :409-410), i.e. the deformation's rotation is ignored; the same covariance
                             is reproduced by handing the rasteriser (scales, canonical rotations), see Renderer.render
"""
import torch
import torch.nn.functional as F

# real spherical-harmonics basis up to degree 3 in the order and sign convention of utils/sh_utils.py:26-55
_C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
       1.445305721320277, -0.5900435899266435)


def sh_basis(deg, dirs):
    """dirs [N,3] (unit) -> basis values [N,(deg+1)^2]."""
    if not 0 <= deg <= 3:
        raise ValueError("SH degree must be 0..3")
    x, y, z = dirs.unbind(-1)
    cols = [torch.full_like(x, _C0)]
    if deg >= 1:
        cols += [-_C1 * y, _C1 * z, -_C1 * x]
    if deg >= 2:
        xx, yy, zz = x * x, y * y, z * z
        cols += [_C2[0] * x * y, _C2[1] * y * z, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * x * z, _C2[4] * (xx - yy)]
    if deg >= 3:
        cols += [_C3[0] * y * (3 * xx - yy), _C3[1] * x * y * z, _C3[2] * y * (4 * zz - xx - yy),
                 _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy), _C3[5] * z * (xx - yy),
                 _C3[6] * x * (xx - 3 * yy)]
    return torch.stack(cols, dim=-1)


def sh_colors(deg, features, xyz, camera_center):
    """features [N,K,3] (get_features layout), xyz [N,3], camera_center [3] -> rgb [N,3] = clamp_min(SH + 0.5, 0)."""
    d = xyz - camera_center.reshape(1, 3)
    d = d / d.norm(dim=1, keepdim=True)
    basis = sh_basis(deg, d)                                               # [N,B]
    rgb = torch.einsum("nb,nbc->nc", basis, features[:, :basis.shape[1], :])
    return torch.clamp_min(rgb + 0.5, 0.0)


def quat_mul(q1, q2):
    """Hamilton product, (w, x, y, z) (:135-147)."""
    w1, x1, y1, z1 = q1.unbind(-1)
    w2, x2, y2, z2 = q2.unbind(-1)
    return torch.stack((w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                        w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2), dim=-1)


def lbs_global_frame(xyz, rot, c_radius, dxyz, dquat, neighbor_indices, neighbor_dists, eps=1e-7):
    """local_frame=False (:1193-1209): xyz [N,3], rot [N,4] raw, c_radius [M,1] ACTIVATED (exp), dxyz [M,3], dquat [M,4],
    neighbor_indices [N,K] int64, neighbor_dists [N,K] -> (means3D [N,3], rotations [N,4] normalised)."""
    r = c_radius[neighbor_indices][:, :, 0]
    w = torch.exp(-1.0 * neighbor_dists ** 2 / (2.0 * r ** 2)) + eps
    w = F.normalize(w, p=1)
    means = xyz + (w[..., None] * dxyz[neighbor_indices]).sum(dim=1)
    blend = (w[..., None] * dquat[neighbor_indices]).sum(dim=1)
    return means, F.normalize(quat_mul(blend, rot))
