import torch

__all__ = ["quaternion_to_matrix"]


def quaternion_to_matrix(quaternions: torch.Tensor) -> torch.Tensor:
    """(w, x, y, z) -> rotation matrices, scaled by 2 / |q|^2 like pytorch3d (no prior normalisation needed);
    used by GaussianModel.get_rotation_matrix, renderer/latent_gs_renderer.py:385-386."""
    w, x, y, z = torch.unbind(quaternions, -1)
    s = 2.0 / (quaternions * quaternions).sum(-1)
    m = torch.stack((1 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
                     s * (x * y + z * w), 1 - s * (x * x + z * z), s * (y * z - x * w),
                     s * (x * z - y * w), s * (y * z + x * w), 1 - s * (x * x + y * y)), -1)
    return m.reshape(quaternions.shape[:-1] + (3, 3))
