"""Oracle: orbit cameras and the MiniCam matrices (host-side NumPy, as in the reference).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates utils/cam_utils.py:21-58
(look_at, orbit_camera) and renderer/latent_gs_renderer.py:927-970 (getProjectionMatrix,
MiniCam) without the hard-coded .cuda() calls.
"""
import math
import numpy as np
import torch


def _safe_normalize(x, eps=1e-20):
    # utils/cam_utils.py:6-19
    return x / np.sqrt(np.maximum(np.sum(x * x, axis=-1, keepdims=True), eps))


def look_at(campos, target):
    # utils/cam_utils.py:21-38, opengl=True branch
    forward = _safe_normalize(campos - target)
    up = np.array([0, 1, 0], dtype=np.float32)
    right = _safe_normalize(np.cross(up, forward))
    up = _safe_normalize(np.cross(forward, right))
    return np.stack([right, up, forward], axis=1)


def orbit_camera(elevation, azimuth, radius=1.0):
    # utils/cam_utils.py:41-58 (is_degree=True, target=None, opengl=True)
    elevation = np.deg2rad(elevation)
    azimuth = np.deg2rad(azimuth)
    x = radius * np.cos(elevation) * np.sin(azimuth)
    y = -radius * np.sin(elevation)
    z = radius * np.cos(elevation) * np.cos(azimuth)
    target = np.zeros([3], dtype=np.float32)
    campos = np.array([x, y, z]) + target
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = look_at(campos, target)
    T[:3, 3] = campos
    return T


def projection_matrix(znear, zfar, fovx, fovy):
    # renderer/latent_gs_renderer.py:927-940
    P = torch.zeros(4, 4)
    P[0, 0] = 1 / math.tan(fovx / 2)
    P[1, 1] = 1 / math.tan(fovy / 2)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class Camera:
    """CPU twin of MiniCam, renderer/latent_gs_renderer.py:943-970 (same attribute names)."""

    def __init__(self, c2w, width, height, fovy, fovx, znear, zfar):
        self.image_width = width
        self.image_height = height
        self.FoVy = fovy
        self.FoVx = fovx
        self.znear = znear
        self.zfar = zfar
        w2c = np.linalg.inv(c2w)
        w2c[1:3, :3] *= -1
        w2c[:3, 3] *= -1
        self.world_view_transform = torch.tensor(w2c).transpose(0, 1).contiguous()
        self.projection_matrix = projection_matrix(znear, zfar, fovx, fovy).transpose(0, 1).contiguous()
        self.full_proj_transform = self.world_view_transform @ self.projection_matrix
        self.camera_center = -torch.tensor(c2w[:3, 3])      # (sic) :970

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)      # renderer/latent_gs_renderer.py:1129

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)


def orbit_cam(view, num_views, width, height, fovy_deg=33.9, radius=2.0, elevation=0.0,
              znear=0.01, zfar=100.0):
    """SURVEY.md §8d camera recipe: orbit_camera(0, 360 v/V, 2) -> MiniCam with fovy 33.9 deg
    (configs/train_config.yaml:26-28, utils/cam_utils.py:62,73-75)."""
    fovy = np.deg2rad(fovy_deg)
    fovx = 2 * np.arctan(np.tan(fovy / 2) * width / height)
    pose = orbit_camera(elevation, 360.0 * view / num_views, radius)
    return Camera(pose, width, height, float(fovy), float(fovx), znear, zfar)
