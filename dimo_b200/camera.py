"""Host-side cameras with the reference's attribute names (MiniCam, renderer/latent_gs_renderer.py:943-970;
orbit_camera, utils/cam_utils.py:41-58).  NumPy on the host exactly like the reference; the only
change is that the target device is a parameter instead of a hard-coded .cuda()."""
import math

import numpy as np
import torch


def _safe_normalize(x, eps=1e-20):
    return x / np.sqrt(np.maximum(np.sum(x * x, axis=-1, keepdims=True), eps))


def look_at(campos, target, opengl=True):
    if not opengl:
        forward = _safe_normalize(target - campos)
        up = np.array([0, 1, 0], dtype=np.float32)
        right = _safe_normalize(np.cross(forward, up))
        up = _safe_normalize(np.cross(right, forward))
    else:
        forward = _safe_normalize(campos - target)
        up = np.array([0, 1, 0], dtype=np.float32)
        right = _safe_normalize(np.cross(up, forward))
        up = _safe_normalize(np.cross(forward, right))
    return np.stack([right, up, forward], axis=1)


def orbit_camera(elevation, azimuth, radius=1, is_degree=True, target=None, opengl=True):
    if is_degree:
        elevation = np.deg2rad(elevation)
        azimuth = np.deg2rad(azimuth)
    x = radius * np.cos(elevation) * np.sin(azimuth)
    y = -radius * np.sin(elevation)
    z = radius * np.cos(elevation) * np.cos(azimuth)
    if target is None:
        target = np.zeros([3], dtype=np.float32)
    campos = np.array([x, y, z]) + target
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = look_at(campos, target, opengl)
    T[:3, 3] = campos
    return T


def getProjectionMatrix(znear, zfar, fovX, fovY):
    P = torch.zeros(4, 4)
    P[0, 0] = 1 / math.tan(fovX / 2)
    P[1, 1] = 1 / math.tan(fovY / 2)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


class MiniCam:
    def __init__(self, c2w, width, height, fovy, fovx, znear, zfar, device="cuda"):
        self.image_width = width
        self.image_height = height
        self.FoVy = fovy
        self.FoVx = fovx
        self.znear = znear
        self.zfar = zfar
        w2c = np.linalg.inv(c2w)
        w2c[1:3, :3] *= -1
        w2c[:3, 3] *= -1
        self.world_view_transform = torch.tensor(w2c).transpose(0, 1).to(device)
        self.projection_matrix = getProjectionMatrix(znear, zfar, fovx, fovy).transpose(0, 1).to(device)
        self.full_proj_transform = self.world_view_transform @ self.projection_matrix
        self.camera_center = -torch.tensor(c2w[:3, 3]).to(device)


def orbit_minicam(view, num_views, width, height, fovy_deg=33.9, radius=2.0, elevation=0.0, znear=0.01,
                  zfar=100.0, device="cuda"):
    """Synthetic-benchmark camera recipe: orbit_camera(0, 360 v/V, 2) with fovy 33.9 deg."""
    fovy = np.deg2rad(fovy_deg)
    fovx = 2 * np.arctan(np.tan(fovy / 2) * width / height)
    return MiniCam(orbit_camera(elevation, 360.0 * view / num_views, radius), width, height, float(fovy),
                   float(fovx), znear, zfar, device=device)
