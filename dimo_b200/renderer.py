"""Host-side mirror of the reference renderer classes for the hot path
(renderer/latent_gs_renderer.py: GaussianModel :248-415 [parameters + activations only], Renderer.render
:1096-1293) on top of the libdimo_b200 kernels.

`Renderer.render(...)` keeps the reference signature and result dict (one frame per call);
`Renderer.render_batch(...)` is the B200 fast path: all S frames of an optimisation step in ONE launch
set -- TimeNet once per unique (motion, t) pair (its output is view-independent), LBS once per pair,
rasterisation batched over all frames.

Out of scope here: densify/prune, PLY/.pth I/O, optimizer surgery.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import deform as _deform
from . import knn as _knn
from . import raster as _raster
from .camera import MiniCam  # noqa: F401  (re-export: the reference imports MiniCam from the renderer module)

C0 = 0.28209479177387814


def inverse_sigmoid(x):
    return torch.log(x / (1 - x))


class GaussianModel:
    """Parameters + activations of the reference GaussianModel, same attribute names."""

    def __init__(self, sh_degree: int, num_latent_code: int = 1, latent_code_dim: int = 32, device="cuda"):
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        self.num_latent_code = num_latent_code
        self.latent_code_dim = latent_code_dim
        self.device = device
        e = torch.empty(0, device=device)
        self._xyz = self._features_dc = self._features_rest = self._scaling = self._rotation = self._opacity = e
        self._c_xyz = self._c_radius = self._r = e
        self.max_radii2D = e
        self.optimizer = None
        self._latent_codes = nn.Parameter(torch.randn(num_latent_code, latent_code_dim, device=device))
        self._timenet = _deform.TimeNet(latent_code_dim=latent_code_dim).to(device)
        self.neighbor_dists = None
        self.neighbor_indices = None

    # -- construction from tensors (synthetic scenes, checkpoints) --------------------------------
    def load_state(self, state: dict):
        for k in ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity", "_c_xyz", "_c_radius"):
            setattr(self, k, nn.Parameter(state[k].to(self.device).float().contiguous()))
        if "_latent_codes" in state:
            self._latent_codes = nn.Parameter(state["_latent_codes"].to(self.device).float().contiguous())
            self.num_latent_code = self._latent_codes.shape[0]
        self._r = torch.empty(0, device=self.device)
        self.max_radii2D = torch.zeros(self._xyz.shape[0], device=self.device)

    # -- activations: renderer/latent_gs_renderer.py:257-265, 340-407 ---------------------------
    @property
    def get_scaling(self):
        if len(self._r) == 0:
            return torch.exp(self._scaling)
        elif self._r.shape[0] != self._xyz.shape[0]:
            return torch.exp(self._r.repeat(self._xyz.shape[0], 3))
        elif self._r.shape[1] == 1:
            return torch.exp(self._r.repeat(1, 3))
        elif self._r.shape == self._xyz.shape:
            return torch.exp(self._r)
        raise ValueError("Shape of _r is not supported.")

    @property
    def get_rotation(self):
        return F.normalize(self._rotation)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_c_xyz(self):
        return self._c_xyz

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return torch.sigmoid(self._opacity)

    @property
    def get_latent_codes(self):
        return self._latent_codes

    def get_c_radius(self, stage="s2"):
        if stage < "s2":
            return torch.exp(self._r.repeat(self._xyz.shape[0], 1))
        return torch.exp(self._c_radius)

    def parameters(self):
        return [self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling, self._rotation,
                self._c_xyz, self._c_radius, self._latent_codes] + list(self._timenet.parameters())

    # reference group names and order: GaussianModel.training_setup, renderer/latent_gs_renderer.py:460-473
    GROUP_NAMES = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation", "latent_code", "deform", "deform_rot",
                   "c_xyz", "c_radius", "r")

    def param_groups(self, lr=0.0):
        """The twelve optimizer groups of training_setup.  lr: float (every group) or {name: lr} (missing -> 0.0,
        the reference constructs Adam with lr=0.0 and per-group values)."""
        mlp, mlp_rot = self._timenet.get_mlp_parameters()
        tensors = {"xyz": [self._xyz], "f_dc": [self._features_dc], "f_rest": [self._features_rest],
                   "opacity": [self._opacity], "scaling": [self._scaling], "rotation": [self._rotation],
                   "latent_code": [self._latent_codes], "deform": list(mlp), "deform_rot": list(mlp_rot),
                   "c_xyz": [self._c_xyz], "c_radius": [self._c_radius], "r": [self._r]}
        get = (lambda n: float(lr.get(n, 0.0))) if isinstance(lr, dict) else (lambda n: float(lr))
        return [{"params": [p for p in tensors[n] if isinstance(p, nn.Parameter)], "lr": get(n), "name": n}
                for n in self.GROUP_NAMES]

    def find_knn(self, k=4):
        """main_train_dimo.py:502-509 (GUI.find_knn): once per optimisation step."""
        self.neighbor_dists, self.neighbor_indices = _knn.knn(self._c_xyz, self._xyz, k)


class Renderer:
    def __init__(self, sh_degree=3, white_background=True, radius=1, delta_t=1 / 32, num_latent_code=1,
                 latent_code_dim=32, add_normal=False, device="cuda"):
        self.sh_degree = sh_degree
        self.white_background = white_background
        self.radius = radius
        self.gaussians = GaussianModel(sh_degree, num_latent_code, latent_code_dim, device=device)
        self.bg_color = torch.tensor([1, 1, 1] if white_background else [0, 0, 0], dtype=torch.float32, device=device)
        self.delta_t = delta_t
        self.add_normal = add_normal

    # ------------------------------------------------------------------------------------------
    def prepare_step(self, cameras, times, latent_indices, bg_color=None, out=None):
        """Host-side packing of one step's frame list into device tensors (no kernels of ours, no host sync):
        cams [S,40], t [U], li [U] i64, pf [S] i64 where U = unique (motion, t) pairs (the deformation is
        view-independent) and pf maps frame -> pair.  `out`: an earlier result whose tensors are
        overwritten in place (static buffers for CUDA-graph replay; U and S must not change)."""
        dev = self.gaussians._xyz.device
        pairs, pair_of_frame = {}, []
        for t, li in zip(times, latent_indices):
            pair_of_frame.append(pairs.setdefault((int(li), float(t)), len(pairs)))
        keys = list(pairs.keys())
        bg = self.bg_color if bg_color is None else bg_color
        V = torch.stack([c.world_view_transform for c in cameras]).reshape(-1, 16).float()
        P = torch.stack([c.full_proj_transform for c in cameras]).reshape(-1, 16).float()
        C = torch.stack([c.camera_center for c in cameras]).float()
        S = len(cameras)
        host = torch.empty(S, 3, dtype=torch.float32)                  # tanfovx, tanfovy, pair index
        for i, c in enumerate(cameras):
            host[i, 0] = math.tan(c.FoVx * 0.5); host[i, 1] = math.tan(c.FoVy * 0.5)
        host[:, 2] = torch.tensor(pair_of_frame, dtype=torch.float32)
        small = host.to(dev, non_blocking=True)
        bgd = bg.to(dev).float().reshape(1, 3).expand(S, 3)
        cams = torch.cat([V, P, C, small[:, 0:2], bgd], dim=1)
        th = torch.tensor([[k[1], float(k[0])] for k in keys], dtype=torch.float32).to(dev, non_blocking=True)
        prep = {"cams": cams, "t": th[:, 0].contiguous(), "li": th[:, 1].long(), "pf": small[:, 2].long(),
                "pf32": small[:, 2].int(),
                "S": S, "U": len(keys), "W": int(cameras[0].image_width), "H": int(cameras[0].image_height),
                "pair_of_frame": pair_of_frame, "expand": pair_of_frame != list(range(S))}
        if out is not None:
            assert out["S"] == prep["S"] and out["U"] == prep["U"] and out["expand"] == prep["expand"]
            for k in ("cams", "t", "li", "pf", "pf32"):
                out[k].copy_(prep[k])
            out["pair_of_frame"] = pair_of_frame
            return out
        return prep

    def render_batch(self, cameras=None, times=None, latent_indices=None, stage="s2", scaling_modifier=1.0,
                     bg_color=None, override_color=None, xyz_detach=False, clamp=True, prepared=None, capacity=None,
                     with_visibility=True):
        """All S frames of a step in ONE launch set.  cameras: list of S MiniCam (same W,H); times: list of S floats;
        latent_indices: list of S ints -- or `prepared` = the result of prepare_step().  capacity: instance-slot
        capacity for the sync-free rasteriser mode (None = exact mode with one host read-back).
        Returns a dict of batched tensors: image [S,3,H,W] (clamped), image_raw, depth, normal, alpha, radii [S,N],
        visibility_filter, pts_t [U,N,3] (one block per unique (motion,t); frame f uses block pair_of_frame[f]),
        cpts_t [U,M,3] + `pair_of_frame`."""
        g = self.gaussians
        prep = prepared if prepared is not None else self.prepare_step(cameras, times, latent_indices, bg_color)
        W, H = prep["W"], prep["H"]
        latents = g._latent_codes.index_select(0, prep["li"])               # [U,L]
        t_dev = prep["t"]
        if stage >= "s2":
            dxyz, dquat = g._timenet.forward_batched(g._c_xyz, t_dev, latents)   # [U,M,3],[U,M,4]
            cpts_t = g._c_xyz[None] + dxyz
            means3D_u, rot_u = _deform.lbs_deform(g._xyz, g._rotation, g._c_xyz, g._c_radius, dxyz, dquat,
                                                  g.neighbor_indices, g.neighbor_dists)
        elif stage == "s1":
            dxyz, dquat = g._timenet.forward_batched(g._xyz, t_dev, latents)
            cpts_t = g._xyz[None] + dxyz
            means3D_u = cpts_t
            rot_u = F.normalize(g._rotation)[None].expand(prep["U"], -1, -1)
        else:
            raise ValueError("Nonexistent stage!!!")
        if xyz_detach:
            means3D_u = means3D_u.detach()
        # (motion, t) pair -> frames: the rasteriser reads block pf[b] for frame b (no [S,N,*] copies) and its
        # backward folds the per-frame gradients back onto the U blocks with one segment-sum launch per tensor
        frame_src = prep["pf32"] if prep["expand"] else None
        means3D, rotations = means3D_u, rot_u

        shs = colors = None
        if override_color is None:
            # no higher-order coefficients (sh_degree 0): the DC tensor IS the feature tensor, skip the concatenation
            shs = g._features_dc if g._features_rest.shape[1] == 0 else g.get_features
        else:
            colors = override_color
        state = []
        color, depth, normal, alpha, radii = _raster.rasterize_batch(
            prep["cams"], means3D, g.get_scaling, rotations, g.get_opacity, W, H, shs=shs, colors_precomp=colors,
            sh_degree=g.active_sh_degree, scale_modifier=scaling_modifier, state_out=state, capacity=capacity,
            frame_src=frame_src)
        return {"image": color.clamp(0, 1) if clamp else None, "image_raw": color, "depth": depth, "normal": normal,
                "alpha": alpha, "radii": radii, "visibility_filter": (radii > 0) if with_visibility else None,
                "pts_t": means3D, "cpts_t": cpts_t, "pair_of_frame": prep["pair_of_frame"], "raster_state": state[0]}

    # ------------------------------------------------------------------------------------------
    def render(self, viewpoint_camera, scaling_modifier=1.0, bg_color=None, override_color=None,
               compute_cov3D_python=False, convert_SHs_python=False, time=0.0, stage="s1", rot_as_res=True,
               xyz_detach=False, local_frame=True, direct_deform=False, vertices_deform=None, latent_index=0):
        """Reference signature and result dict (renderer/latent_gs_renderer.py:1096-1293), one frame."""
        if compute_cov3D_python or convert_SHs_python or not local_frame:
            raise NotImplementedError("dimo_b200 render(): python-side cov3D / SH conversion and local_frame=False "
                                      "are off the reference's default path")
        g = self.gaussians
        dev = g._xyz.device
        screenspace_points = torch.zeros_like(g._xyz, requires_grad=True) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
        t_dev = torch.tensor([float(time)], dtype=torch.float32).to(dev, non_blocking=True)
        latents = g._latent_codes[latent_index][None]
        if stage >= "s2":
            dxyz, dquat = g._timenet.forward_batched(g._c_xyz, t_dev, latents)
            cpts_t = g._c_xyz + dxyz[0]
            means3D, rotations = _deform.lbs_deform(g._xyz, g._rotation, g._c_xyz, g._c_radius, dxyz[0], dquat[0],
                                                    g.neighbor_indices, g.neighbor_dists)
        elif stage == "s1":
            dxyz, dquat = g._timenet.forward_batched(g._xyz, t_dev, latents)
            cpts_t = g._xyz + dxyz[0]
            means3D = cpts_t
            rotations = F.normalize(g._rotation)
        else:
            raise ValueError("Nonexistent stage!!!")
        if xyz_detach:
            means3D = means3D.detach()
        bg = self.bg_color if bg_color is None else bg_color
        cams = _raster.pack_cameras(viewpoint_camera.world_view_transform, viewpoint_camera.full_proj_transform,
                                    viewpoint_camera.camera_center, math.tan(viewpoint_camera.FoVx * 0.5),
                                    math.tan(viewpoint_camera.FoVy * 0.5), bg)
        shs = colors = None
        if override_color is None:
            shs = g.get_features
        else:
            colors = override_color
        color, depth, normal, alpha, radii = _raster.rasterize_batch(
            cams, means3D, g.get_scaling, rotations, g.get_opacity, int(viewpoint_camera.image_width),
            int(viewpoint_camera.image_height), shs=shs, colors_precomp=colors, sh_degree=g.active_sh_degree,
            scale_modifier=scaling_modifier, means2D=screenspace_points)
        radii = radii[0]
        return {"image": color[0].clamp(0, 1), "depth": depth[0], "normal": normal[0], "alpha": alpha[0],
                "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
                "pts_t": means3D, "cpts_t": cpts_t}
