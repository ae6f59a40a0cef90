"""Host-side mirror of the reference `Renderer` (renderer/latent_gs_renderer.py:973-1293; `GaussianModel` lives in
gaussian_model.py) on top of the libdimo_b200 kernels.

`Renderer.render(...)` keeps the reference signature and result dict (one frame per call);
`Renderer.render_batch(...)` is the B200 fast path: all S frames of an optimisation step in ONE launch
set -- TimeNet once per unique (motion, t) pair (its output is view-independent), LBS once per pair,
rasterisation batched over all frames.  `initialize` / `initialize_ag` / `arap_loss_v2` / `reparameterize` complete
the class surface main_train_dimo.py and main_test_dimo.py use.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import deform as _deform
from . import flags as _flags
from . import raster as _raster
from . import regularisers as _reg
from .camera import MiniCam  # noqa: F401  (re-export: the reference imports MiniCam from the renderer module)
from .gaussian_model import (BasicPointCloud, GaussianModel, RGB2SH, SH2RGB, C0, get_expon_lr_func,  # noqa: F401
                             inverse_sigmoid)


def _ball_points(n, radius):
    """n points uniform in a ball, drawn from NumPy's global stream in the reference's order: phi, cos(theta), mu
    (renderer/latent_gs_renderer.py:999-1007)."""
    phis = np.random.random((n,)) * 2 * np.pi
    costheta = np.random.random((n,)) * 2 - 1
    thetas = np.arccos(costheta)
    mu = np.random.random((n,))
    r = radius * np.cbrt(mu)
    return np.stack((r * np.sin(thetas) * np.cos(phis), r * np.sin(thetas) * np.sin(phis), r * np.cos(thetas)), axis=1)


class Renderer:
    def __init__(self, sh_degree=3, white_background=True, radius=1, delta_t=1 / 32, num_latent_code=1,
                 latent_code_dim=32, add_normal=False, device="cuda", vae_latent=False):
        """vae_latent=True: the renderer/gaussian_gs_renderer.py twin (`_mu` / `_log_var` + reparameterisation)."""
        self.sh_degree = sh_degree
        self.white_background = white_background
        self.radius = radius
        self.gaussians = GaussianModel(sh_degree, num_latent_code, latent_code_dim, device=device,
                                       vae_latent=vae_latent)
        self.bg_color = torch.tensor([1, 1, 1] if white_background else [0, 0, 0], dtype=torch.float32, device=device)
        self.delta_t = delta_t
        self.add_normal = add_normal
        self.vae_latent = bool(vae_latent)

    # -- initialisation: :995-1058 -----------------------------------------------------------------
    def initialize(self, input=None, num_pts=5000, num_cpts=512, radius=0.5, radius2=0.5, only_init_gaussians=False,
                   dist3nn=None):
        """Random Gaussians in a ball of `radius` and control points in a ball of `radius2` (NumPy global stream, same
        draw order as the reference), or a given BasicPointCloud."""
        if input is None:
            xyz = _ball_points(num_pts, radius)
            shs = np.random.random((num_pts, 3)) / 255.0
            pcd = BasicPointCloud(points=xyz, colors=SH2RGB(shs), normals=np.zeros((num_pts, 3)))
            cxyz = _ball_points(num_cpts, radius2)
            cshs = np.random.random((num_cpts, 3)) / 255.0
            pcd2 = BasicPointCloud(points=cxyz, colors=SH2RGB(cshs), normals=np.zeros((num_cpts, 3)))
            self.gaussians.create_from_pcd(pcd, pcd2, 1, only_init_gaussians=only_init_gaussians, dist3nn=dist3nn)
        elif isinstance(input, BasicPointCloud):
            self.gaussians.create_from_pcd(input, input, 1, dist3nn=dist3nn)
        else:
            raise ValueError("Unsupported initialization type!!!")

    def initialize_ag(self, c_xyz, c_radius, num_cpts=512, num_pts_per_cpt=200, init_ratio=1, dist3nn=None):
        """Adaptive Gaussian initialisation (:1038-1058): the SAME ball of num_pts_per_cpt offsets, radius
        mean(c_radius) * init_ratio, around every control point."""
        offs = _ball_points(num_pts_per_cpt, c_radius.mean().item() * init_ratio)
        xyz = torch.tensor(offs)[None].repeat(num_cpts, 1, 1).flatten(0, 1)
        centres = c_xyz.cpu().data[:, None].repeat(1, num_pts_per_cpt, 1).flatten(0, 1)
        xyz = (xyz + centres).numpy()
        n = num_pts_per_cpt * num_cpts
        shs = np.random.random((n, 3)) / 255.0
        pcd = BasicPointCloud(points=xyz, colors=SH2RGB(shs), normals=np.zeros((n, 3)))
        self.gaussians.create_from_pcd(pcd, pcd, 1, only_init_gaussians=True, dist3nn=dist3nn)

    # -- latent codes ---------------------------------------------------------------------------------
    def reparameterize(self, mu, log_var):
        """gaussian_gs_renderer.py:1088-1098: z = mu + eps * exp(log_var / 2), eps ~ N(0, I)."""
        std = torch.exp(0.5 * log_var)
        return torch.randn_like(std) * std + mu

    def latent_code(self, latent_index):
        g = self.gaussians
        if self.vae_latent:
            return self.reparameterize(g._mu[latent_index], g._log_var[latent_index])
        return g._latent_codes[latent_index]

    # -- ARAP over T random time samples: :1081-1094 ----------------------------------------------------
    def arap_loss_v2(self, delta_t=0.05, t_samp_num=8, stage="s1", latent_index=0):
        g = self.gaussians
        q_times = torch.rand(t_samp_num).to(g._xyz.device)
        means3D = g._xyz if stage == "s1" else g._c_xyz                                  # [M,3]
        lat = self.latent_code(latent_index)
        deform, _ = g._timenet.forward_batched(means3D, q_times, lat[None, :].expand(t_samp_num, -1).contiguous())
        means3D_t = means3D[None].detach() + deform                                      # [T,M,3]
        return _reg.arap_loss_points(means3D_t)

    # ------------------------------------------------------------------------------------------
    def _camera_row(self, c):
        """35 host floats of one camera (view 16 | full projection 16 | centre 3), cached on the camera object: cameras
        are reused step after step, so the device->host read happens once per camera, not once per step."""
        row = getattr(c, "_dimo_row", None)
        if row is None:
            row = torch.cat([c.world_view_transform.reshape(-1).float().cpu(), c.full_proj_transform.reshape(-1).float().cpu(),
                             c.camera_center.reshape(-1).float().cpu()]).numpy().copy()
            try:
                c._dimo_row = row
            except Exception:
                pass
        return row

    def prepare_step(self, cameras, times, latent_indices, bg_color=None, out=None):
        """Host-side packing of one step's frame list: cams [S,40], t [U], li [U] i64, pf [S] i64, pf32 [S] i32 where
        U = unique (motion, t) pairs (the deformation is view-independent) and pf maps frame -> pair.
        Everything is laid out in ONE host staging buffer and reaches the device with one small copy (no kernels, no
        host sync); the returned tensors are views of the device buffer.  `out`: an earlier result whose buffer is
        overwritten in place (static buffers for CUDA-graph replay; U and S must not change)."""
        dev = self.gaussians._xyz.device
        pairs, pair_of_frame = {}, []
        for t, li in zip(times, latent_indices):
            pair_of_frame.append(pairs.setdefault((int(li), float(t)), len(pairs)))
        keys = list(pairs.keys())
        S, U = len(cameras), len(keys)
        bg = self.bg_color if bg_color is None else bg_color
        if bg is self.bg_color:
            cache = getattr(self, "_bg_host", None)
            if cache is None or cache[0] is not bg:
                cache = (bg, bg.detach().float().cpu().reshape(3).numpy().copy())
                self._bg_host = cache
            bg_host = cache[1]
        else:
            bg_host = bg.detach().float().cpu().reshape(3).numpy()
        o_t, o_li, o_pf, o_pf32, nbytes = self._prep_layout(S, U)
        # staging in PAGEABLE host memory on purpose: a copy this small (a few KB) from pageable memory travels inside the
        # command stream, while a pinned-memory copy is a DMA-engine job that queues behind the step's ground-truth
        # upload on the copy stream (measured on the bench's end-to-end arm: 4477 vs 4186 frames/s).  The source may be
        # reused as soon as copy_() returns.
        hb = getattr(self, "_prep_host", None)
        if hb is None or hb.numel() != nbytes:
            hb = self._prep_host = torch.zeros(nbytes, dtype=torch.uint8)
        hn = hb.numpy()
        cams_h = hn[:o_t].view(np.float32).reshape(S, 40)
        for i, c in enumerate(cameras):
            cams_h[i, :35] = self._camera_row(c)
            cams_h[i, 35] = math.tan(c.FoVx * 0.5); cams_h[i, 36] = math.tan(c.FoVy * 0.5)
            cams_h[i, 37:40] = bg_host
        hn[o_t:o_t + U * 4].view(np.float32)[:] = [k_[1] for k_ in keys]
        hn[o_li:o_pf].view(np.int64)[:] = [k_[0] for k_ in keys]
        hn[o_pf:o_pf32].view(np.int64)[:] = pair_of_frame
        hn[o_pf32:nbytes].view(np.int32)[:] = pair_of_frame
        if out is not None:
            assert out["S"] == S and out["U"] == U and out["expand"] == (pair_of_frame != list(range(S)))
            buf = out["_buf"]
        else:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        buf.copy_(hb, non_blocking=True)
        if out is not None:
            out["pair_of_frame"] = pair_of_frame
            return out
        prep = {"S": S, "U": U, "W": int(cameras[0].image_width), "H": int(cameras[0].image_height),
                "pair_of_frame": pair_of_frame, "expand": pair_of_frame != list(range(S))}
        return self._prep_views(prep, buf)

    @staticmethod
    def _prep_layout(S, U):
        """byte layout (8-byte aligned sections): cams f32 [S,40] | t f32 [U] (+pad) | li i64 [U] | pf i64 [S] | pf32 i32 [S]"""
        o_t = S * 40 * 4
        o_li = o_t + (U * 4 + 7) // 8 * 8
        o_pf = o_li + U * 8
        o_pf32 = o_pf + S * 8
        return o_t, o_li, o_pf, o_pf32, o_pf32 + S * 4

    @classmethod
    def _prep_views(cls, prep, buf):
        S, U = prep["S"], prep["U"]
        o_t, o_li, o_pf, o_pf32, nbytes = cls._prep_layout(S, U)
        prep.update({"_buf": buf, "cams": buf[:o_t].view(torch.float32).view(S, 40),
                     "t": buf[o_t:o_t + U * 4].view(torch.float32), "li": buf[o_li:o_pf].view(torch.int64),
                     "pf": buf[o_pf:o_pf32].view(torch.int64), "pf32": buf[o_pf32:nbytes].view(torch.int32)})
        return prep

    @classmethod
    def clone_prep(cls, prep):
        """Deep copy of a prepare_step() result (own device buffer): the static inputs of a captured graph."""
        return cls._prep_views({k: v for k, v in prep.items() if not torch.is_tensor(v)}, prep["_buf"].clone())

    def render_batch(self, cameras=None, times=None, latent_indices=None, stage="s2", scaling_modifier=1.0,
                     bg_color=None, override_color=None, xyz_detach=False, clamp=True, prepared=None, capacity=None,
                     with_visibility=True, depth_normal=True, with_cpts=True):
        """All S frames of a step in ONE launch set.  cameras: list of S MiniCam (same W,H); times: list of S floats;
        latent_indices: list of S ints -- or `prepared` = the result of prepare_step().  capacity: instance-slot
        capacity for the sync-free rasteriser mode (None = exact mode with one host read-back).  depth_normal=False:
        depth / normal are not rendered (None in the result) -- for steps whose loss reads image + alpha only.
        Returns a dict of batched tensors: image [S,3,H,W] (clamped), image_raw, depth, normal, alpha, radii [S,N],
        visibility_filter, pts_t [U,N,3] (one block per unique (motion,t); frame f uses block pair_of_frame[f]),
        cpts_t [U,M,3] + `pair_of_frame`."""
        g = self.gaussians
        prep = prepared if prepared is not None else self.prepare_step(cameras, times, latent_indices, bg_color)
        W, H = prep["W"], prep["H"]
        # direct_grads (set by trainstep.TrainStep): backward kernels add parameter gradients straight into the flat
        # gradient buffer's views (_lib.grad_sink) and tell the reducer, instead of zero-filled temporaries + one
        # AccumulateGrad add per parameter (~20 small launches per step)
        direct = bool(getattr(self, "direct_grads", False)) and torch.is_grad_enabled()
        reducer = getattr(g, "reducer", None)
        done = (reducer.direct_written if reducer is not None else (lambda ps: None)) if direct else None
        if self.vae_latent:                                                  # one reparameterised draw per pair
            latents = self.reparameterize(g._mu.index_select(0, prep["li"]), g._log_var.index_select(0, prep["li"]))
        elif direct:
            latents = _deform.gather_rows(g._latent_codes, prep["li"], lambda: done([g._latent_codes]))
        else:
            latents = g._latent_codes.index_select(0, prep["li"])           # [U,L]
        t_dev = prep["t"]
        if stage >= "s2":
            dxyz, dquat = g._timenet.forward_batched(g._c_xyz, t_dev, latents, pts_direct=done)   # [U,M,3],[U,M,4]
            cpts_t = g._c_xyz[None] + dxyz if with_cpts else None
            means3D_u, rot_u = _deform.lbs_deform(g._xyz, g._rotation, g._c_xyz, g._c_radius, dxyz, dquat,
                                                  g.neighbor_indices, g.neighbor_dists, direct_grads=direct, on_done=done)
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
        # exp / sigmoid of the raw _scaling / _opacity run inside the projection kernels (A4 folded in); while the shared
        # radius `_r` of stage s1 is alive the scales come from it (get_scaling, :340-350) through torch
        own_scales = len(g._r) == 0
        color, depth, normal, alpha, radii = _raster.rasterize_batch(
            prep["cams"], means3D, g._scaling if own_scales else g.get_scaling, rotations, g._opacity, W, H, shs=shs,
            colors_precomp=colors, sh_degree=g.active_sh_degree, scale_modifier=scaling_modifier, state_out=state,
            capacity=capacity, frame_src=frame_src, depth_normal=depth_normal, raw_activations=(1 if own_scales else 0) | 2,
            direct_grads=direct, on_done=done)
        return {"image": color.clamp(0, 1) if clamp else None, "image_raw": color, "depth": depth, "normal": normal,
                "alpha": alpha, "radii": radii, "visibility_filter": (radii > 0) if with_visibility else None,
                "pts_t": means3D, "cpts_t": cpts_t, "pair_of_frame": prep["pair_of_frame"], "raster_state": state[0]}

    # ------------------------------------------------------------------------------------------
    def render(self, viewpoint_camera, scaling_modifier=1.0, bg_color=None, override_color=None,
               compute_cov3D_python=False, convert_SHs_python=False, time=0.0, stage="s1", rot_as_res=True,
               xyz_detach=False, local_frame=True, direct_deform=False, vertices_deform=None, latent_index=0):
        """Reference signature and result dict (renderer/latent_gs_renderer.py:1096-1293), one frame."""
        if compute_cov3D_python and stage >= "s2":
            # the reference itself cannot do this: it leaves rotations = None and then calls quat_mul(rots3D, None) (:1209)
            raise ValueError("compute_cov3D_python=True is only defined for stage 's1' (as in the reference)")
        g = self.gaussians
        dev = g._xyz.device
        screenspace_points = torch.zeros_like(g._xyz, requires_grad=True) + 0
        try:
            screenspace_points.retain_grad()
        except Exception:
            pass
        t_dev = torch.tensor([float(time)], dtype=torch.float32).to(dev, non_blocking=True)
        latents = self.latent_code(latent_index)[None]
        if stage >= "s2":
            dxyz, dquat = g._timenet.forward_batched(g._c_xyz, t_dev, latents)
            cpts_t = g._c_xyz + dxyz[0]
            if local_frame:
                means3D, rotations = _deform.lbs_deform(g._xyz, g._rotation, g._c_xyz, g._c_radius, dxyz[0], dquat[0],
                                                        g.neighbor_indices, g.neighbor_dists)
            else:                                          # off-default flag: tensor expressions (flags.py)
                means3D, rotations = _flags.lbs_global_frame(g._xyz, g._rotation, g.get_c_radius(stage), dxyz[0],
                                                             dquat[0], g.neighbor_indices, g.neighbor_dists)
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
            if convert_SHs_python:                         # off-default flag (:1227-1238): colours from the CANONICAL centres
                colors = _flags.sh_colors(g.active_sh_degree, g.get_features, g.get_xyz, viewpoint_camera.camera_center)
            else:
                shs = g.get_features
        else:
            colors = override_color
        # compute_cov3D_python (stage s1 only): the reference's precomputed covariance R (m s)^2 R^T from the canonical
        # rotations is what the kernel builds from (scales, rotations, scale_modifier) -- in s1 `rotations` ARE the
        # canonical ones, so the flag needs no separate path
        color, depth, normal, alpha, radii = _raster.rasterize_batch(
            cams, means3D, g.get_scaling, rotations, g.get_opacity, int(viewpoint_camera.image_width),
            int(viewpoint_camera.image_height), shs=shs, colors_precomp=colors, sh_degree=g.active_sh_degree,
            scale_modifier=scaling_modifier, means2D=screenspace_points)
        radii = radii[0]
        return {"image": color[0].clamp(0, 1), "depth": depth[0], "normal": normal[0], "alpha": alpha[0],
                "viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii,
                "pts_t": means3D, "cpts_t": cpts_t}
