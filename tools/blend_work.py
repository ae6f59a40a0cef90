#!/usr/bin/env python
"""Work statistics of the blend kernels on the bench scene: tile list lengths against the depth each tile is actually
walked to (largest n_contrib of its pixels).  gpurun -- python tools/blend_work.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"])
dev = torch.device("cuda:0")
r, sc = bench.build_model(wl, 0, dev)
from dimo_b200.camera import orbit_minicam
W, H = wl["W"], wl["H"]
frames = bench.step_schedule(wl, 0)
cams = [orbit_minicam(v, wl["views"], W, H, device=dev) for (_, v, _) in frames]
times = [f / wl["frames"] for (_, _, f) in frames]
lat = [m for (m, _, _) in frames]
r.gaussians.find_knn(4)
with torch.no_grad():
    out = r.render_batch(cams, times, lat, stage="s2", with_visibility=False, depth_normal=False)
st = out["raster_state"]
S = len(cams)
ranges = st.ranges.view(-1, 2).long()
length = (ranges[:, 1] - ranges[:, 0]).float()
gx, gy = (W + 15) // 16, (H + 15) // 16
nc = st.n_contrib.view(S, H, W).float()
pad = torch.zeros(S, gy * 16, gx * 16, device=dev); pad[:, :H, :W] = nc
walked = pad.view(S, gy, 16, gx, 16).amax(dim=(2, 4)).reshape(-1)
print(f"tiles {length.numel()}  instances R {int(length.sum())}  list length mean {length.mean():.1f} max {length.max():.0f}")
print(f"walked (max n_contrib per tile) mean {walked.mean():.1f}  -> fraction of the lists walked {walked.sum() / length.sum():.3f}")
print(f"per-pixel last contributor mean {nc.mean():.1f}; pixel-pairs walked by the tile {walked.sum() * 256 / 1e9:.3f} G, "
      f"pairs up to each pixel's own last contributor {nc.sum() / 1e9:.3f} G")
T = st.final_T.view(S, H, W)
print(f"final T: mean {T.mean():.4f}, saturated (<1e-4) fraction {(T < 1e-4).float().mean():.3f}")
