"""GPU micro-benchmark of the rasteriser front end (preprocess + depth sort, counting sort) vs the number of frames.
gpurun -- python tools/sort_bench.py"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gpu_parity as gp
from dimo_b200 import _lib, raster as draster
from dimo_b200.camera import orbit_minicam

print('max co-resident depth-sort clusters:', _lib.lib().dimo_debug_max_sort_clusters())
N, W, H = 100000, 512, 512
xyz, scales, rot, op, shs = [t.cuda() for t in gp.scene_inputs(N)]
for B in (1, 2, 4, 8, 16):
    cams = []
    for v in range(B):
        cam = orbit_minicam(v % 8, 8, W, H)
        cams.append(draster.pack_cameras(cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                                         math.tan(cam.FoVx / 2), math.tan(cam.FoVy / 2), torch.ones(3, device="cuda")))
    cams = torch.cat(cams)
    for it in range(3):
        draster.rasterize_batch(cams, xyz, scales, rot, op, W, H, shs=shs, depth_normal=False)
    torch.cuda.synchronize()
    _lib.PROFILE.reset(enabled=True)
    for it in range(5):
        draster.rasterize_batch(cams, xyz, scales, rot, op, W, H, shs=shs, depth_normal=False)
    s = _lib.PROFILE.summary()
    _lib.PROFILE.enabled = False
    print(f"B={B:2d} R={_lib.PROFILE.extra.get('R')}: " + "  ".join(f"{k[5:]} {v['ms'] / v['calls'] * 1000:.1f} us" for k, v in s.items()))
