#!/usr/bin/env python
"""Finds Gaussians whose integer outputs differ between the oracle and the CUDA preprocess at a given shape and prints
the oracle's intermediates for them (bring-up / parity debugging)."""
import math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_parity as gp
from dimo_b200 import raster as draster
from dimo_b200.camera import orbit_minicam
from oracle import raster as oraster, camera as ocamera

N, W, H, view, nviews = 30000, 1024, 1024, 17, 120
xyz, scales, rot, op, shs = gp.scene_inputs(N)
ocam = ocamera.orbit_cam(view, nviews, W, H)
pre = oraster.preprocess(xyz, scales, rot, op, ocam.world_view_transform, ocam.full_proj_transform, ocam.camera_center,
                         ocam.tanfovx, ocam.tanfovy, W, H, 1.0, shs, 0)
print("oracle keys:", sorted(pre.keys()))
cam = orbit_minicam(view, nviews, W, H)
cams = draster.pack_cameras(cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                            math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), torch.ones(3, device="cuda"))
st = []
draster.rasterize_batch(cams, xyz.cuda(), scales.cuda(), rot.cuda(), op.cuda(), W, H, shs=shs.cuda(), state_out=st)
s = st[0]
r_c = s.radii.cpu(); r_o = pre["radii"]
bad = torch.nonzero(r_c != r_o).flatten()
print("mismatches:", bad.tolist())
for i in bad.tolist():
    print("idx", i, "cuda radius", int(r_c[i]), "oracle radius", int(r_o[i]), "tiles cuda", int(s.tiles_touched[i]), "oracle", int(pre["tiles_touched"][i]))
    ca, cb, cc = [pre["cov2d"][i, k] for k in range(3)]
    det = ca * cc - cb * cb
    mid = 0.5 * (ca + cc)
    root = oraster._sqrt_rn(torch.clamp_min(mid * mid - det, oraster.LAMBDA_FLOOR))
    lam = torch.maximum(mid + root, mid - root)
    rf = 3.0 * oraster._sqrt_rn(lam)
    hx = lambda v: hex(int(v.reshape(1).view(torch.int32)) & 0xFFFFFFFF)
    for nm, v in (("ca", ca), ("cb", cb), ("cc", cc), ("det", det), ("mid", mid), ("root", root), ("lam", lam), ("3sqrt", rf),
                  ("depth", pre["depth"][i]), ("x", pre["xy"][i, 0]), ("y", pre["xy"][i, 1])):
        print("   ", nm, repr(float(v)), hx(v))
    print("    rect", pre["rect"][i].tolist())
    rec = s.splats[i]
    print("    cuda record x,y,depth:", float(rec[0]), float(rec[1]), float(rec[10]), hx(rec[0].cpu()), hx(rec[1].cpu()), hx(rec[10].cpu()))
    # conic from the cuda record: a2 = -0.5*log2e*conic_a ...
    print("    cuda a2,b2,c2:", float(rec[2]), float(rec[3]), float(rec[4]))
    con = pre["conic"][i]
    L2E = 1.4426950408889634
    print("    oracle a2,b2,c2:", float(-0.5 * L2E * con[0]), float(-L2E * con[1]), float(-0.5 * L2E * con[2]))
