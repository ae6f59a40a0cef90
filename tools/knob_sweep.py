#!/usr/bin/env python
"""Runs the bench step eagerly with profiling for a few tuning-knob settings and prints the per-call breakdown."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from dimo_b200 import _lib, trainstep
from dimo_b200.camera import orbit_minicam

wl = bench.WORKLOADS["c3"]
dev = torch.device("cuda:0")
r, _ = bench.build_model(wl, 0, dev)
ts = trainstep.TrainStep(r, lr=1e-5)
H, W = wl["H"], wl["W"]; S = wl["bm"] * wl["bv"] * wl["bf"]
cams_all = [orbit_minicam(v, wl["views"], W, H, device=dev) for v in range(wl["views"])]
gt = torch.rand(S, 3, H, W, device=dev); mk = torch.rand(S, 1, H, W, device=dev)

def run(steps):
    for i in range(steps):
        fr = bench.step_schedule(wl, i)
        ts.run([cams_all[v] for (_, v, _) in fr], [f / wl["frames"] for (_, _, f) in fr], [m for (m, _, _) in fr], gt, mk, wl["bm"])

run(3)
for knob2 in (0, 64, 32, 16):
    _lib.call("dimo_tc_debug_set", 2, knob2)
    run(1); torch.cuda.synchronize()
    _lib.PROFILE.reset(enabled=True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(5); e1.record()
    prof = _lib.PROFILE.summary(); _lib.PROFILE.enabled = False
    print(f"wgrad_ctas={knob2 or 148}: step {e0.elapsed_time(e1)/5:.3f} ms", {k: round(v['ms']/5, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])[:7]})
