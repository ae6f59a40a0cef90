#!/usr/bin/env python
"""Times the TimeNet forward and forward+backward launch sets at the bench shape (G=8 pairs x M=512 control points =
4096 rows) as CUDA-graph replays, i.e. what they cost inside the whole-step graph (the per-call CUDA events of the
eager profile include launch gaps).  Run on the GPU box: python tools/timenet_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dimo_b200.deform import TimeNet  # noqa: E402


def main():
    torch.manual_seed(0)
    G, M, L = int(os.environ.get("TN_G", "8")), 512, 32
    net = TimeNet(latent_code_dim=L).cuda()
    with torch.no_grad():
        for lin in (net.pts_layers[-1], net.rot_layers[-1]):
            lin.weight.copy_(0.01 * torch.randn_like(lin.weight))
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    net.direct_grads = True
    pts = (torch.rand(M, 3, device="cuda") - 0.5).requires_grad_(True)
    times = torch.rand(G, device="cuda")
    lat = torch.randn(G, L, device="cuda", requires_grad=True)
    wx = torch.randn(G, M, 3, device="cuda"); wq = torch.randn(G, M, 4, device="cuda")

    def fwd():
        with torch.no_grad():
            return net.forward_batched(pts, times, lat)

    def fwd_bwd():
        dx, dq = net.forward_batched(pts, times, lat)
        torch.autograd.backward([dx, dq], [wx, wq])

    for name, fn in (("forward", fwd), ("forward+backward", fwd_bwd)):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(5):
            g.replay()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"TimeNet {name}: {e0.elapsed_time(e1) / 200 * 1000:.1f} us per replay (G={G}, M={M}, DIMO_TC={os.environ.get('DIMO_TC', '1')})")


if __name__ == "__main__":
    main()
