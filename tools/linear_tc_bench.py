#!/usr/bin/env python
"""Per-launch cost model of dimo_linear_tc inside a CUDA graph: 20 chained launches per graph, for several K (number
of 64-wide K tiles) and both MMA modes -> fixed cost per launch (intercept) and cost per K tile (slope)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dimo_b200 import _lib  # noqa: E402


def bench(R, K, No, relu=1, chain=20, reps=50):
    X = torch.randn(R, K, device="cuda"); W = torch.randn(No, K, device="cuda") / K ** 0.5
    b = torch.randn(No, device="cuda"); Y = torch.empty(R, No, device="cuda")

    def run():
        for _ in range(chain):
            _lib.call("dimo_linear_tc", R, K, No, _lib.ptr(X), K, None, 0, _lib.ptr(W), _lib.ptr(b), _lib.ptr(Y), No,
                      relu, 0, _lib.stream())
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        run()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    for _ in range(3):
        g.replay()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / chain * 1000.0


def main():
    for single in (0, 1):
        _lib.call("dimo_tc_debug_set", 1, single)
        for R in (4096, 1024):
            for K in (64, 128, 256, 512, 1024):
                print(f"single_pass={single} R={R} K={K} No=256: {bench(R, K, 256):.2f} us per launch")
    _lib.call("dimo_tc_debug_set", 1, 0)
    # ablations (results are garbage, only the time matters): 1 = no global loads, 2 = no shared stores, 4 = no MMAs
    for ab in (1, 2, 4, 3, 5, 6, 7):
        _lib.call("dimo_tc_debug_set", 7, ab)
        print(f"ablate={ab} R=4096 No=256: K=256 {bench(4096, 256, 256):.2f} us, K=1024 {bench(4096, 1024, 256):.2f} us per launch")
    _lib.call("dimo_tc_debug_set", 7, 0)
    # an empty-ish reference: the same chain with a trivially small problem
    print(f"R=128 K=64 No=64: {bench(128, 64, 64):.2f} us per launch")


if __name__ == "__main__":
    main()
