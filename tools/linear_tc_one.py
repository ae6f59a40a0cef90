#!/usr/bin/env python
"""One warm dimo_linear_tc launch (R=4096, K=1024, No=256) for an ncu source-level capture."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dimo_b200 import _lib
R, K, No = 4096, int(os.environ.get("K", "1024")), 256
X = torch.randn(R, K, device="cuda"); W = torch.randn(No, K, device="cuda") / K ** 0.5
b = torch.randn(No, device="cuda"); Y = torch.empty(R, No, device="cuda")
for _ in range(3):
    _lib.call("dimo_linear_tc", R, K, No, _lib.ptr(X), K, None, 0, _lib.ptr(W), _lib.ptr(b), _lib.ptr(Y), No, 1, 0, _lib.stream())
torch.cuda.synchronize()
