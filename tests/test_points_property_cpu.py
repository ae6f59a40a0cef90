"""Size-independent properties of the point-set oracles (oracle/points.py) -- the same properties the GPU tests
check on the kernels at sizes where a brute-force comparison is not practical."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import points as op


@settings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(st.integers(2, 300), st.integers(0, 2 ** 31 - 1))
def test_fps_properties(n, seed):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(n, 3, generator=g)
    k = max(1, n // 3)
    idx = op.fps(pts, k)
    assert idx[0] == 0 and len(set(idx.tolist())) == k                       # distinct picks, fixed start
    p = pts[idx].double()
    d = torch.cdist(p, p)
    gaps = [float(d[i, :i].min()) for i in range(1, k)]
    assert all(b <= a * (1 + 1e-6) for a, b in zip(gaps, gaps[1:]))          # coverage radius never grows
    if k > 1:                                                                # every pick is the farthest point at its turn
        rest = torch.cdist(pts.double(), p[:k - 1]).min(dim=1).values
        assert abs(float(rest.max()) - gaps[-1]) <= 1e-6


@settings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(st.integers(1, 120), st.integers(1, 12), st.floats(0.02, 0.6), st.integers(0, 2 ** 31 - 1))
def test_ball_query_properties(n, k, radius, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand(2, n, 3, generator=g)
    d, idx, nn = op.ball_query(p, p, K=k, radius=radius)
    d2 = torch.cdist(p.double(), p.double()) ** 2
    for b in range(2):
        for i in range(n):
            got = [j for j in idx[b, i].tolist() if j >= 0]
            inside = [j for j in range(n) if float(d2[b, i, j]) < radius * radius * (1 - 1e-5)]
            maybe = [j for j in range(n) if float(d2[b, i, j]) < radius * radius * (1 + 1e-5)]
            assert got == sorted(got) and set(got) <= set(maybe)               # index order, all inside the ball
            assert len(got) >= min(k, len(inside)) and len(got) <= k           # the FIRST K of them
            if len(got) < k:
                assert set(inside) <= set(got)
            else:
                assert all(j in got or j > got[-1] for j in inside)
            assert idx[b, i, len(got):].tolist() == [-1] * (k - len(got))
            assert float(d[b, i, len(got):].abs().sum()) == 0.0
            assert torch.equal(nn[b, i, :len(got)], p[b, got]) if got else True
