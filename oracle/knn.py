"""Oracle: nearest-neighbour terms.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED: unlimblue/KNN_CUDA @ 619617b5 and camenduru/simple-knn @ 60f461f4 are absent
from /root/reference.  Semantics are fixed by the reference's call sites:

* ``knn``: main_train_dimo.py:502-509 -- KNN(k=4, transpose_mode=True)(ref[1,M,3], query[1,N,3])
  -> (dist[1,N,k], idx[1,N,k] int64); the consumer squares dist (renderer/latent_gs_renderer.py:1197),
  so dist is the Euclidean (unsquared) distance, ascending, ties -> lower ref index.
* ``dist3nn``: renderer/latent_gs_renderer.py:426 -- distCUDA2(points[N,3]) -> [N] mean squared
  distance to the 3 nearest *other* points.

Squared distances are the fp32 sequence (dx*dx + dy*dy) + dz*dz (separately rounded), which the
CUDA kernels reproduce (-fmad=false) so indices are bit-exact.
"""
import torch


def _sqdist(q, r):
    dx = q[:, None, 0] - r[None, :, 0]
    dy = q[:, None, 1] - r[None, :, 1]
    dz = q[:, None, 2] - r[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def _sqrt_rn(x):
    """correctly rounded fp32 sqrt (torch.sqrt on CPU is not; see oracle/raster.py::_sqrt_rn)"""
    return x.double().sqrt().float() if x.dtype == torch.float32 else torch.sqrt(x)


def knn(ref, query, k=4, chunk=8192):
    """ref [M,3], query [N,3] -> dist [N,k] (Euclidean), idx [N,k] int64."""
    ds, ix = [], []
    for s in range(0, query.shape[0], chunk):
        d2 = _sqdist(query[s:s + chunk], ref)
        order = torch.sort(d2, dim=1, stable=True)
        ds.append(_sqrt_rn(order.values[:, :k]))
        ix.append(order.indices[:, :k])
    return torch.cat(ds), torch.cat(ix)


def dist3nn(points, chunk=2048):
    """points [N,3] -> [N] mean squared distance to the 3 nearest other points."""
    N = points.shape[0]
    out = []
    for s in range(0, N, chunk):
        d2 = _sqdist(points[s:s + chunk], points)
        rows = torch.arange(d2.shape[0])
        d2[rows, rows + s] = float("inf")
        v = torch.sort(d2, dim=1, stable=True).values[:, :3]
        out.append(((v[:, 0] + v[:, 1]) + v[:, 2]) / 3.0)
    return torch.cat(out)
