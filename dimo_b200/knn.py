"""Nearest-neighbour ops (host side).  knn(): Gaussian -> k nearest control points
(knn_cuda.KNN, main_train_dimo.py:502-509); dist3nn(): simple_knn._C.distCUDA2
(renderer/latent_gs_renderer.py:426).  No gradients (the reference detaches both)."""
import torch

from . import _lib


def knn(ref, query, k=4):
    """ref [M,3], query [N,3] (CUDA fp32) -> dist [N,k] fp32 Euclidean ascending, idx [N,k] int64."""
    ref = ref.detach().contiguous().float()
    query = query.detach().contiguous().float()
    M, N = ref.shape[0], query.shape[0]
    dist = torch.empty(N, k, dtype=torch.float32, device=query.device)
    idx = torch.empty(N, k, dtype=torch.int64, device=query.device)
    _lib.call("dimo_knn", M, N, k, _lib.ptr(ref), _lib.ptr(query), _lib.ptr(dist), _lib.ptr(idx), _lib.stream())
    return dist, idx


def dist3nn(points):
    points = points.detach().contiguous().float()
    out = torch.empty(points.shape[0], dtype=torch.float32, device=points.device)
    _lib.call("dimo_dist3nn", points.shape[0], _lib.ptr(points), _lib.ptr(out), _lib.stream())
    return out
