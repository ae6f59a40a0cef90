"""`from simple_knn._C import distCUDA2` (renderer/latent_gs_renderer.py:17,426)."""
from dimo_b200.knn import dist3nn as _dist3nn


def distCUDA2(points):
    return _dist3nn(points)
