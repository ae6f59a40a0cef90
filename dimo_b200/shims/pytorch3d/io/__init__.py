import numpy as np
import torch

from dimo_b200 import ply as _ply

__all__ = ["load_ply"]


def load_ply(f):
    """(verts [V,3] float32, faces [F,3] int64) -- imported by utils/deform_utils.py:5 (vertex-only files here:
    faces come back empty; list properties are not parsed)."""
    v = _ply.read_ply(f)["vertex"]
    verts = torch.from_numpy(np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float32))
    return verts, torch.zeros(0, 3, dtype=torch.int64)
