from . import mesh_laplacian_smoothing  # noqa: F401
