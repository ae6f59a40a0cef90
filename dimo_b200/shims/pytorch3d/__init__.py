"""Drop-in for the handful of pytorch3d entry points the reference touches (utils/deform_utils.py:3-13,
renderer/latent_gs_renderer.py:23-24, main_train_dimo.py:25,513):
    pytorch3d.ops.sample_farthest_points, pytorch3d.ops.ball_query, pytorch3d.ops.knn_points,
    pytorch3d.transforms.quaternion_to_matrix, pytorch3d.io.load_ply,
    pytorch3d.loss.mesh_laplacian_smoothing.cot_laplacian (imported by the reference, never called).
Not a re-implementation of pytorch3d: everything else raises AttributeError."""
from . import ops, transforms, io, loss  # noqa: F401

__version__ = "0.0.0+dimo_b200"
