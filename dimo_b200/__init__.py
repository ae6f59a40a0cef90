"""dimo_b200 -- B200-native (sm_100a) implementation of DIMO's deform -> raster -> loss hot path.

Product code.  Hand-written CUDA behind a C ABI (include/dimo_b200.h, dimo_b200/csrc), bound with
ctypes; PyTorch only owns memory/streams/autograd plumbing.  There is no CPU fallback and nothing
here imports ``oracle``.

``install_shims()`` puts drop-in modules named exactly as the reference imports them
(diff_gauss, diff_gaussian_rasterization, knn_cuda, simple_knn, fused_ssim, plus the slices of pytorch3d,
chamferdist and plyfile the path calls) on sys.path, so the reference's renderer / deform_utils modules import and
run unmodified.
"""
import os
import sys

__version__ = "0.1.0"

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def install_shims():
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)
    return SHIM_DIR
