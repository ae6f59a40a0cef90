"""Drop-in for `fused_ssim` (rahul-goel/fused-ssim): fused_ssim(img1, img2) -> scalar, differentiable in
img1 (main_test_dimo.py:29,979,1160,1284)."""
from dimo_b200.loss import fused_ssim

__all__ = ["fused_ssim"]
