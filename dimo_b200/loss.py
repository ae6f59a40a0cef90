"""Fused image losses (host side): SSIM + L1 + MSE in one forward and one backward kernel.
ssim() == src/loss.py:144 ssim(img1, img2) == fused_ssim.fused_ssim(img1, img2) (main_test_dimo.py:979);
l1 == src/loss.py:178; mse == F.mse_loss (main_train_dimo.py:333)."""
import torch

from . import _lib


class _ImageLoss(torch.autograd.Function):
    """returns the three means [ssim, l1, mse] as a [3] tensor; differentiable w.r.t. img1 only
    (the reference's second argument is ground truth)."""

    @staticmethod
    def forward(ctx, img1, img2, need_ssim_grad, clamp01=False):
        img1 = img1.contiguous().float()
        img2 = img2.contiguous().float()
        if img1.dim() == 3:
            img1 = img1[None]; img2 = img2[None]
        B, C, H, W = img1.shape
        dev = img1.device
        sums = torch.empty(3, dtype=torch.float32, device=dev)
        keep = bool(need_ssim_grad) and ctx.needs_input_grad[0]      # (img1 itself may be a no-grad copy by now)
        dm = torch.empty(3, B, C, H, W, dtype=torch.float32, device=dev) if keep else None
        _lib.call("dimo_ssim_fwd", B, C, H, W, int(clamp01), _lib.ptr(img1), _lib.ptr(img2), _lib.ptr(sums), _lib.ptr(dm),
                  None, None, 0.0, 0.0, 0.0, _lib.stream())
        ctx.save_for_backward(img1, img2, dm)
        ctx.dims = (B, C, H, W)
        ctx.clamp01 = int(clamp01)
        return sums / float(B * C * H * W)

    @staticmethod
    def backward(ctx, g):
        img1, img2, dm = ctx.saved_tensors
        B, C, H, W = ctx.dims
        # weights are host scalars in the C ABI; one small D2H read of the three upstream grads
        gw = (g.float() / float(B * C * H * W)).tolist()
        if dm is None and gw[0] != 0.0:
            raise RuntimeError("ssim gradient requested but the derivative maps were not kept")
        out = torch.empty_like(img1)
        _lib.call("dimo_ssim_bwd", B, C, H, W, ctx.clamp01, _lib.ptr(img1), _lib.ptr(img2), _lib.ptr(dm), gw[0], gw[1], gw[2],
                  None, None, _lib.ptr(out), _lib.stream())
        return out, None, None, None


def image_losses(img1, img2, need_ssim_grad=True, clamp01=False):
    """-> tensor [3] = (mean ssim map, mean |a-b|, mean (a-b)^2); clamp01 clamps img1 to [0,1] on load"""
    return _ImageLoss.apply(img1, img2, need_ssim_grad, clamp01)


def ssim(img1, img2, window_size=11, size_average=True):
    if window_size != 11 or not size_average:
        raise NotImplementedError("dimo_b200 ssim: window 11 / size_average=True only (what the reference calls)")
    return image_losses(img1, img2)[0]


def fused_ssim(img1, img2, padding="same", train=True):
    if padding != "same":
        raise NotImplementedError("dimo_b200 fused_ssim: 'same' (zero) padding only")
    return image_losses(img1, img2, need_ssim_grad=train)[0]


class _Smoothness(torch.autograd.Function):
    """l_smooth * edge_aware_smoothness(depth, rgb) + l_bilateral * bilateral_normal_smoothness(normal, rgb) summed over
    `groups` equal groups of frames (the reference evaluates both per motion, main_train_dimo.py:363-372).
    rgb [B,3,H,W], depth [B,1,H,W], normal [B,3,H,W]; differentiable in all three."""

    @staticmethod
    def forward(ctx, rgb, depth, normal, groups, l_smooth, l_bilateral, clamp01):
        rgb = rgb.contiguous().float(); depth = depth.contiguous().float(); normal = normal.contiguous().float()
        B, _, H, W = rgb.shape
        per = B // groups
        nx, ny = float(per * H * max(W - 1, 1)), float(per * max(H - 1, 1) * W)
        w = (l_smooth / nx, l_smooth / ny, l_bilateral / (3.0 * nx), l_bilateral / (3.0 * ny))
        sums = torch.empty(4, dtype=torch.float32, device=rgb.device)
        loss = torch.zeros((), dtype=torch.float32, device=rgb.device)
        _lib.call("dimo_smooth_fwd", B, H, W, int(clamp01), _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(normal),
                  _lib.ptr(sums), _lib.ptr(loss), w[0], w[1], w[2], w[3], _lib.stream())
        ctx.save_for_backward(rgb, depth, normal)
        ctx.w, ctx.clamp01 = w, int(clamp01)
        return loss

    @staticmethod
    def backward(ctx, g):
        rgb, depth, normal = ctx.saved_tensors
        B, _, H, W = rgb.shape
        w = ctx.w
        g = g.contiguous().float()
        d_rgb = torch.empty_like(rgb); d_depth = torch.empty_like(depth); d_normal = torch.empty_like(normal)
        _lib.call("dimo_smooth_bwd", B, H, W, ctx.clamp01, _lib.ptr(rgb), _lib.ptr(depth), _lib.ptr(normal),
                  w[0], w[1], w[2], w[3], _lib.ptr(g), _lib.ptr(d_rgb), 0, _lib.ptr(d_depth), _lib.ptr(d_normal),
                  _lib.stream())
        return d_rgb, d_depth, d_normal, None, None, None, None


def smoothness_losses(rgb, depth, normal, groups=1, lambda_smooth=1.0, lambda_bilateral=1.0, clamp01=False):
    """NCHW entry point of the two step regularisers; returns their weighted sum as a scalar tensor."""
    return _Smoothness.apply(rgb, depth, normal, groups, float(lambda_smooth), float(lambda_bilateral), clamp01)


def compute_edge_aware_smoothness_loss(depth, rgb):
    """src/loss.py:64 signature: depth [batch,H,W,1], rgb [batch,H,W,3] (channel-last, as the reference calls it)."""
    d = depth.permute(0, 3, 1, 2)
    c = rgb.permute(0, 3, 1, 2)
    n = torch.zeros_like(c)
    return _Smoothness.apply(c, d, n, 1, 1.0, 0.0, False)


def compute_bilateral_normal_smoothness_loss(normal, rgb):
    """src/loss.py:87 signature: normal [batch,H,W,3], rgb [batch,H,W,3]."""
    n = normal.permute(0, 3, 1, 2)
    c = rgb.permute(0, 3, 1, 2)
    d = torch.zeros_like(c[:, :1])
    return _Smoothness.apply(c, d, n, 1, 0.0, 1.0, False)


def l1_loss(a, b):
    return image_losses(a, b, need_ssim_grad=False)[1]


def mse_loss(a, b):
    return image_losses(a, b, need_ssim_grad=False)[2]


def compute_tv_norm(values, losstype="l2"):
    """src/loss.py:109-129 (imported by main_train_dimo.py:28, never called): squared / absolute forward differences of
    a channel-last [batch,H,W,C] tensor -> [batch,H-1,W-1,C].  Plain tensor arithmetic: not on any measured path."""
    v00, v01, v10 = values[..., :-1, :-1, :], values[..., :-1, 1:, :], values[..., 1:, :-1, :]
    if losstype == "l2":
        return (v00 - v01) ** 2 + (v00 - v10) ** 2
    if losstype == "l1":
        return (v00 - v01).abs() + (v00 - v10).abs()
    raise ValueError(f"losstype must be l2 or l1 but is {losstype}")
