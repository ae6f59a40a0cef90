"""Host side of the tile rasteriser: torch owns the memory, libdimo_b200 does the work.

`rasterize_batch` renders B frames that share the image size in one set of kernel launches
(preprocess -> per-frame depth sort -> counting sort by tile -> blend) and is differentiable through
`_Rasterize` (blend backward -> preprocess backward).  The single-frame reference API
(diff_gauss / diff_gaussian_rasterization GaussianRasterizer, renderer/latent_gs_renderer.py:1147,
1256-1277) is B=1 of the same path (dimo_b200/shims).
"""
import torch

from . import _lib

CAM_FLOATS = 40
SPLAT_FLOATS = 16
TILE = 16


def pack_cameras(viewmatrix, projmatrix, campos, tanfovx, tanfovy, bg):
    """[B?,4,4],[B?,4,4],[B?,3], python floats or [B] tensors, [B?,3] -> [B,40] device block
    (layout: include/dimo_b200.h DIMO_CAM_FLOATS).  No host sync."""
    dev = viewmatrix.device
    V = viewmatrix.reshape(-1, 16).float()
    B = V.shape[0]
    P = projmatrix.reshape(-1, 16).float().expand(B, 16)
    C = campos.reshape(-1, 3).float().expand(B, 3)
    if not torch.is_tensor(tanfovx):
        tan = torch.tensor([[float(tanfovx), float(tanfovy)]], dtype=torch.float32).to(dev, non_blocking=True).expand(B, 2)
    else:
        tan = torch.stack([tanfovx.float().reshape(-1), tanfovy.float().reshape(-1)], dim=1).to(dev).expand(B, 2)
    G = bg.reshape(-1, 3).float().to(dev).expand(B, 3)
    return torch.cat([V, P, C, tan, G], dim=1).contiguous()


def _bstride(t, B, per_frame_numel, blocks=None):
    """element stride between frames: 0 if the tensor is shared by all frames.  blocks: number of blocks the
    tensor may hold instead of B (inputs addressed through frame_src)."""
    if t is None:
        return 0
    if t.numel() == per_frame_numel:
        return 0
    assert t.numel() == (B if blocks is None else blocks) * per_frame_numel, "batched argument has wrong size"
    return per_frame_numel


class RasterState:
    """Buffers kept from forward for backward / inspection (all torch-owned)."""
    __slots__ = ("B", "N", "W", "H", "R", "count_overflow", "cams", "splats", "radii", "tiles_touched", "rects", "perm",
                 "keys_sorted",
                 "vals_sorted", "ranges", "final_T", "n_contrib", "sh_degree", "sh_coeffs",
                 "scale_modifier", "frame_src", "n_src", "value_bits", "act_flags")

    # The sorted instance list is either (keys_sorted, vals_sorted) = (frame*tiles + tile, index into B*N) or, packed
    # (value_bits > 0, include/dimo_b200.h dimo_raster_packed_value_bits), single words in vals_sorted.  These two
    # accessors give the unpacked view in both cases (tests / inspection; the kernels decode on the fly).
    def tile_keys(self, count=None):
        w = self.vals_sorted if count is None else self.vals_sorted[:count]
        if self.value_bits > 0:
            return (w.long() & 0xFFFFFFFF) >> self.value_bits
        k = self.keys_sorted if count is None else self.keys_sorted[:count]
        return k.long() & 0xFFFFFFFF

    def record_ids(self, count=None):
        w = self.vals_sorted if count is None else self.vals_sorted[:count]
        w = w.long() & 0xFFFFFFFF
        if self.value_bits > 0:
            tiles = ((self.W + TILE - 1) // TILE) * ((self.H + TILE - 1) // TILE)
            frame = torch.clamp((w >> self.value_bits) // tiles, max=self.B - 1)      # sentinel slots: clamped, unused
            return (w & ((1 << self.value_bits) - 1)) + frame * self.N
        return w


def _forward_impl(cams, means3D, scales, rotations, opacities, shs, colors_precomp, B, N, W, H, sh_degree,
                  scale_modifier, capacity=None, frame_src=None, n_src=None, depth_normal=True, act_flags=0):
    """capacity=None: exact mode -- the instance count R is read back from the device once (a host sync, like the
    upstream rasterisers) and buffers are sized to it.  capacity=int: sync-free mode for CUDA graphs -- buffers hold
    `capacity` instance slots, `st.count_overflow` (device i32[2]) receives the true count and an overflow flag."""
    import ctypes
    dev = means3D.device
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    st = RasterState()
    st.B, st.N, st.W, st.H = B, N, W, H
    st.cams = cams
    st.sh_degree = int(sh_degree)
    st.sh_coeffs = 0 if shs is None else int(shs.shape[-2])
    st.scale_modifier = float(scale_modifier)
    st.frame_src, st.n_src = frame_src, n_src
    st.act_flags = int(act_flags)
    BN = B * N
    st.splats = torch.empty(BN, SPLAT_FLOATS, **f32)
    st.radii = torch.empty(BN, **i32)
    st.tiles_touched = torch.empty(BN, **i32)
    st.rects = torch.empty(BN, 2, **i32)
    L = _lib.lib()
    sort_scratch = torch.empty(3 * BN, **i32)
    perm2 = torch.empty(2 * BN, **i32)
    st.perm = perm2[BN:]
    total = torch.empty(1, dtype=torch.int64, device=dev)
    R_host = ctypes.c_int64(0)
    s = _lib.stream()
    _lib.call("dimo_raster_preprocess", B, N, W, H, st.sh_degree, st.sh_coeffs, st.scale_modifier, st.act_flags, _lib.ptr(cams),
              _lib.ptr(frame_src),
              _lib.ptr(means3D), _bstride(means3D, B, N * 3, n_src),
              _lib.ptr(scales), _bstride(scales, B, N * 3),
              _lib.ptr(rotations), _bstride(rotations, B, N * 4, n_src),
              _lib.ptr(opacities), _bstride(opacities, B, N),
              _lib.ptr(shs), _bstride(shs, B, N * st.sh_coeffs * 3),
              _lib.ptr(colors_precomp), _bstride(colors_precomp, B, N * 3),
              _lib.ptr(st.splats), _lib.ptr(st.radii), _lib.ptr(st.tiles_touched), _lib.ptr(st.rects),
              _lib.ptr(sort_scratch), _lib.ptr(perm2), _lib.ptr(total),
              ctypes.addressof(R_host) if capacity is None else None, s)
    if capacity is None:
        R = int(R_host.value)
        st.R = R
        st.count_overflow = None
        _lib.PROFILE.extra["R"] = R
    else:
        R = int(capacity)
        st.R = None                                  # true count lives on the device: st.count_overflow[0]
        st.count_overflow = torch.zeros(2, **i32)
    gx, gy = (W + TILE - 1) // TILE, (H + TILE - 1) // TILE
    Ra = max(R, 1)
    st.value_bits = int(L.dimo_raster_packed_value_bits(B, N, W, H))
    packed = st.value_bits > 0
    st.keys_sorted = None if packed else torch.empty(Ra, **i32)   # frame*tiles + tile, tile-major
    st.vals_sorted = torch.empty(Ra, **i32)          # record indices, or packed (key | index-in-frame) words
    st.ranges = torch.empty(B * gx * gy, 2, **i32)
    bin_bytes = L.dimo_raster_bin_temp_bytes(B, N, W, H)
    bin_temp = torch.empty(bin_bytes, dtype=torch.uint8, device=dev)
    _lib.call("dimo_raster_bin", B, N, W, H, R, _lib.ptr(st.rects), _lib.ptr(st.perm),
              _lib.ptr(st.keys_sorted), _lib.ptr(st.vals_sorted), _lib.ptr(bin_temp), bin_bytes, _lib.ptr(st.ranges),
              _lib.ptr(st.count_overflow), s)
    color = torch.empty(B, 3, H, W, **f32)
    # depth_normal=False: nobody reads depth / normal (the MSE + SSIM + mask step) -> the blend kernel's 3-channel
    # variant (dimo_raster_blend_fwd with out_depth = out_normal = NULL)
    depth = torch.empty(B, 1, H, W, **f32) if depth_normal else None
    normal = torch.empty(B, 3, H, W, **f32) if depth_normal else None
    alpha = torch.empty(B, 1, H, W, **f32)
    st.final_T = torch.empty(B, H, W, **f32)
    st.n_contrib = torch.empty(B, H, W, **i32)
    _lib.call("dimo_raster_blend_fwd", B, N, W, H, st.value_bits, _lib.ptr(cams), _lib.ptr(st.splats), _lib.ptr(st.vals_sorted),
              _lib.ptr(st.ranges),
              _lib.ptr(color), _lib.ptr(depth), _lib.ptr(normal), _lib.ptr(alpha), _lib.ptr(st.final_T),
              _lib.ptr(st.n_contrib), s)
    return color, depth, normal, alpha, st


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, scales, rotations, opacities, shs, colors_precomp, cams, B, N, W, H,
                sh_degree, scale_modifier, state_out, capacity, frame_src, n_src, depth_normal=True, act_flags=0,
                grad_sink=None):
        color, depth, normal, alpha, st = _forward_impl(cams, means3D, scales, rotations, opacities, shs,
                                                        colors_precomp, B, N, W, H, sh_degree, scale_modifier,
                                                        capacity, frame_src, n_src, depth_normal, act_flags)
        ctx.st = st
        ctx.grad_sink = grad_sink
        ctx.set_materialize_grads(False)     # unused outputs (depth / normal in the image-loss step) arrive as None
        ctx.save_for_backward(means3D, scales, rotations, shs, opacities)
        ctx.shapes = (means3D.shape, None if means2D is None else means2D.shape, scales.shape, rotations.shape,
                      opacities.shape, None if shs is None else shs.shape,
                      None if colors_precomp is None else colors_precomp.shape)
        if state_out is not None:
            state_out.append(st)
        radii = st.radii.view(B, N)
        ctx.mark_non_differentiable(radii)
        return color, depth, normal, alpha, radii

    @staticmethod
    def backward(ctx, g_color, g_depth, g_normal, g_alpha, _g_radii):
        st = ctx.st
        means3D, scales, rotations, shs, opacities = ctx.saved_tensors
        B, N, W, H = st.B, st.N, st.W, st.H
        dev = means3D.device
        f32 = dict(dtype=torch.float32, device=dev)
        zeros = lambda *s: torch.zeros(*s, **f32)
        g_color = g_color.contiguous() if g_color is not None else zeros(B, 3, H, W)
        if g_depth is None and g_normal is None:
            pass        # NULL for both: the kernel drops the four channels (dimo_raster_blend_bwd, DN = false)
        else:
            g_depth = g_depth.contiguous() if g_depth is not None else zeros(B, 1, H, W)
            g_normal = g_normal.contiguous() if g_normal is not None else zeros(B, 3, H, W)
        g_alpha = g_alpha.contiguous() if g_alpha is not None else zeros(B, 1, H, W)
        s = _lib.stream()
        # deterministic mode: the kernel accumulates into an int64 record table (zeroed by the call like the fp32 one)
        dsplats = torch.empty(B * N, SPLAT_FLOATS, dtype=torch.int64 if _lib.deterministic() else torch.float32, device=dev)
        _lib.call("dimo_raster_blend_bwd", B, N, W, H, st.value_bits, _lib.ptr(st.cams), _lib.ptr(st.splats),
                  _lib.ptr(st.vals_sorted), _lib.ptr(st.ranges), _lib.ptr(st.final_T), _lib.ptr(st.n_contrib), _lib.ptr(g_color),
                  _lib.ptr(g_depth), _lib.ptr(g_normal), _lib.ptr(g_alpha), _lib.ptr(dsplats), s)
        dsplats = _lib.acc_result(dsplats)
        use_sh = shs is not None
        # parameters shared by all frames (the training step: one set of scales / opacities / SHs): their gradients are
        # summed over the frames inside the kernel -> [N, *] outputs, no per-frame tensors and no dimo_segment_sum
        shared = (B > 1 and use_sh and _bstride(scales, B, N * 3) == 0 and _bstride(opacities, B, N) == 0
                  and _bstride(shs, B, N * st.sh_coeffs * 3) == 0)
        lead = (N,) if shared else (B, N)
        # grad_sink = (scales.grad, opacities.grad, shs.grad, on_done): the kernel adds the frame sums straight into the
        # parameters' gradient buffers (all three or none)
        sink = ctx.grad_sink if (shared and ctx.grad_sink is not None and not _lib.deterministic()
                                 and all(t is not None for t in ctx.grad_sink[:3])) else None
        d_means3D = torch.empty(B, N, 3, **f32)
        d_means2D = torch.empty(B, N, 3, **f32) if ctx.needs_input_grad[1] else None
        d_scales = sink[0] if sink else torch.empty(*lead, 3, **f32)
        d_rot = torch.empty(B, N, 4, **f32)
        d_op = sink[1] if sink else torch.empty(*lead, **f32)
        d_shs = sink[2] if sink else (torch.empty(*lead, st.sh_coeffs, 3, **f32) if use_sh else None)
        d_col = None if use_sh else torch.empty(B, N, 3, **f32)
        _lib.call("dimo_raster_preprocess_bwd", B, N, W, H, st.sh_degree, st.sh_coeffs, st.scale_modifier, st.act_flags,
                  _lib.ptr(st.cams), _lib.ptr(st.frame_src),
                  _lib.ptr(means3D), _bstride(means3D, B, N * 3, st.n_src),
                  _lib.ptr(scales), _bstride(scales, B, N * 3),
                  _lib.ptr(rotations), _bstride(rotations, B, N * 4, st.n_src),
                  _lib.ptr(opacities), _bstride(opacities, B, N),
                  _lib.ptr(shs), _bstride(shs, B, N * st.sh_coeffs * 3) if use_sh else 0,
                  _lib.ptr(st.radii), _lib.ptr(dsplats), _lib.ptr(d_means3D), _lib.ptr(d_means2D),
                  _lib.ptr(d_scales), _lib.ptr(d_rot), _lib.ptr(d_op), _lib.ptr(d_shs), _lib.ptr(d_col),
                  (2 if sink else 1) if shared else 0, s)
        if sink:
            d_scales = d_op = d_shs = None
            if sink[3] is not None:
                sink[3]()
        sh_m3, sh_m2, sh_sc, sh_rot, sh_op, sh_shs, sh_col = ctx.shapes

        def fit(g, shape, per_frame, mapped=False):
            """[B, ...per-frame] -> the input's own shape: summed over all frames when the input was shared, over the
            frames of each block when it was addressed through frame_src (one dimo_segment_sum launch either way)"""
            if shape is None or g is None:
                return None
            if g.numel() == per_frame:         # already summed over the frames by the kernel
                return g.reshape(shape)
            numel = 1
            for d in shape:
                numel *= d
            if numel == B * per_frame and not (mapped and st.frame_src is not None):
                return g.reshape(shape)
            U = numel // per_frame
            if B > 1024:                       # beyond the kernel's row table: plain torch reduction (shared inputs only)
                assert U == 1
                return g.reshape(B, per_frame).sum(dim=0).reshape(shape)
            out = torch.empty(U, per_frame, **f32)
            _lib.call("dimo_segment_sum", B, U, per_frame, _lib.ptr(st.frame_src) if U > 1 else None, _lib.ptr(g),
                      _lib.ptr(out), s)
            return out.reshape(shape)

        return (fit(d_means3D, sh_m3, N * 3, True), fit(d_means2D, sh_m2, N * 3) if ctx.needs_input_grad[1] else None,
                fit(d_scales, sh_sc, N * 3), fit(d_rot, sh_rot, N * 4, True), fit(d_op, sh_op, N),
                fit(d_shs, sh_shs, N * st.sh_coeffs * 3) if use_sh else None,
                fit(d_col, sh_col, N * 3) if not use_sh else None,
                None, None, None, None, None, None, None, None, None, None, None, None, None, None)


def rasterize_batch(cams, means3D, scales, rotations, opacities, W, H, shs=None, colors_precomp=None,
                    sh_degree=0, scale_modifier=1.0, means2D=None, state_out=None, capacity=None, frame_src=None,
                    depth_normal=True, raw_activations=False, direct_grads=False, on_done=None):
    """cams [B,40]; means3D [B,N,3] or [N,3]; scales [N,3]; rotations [B,N,4] or [N,4]; opacities [N,1]/[N];
    shs [N,K,3] xor colors_precomp [B?,N,3].  frame_src [B] int32 (device): means3D / rotations are [U,N,*] and
    frame b uses block frame_src[b] (frames that differ only in the view share one deformation).
    depth_normal=False: depth and normal are not rendered (returned as None).
    raw_activations=True (or a bit mask: 1 = scales, 2 = opacities): `scales` are log-scales and `opacities` logits (the model's raw _scaling / _opacity); exp and
    sigmoid run inside the projection kernels and the gradients come back w.r.t. the raw parameters.
    direct_grads: when scales, opacities and shs are leaf parameters shared by all frames whose .grad is preallocated
    (_lib.grad_sink), the backward adds their gradients straight into it and calls on_done([the three parameters]).
    Returns color [B,3,H,W], depth [B,1,H,W], normal [B,3,H,W], alpha [B,1,H,W], radii [B,N] int32."""
    if (shs is None) == (colors_precomp is None):
        raise ValueError("Please provide exactly one of either SHs or precomputed colors!")
    B = cams.shape[0]
    N = scales.shape[-2]
    c = lambda t: None if t is None else t.contiguous().float()
    n_src = None
    if frame_src is not None:
        assert frame_src.dtype == torch.int32 and frame_src.numel() == B and means3D.dim() == 3
        n_src = int(means3D.shape[0])
        if B > 1024:
            raise ValueError("frame_src supports at most 1024 frames per launch set")
    sink = None
    if direct_grads and torch.is_grad_enabled() and shs is not None:
        ps = (scales, opacities, shs)
        sk = [_lib.grad_sink(p) if (p.dtype == torch.float32 and p.is_contiguous()) else None for p in ps]
        if all(t is not None for t in sk):
            sink = (*sk, (lambda: on_done(list(ps))) if on_done is not None else None)
    return _Rasterize.apply(c(means3D), means2D, c(scales), c(rotations), c(opacities), c(shs), c(colors_precomp),
                            cams.contiguous(), B, N, int(W), int(H), int(sh_degree), float(scale_modifier), state_out,
                            capacity, frame_src, n_src, bool(depth_normal),
                            (3 if raw_activations is True else int(raw_activations or 0)), sink)
