"""Deformation ops (host side): TimeNet (positional encoding + 12 fused Linear layers) and K-neighbour
linear-blend skinning, batched over the G unique (motion, t) pairs of a step.

`TimeNet` keeps the reference module's structure and parameter names (renderer/latent_gs_renderer.py:
184-203: deformnet.{0..7}, pts_layers.{0,2}, rot_layers.{0,2}) so released `timenet.pth` checkpoints
load unchanged; its forward runs libdimo_b200 kernels through `_TimeNetFn` instead of F.linear.
"""
import ctypes
import os

import torch
import torch.nn as nn
import torch.nn.init as init

from . import _lib

PTS_FREQS, TIME_FREQS, HIDDEN, DEPTH, SKIP_AFTER = 10, 6, 256, 8, 4

# TimeNet GEMMs run on the tcgen05 tensor cores (3xTF32-compensated, csrc/mlp_tc.cu); DIMO_TC=0 selects the FP32
# SIMT kernels (csrc/mlp.cu) for A/B comparison.  Both are hand-written kernels of this library.
USE_TC = os.environ.get("DIMO_TC", "1") != "0"
TC_FWD = TC_DGRAD = TC_WGRAD = True      # per-operation switches (bring-up / A-B tests); all on by default
DEBUG_CAPTURE = None                      # list -> forward activations are cloned into it (bring-up only)
# DIMO_TIMENET=layers: round 1's per-layer kernels (operands split inside every GEMM); default: dimo_timenet_fwd / _bwd,
# one C-ABI call per direction over pre-split, TMA-fed operands (csrc/timenet_tc.cu)
USE_CHAIN = os.environ.get("DIMO_TIMENET", "chain") != "layers"


def _untile(ws, off, rows_pad, nkt):
    """split tiles A[rb][kt] (csrc/timenet_tc.cu) -> plain fp32 [rows_pad, 32 * nkt] (hi + lo); tests / debugging"""
    nrb = rows_pad // 128
    t = ws[off:off + nrb * nkt * 32768].view(torch.float32).view(nrb, nkt, 2, 16, 8, 8, 4)   # rb, kt, plane, row group, q, row, e
    t = t.sum(dim=2)
    return t.permute(0, 2, 4, 1, 3, 5).reshape(nrb * 128, nkt * 32)


class _TimeNetChainFn(torch.autograd.Function):
    """Same contract as _TimeNetFn; forward and backward are one C-ABI call each (dimo_timenet_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, pts, times, latents, sink, pts_sink, *params):
        ctx.sink = sink
        ctx.pts_sink = pts_sink            # (tensor, on_done) or None: d(pts) is accumulated into it
        dev = pts.device
        f32 = dict(dtype=torch.float32, device=dev)
        pts = pts.contiguous().float(); times = times.contiguous().float(); latents = latents.contiguous().float()
        params = [p.contiguous().float() for p in params]
        Ws, bs = params[0::2], params[1::2]
        M, G, L = pts.shape[0], times.shape[0], latents.shape[1]
        lib = _lib.lib()
        nbytes = int(lib.dimo_timenet_workspace_bytes(G, M, L))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dxyz = torch.empty(G, M, 3, **f32)
        dquat = torch.empty(G, M, 4, **f32)
        wp = (ctypes.c_void_p * 12)(*[w.data_ptr() for w in Ws])
        bp = (ctypes.c_void_p * 12)(*[b.data_ptr() for b in bs])
        _lib.call("dimo_timenet_fwd", G, M, L, _lib.ptr(pts), _lib.ptr(times), _lib.ptr(latents), wp, bp, _lib.ptr(ws),
                  nbytes, _lib.ptr(dxyz), _lib.ptr(dquat), _lib.stream())
        ctx.save_for_backward(ws, *Ws)
        ctx.dims = (M, G, L, nbytes)
        if DEBUG_CAPTURE is not None:
            lay = (ctypes.c_int64 * 12)()
            lib.dimo_timenet_layout(G, M, L, lay)
            R, E = G * M, 72 + L
            acts = [_untile(ws, int(lay[2 + i]), int(lay[0]), 8)[:R].clone() for i in range(10)]
            h0 = _untile(ws, int(lay[1]), int(lay[0]), 4)[:R, :E]
            cat = torch.cat([h0, acts[SKIP_AFTER]], dim=1)
            DEBUG_CAPTURE.append([cat, acts[8], acts[9]] + [a for i, a in enumerate(acts[:DEPTH]) if i != SKIP_AFTER])
        return dxyz, dquat

    @staticmethod
    def backward(ctx, g_dxyz, g_dquat):
        M, G, L, nbytes = ctx.dims
        ws, *Ws = ctx.saved_tensors
        dev = ws.device
        f32 = dict(dtype=torch.float32, device=dev)
        g_dxyz = (g_dxyz if g_dxyz is not None else torch.zeros(G, M, 3, **f32)).contiguous().float()
        g_dquat = (g_dquat if g_dquat is not None else torch.zeros(G, M, 4, **f32)).contiguous().float()
        sink = ctx.sink
        det = _lib.deterministic()
        if sink is not None and not det:
            dWs, dbs = list(sink[0::2]), list(sink[1::2])
        else:                             # fresh accumulators (int64 in deterministic mode), folded into the sink below
            dWs = [_lib.acc_zeros(W.shape, dev) for W in Ws]
            dbs = [_lib.acc_zeros((W.shape[0],), dev) for W in Ws]
        need_pts, need_lat = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
        pts_sink = ctx.pts_sink if (ctx.pts_sink is not None and need_pts and not det) else None
        dpts = pts_sink[0] if pts_sink else (_lib.acc_zeros((M, 3), dev) if need_pts else None)
        dlat = _lib.acc_zeros((G, L), dev) if need_lat else None
        wp = (ctypes.c_void_p * 12)(*[w.data_ptr() for w in Ws])
        dwp = (ctypes.c_void_p * 12)(*[w.data_ptr() for w in dWs])
        dbp = (ctypes.c_void_p * 12)(*[b.data_ptr() for b in dbs])
        _lib.call("dimo_timenet_bwd", G, M, L, wp, _lib.ptr(ws), nbytes, _lib.ptr(g_dxyz), _lib.ptr(g_dquat), dwp, dbp,
                  _lib.ptr(dpts), _lib.ptr(dlat), _lib.stream())
        if pts_sink:
            dpts = None
            if pts_sink[1] is not None:
                pts_sink[1]()
        else:
            dpts = _lib.acc_result(dpts) if dpts is not None else None
        dlat = _lib.acc_result(dlat) if dlat is not None else None
        if sink is not None:
            if det:
                for acc, dst in zip(dWs + dbs, list(sink[0::2]) + list(sink[1::2])):
                    _lib.acc_result(acc, into=dst)
            return (dpts, None, dlat, None, None, *([None] * (2 * len(Ws))))
        dWs = [_lib.acc_result(a) for a in dWs]
        dbs = [_lib.acc_result(a) for a in dbs]
        grads = []
        for W, b in zip(dWs, dbs):
            grads += [W, b]
        return (dpts, None, dlat, None, None, *grads)


def _linear_fwd(R, K, No, X, ldx, W, b, Y, ldy, relu, s):
    if USE_TC and TC_FWD and K % 4 == 0 and ldx % 4 == 0:
        _lib.call("dimo_linear_tc", R, K, No, X, ldx, None, 0, _lib.ptr(W), _lib.ptr(b), Y, ldy, int(relu), 0, s)
    else:
        _lib.call("dimo_linear_fwd", R, K, No, X, ldx, _lib.ptr(W), _lib.ptr(b), Y, ldy, int(relu), s)


class _TimeNetFn(torch.autograd.Function):
    """(pts [M,3], times [G], latents [G,L], sink, 24 params) -> dxyz [G,M,3], dquat [G,M,4].
    `sink`: None, or the 24 gradient tensors of the parameters (views of the flat gradient buffer, dist.py): the
    backward kernels then accumulate straight into them and autograd sees no parameter gradients (no temporaries,
    no AccumulateGrad add per parameter -- 48 small launches per step)."""

    @staticmethod
    def forward(ctx, pts, times, latents, sink, *params):
        ctx.sink = sink
        dev = pts.device
        f32 = dict(dtype=torch.float32, device=dev)
        pts = pts.contiguous().float()
        times = times.contiguous().float()
        latents = latents.contiguous().float()
        params = [p.contiguous().float() for p in params]
        Ws, bs = params[0::2], params[1::2]
        M, G, L = pts.shape[0], times.shape[0], latents.shape[1]
        R = G * M
        E = 3 * 2 * PTS_FREQS + 2 * TIME_FREQS + L           # 104
        CAT = E + HIDDEN                                       # 360
        s = _lib.stream()
        # activations: cat = [h0 | out of layer 4]; acts[i] = out of layer i (i != 4)
        cat = torch.empty(R, CAT, **f32)
        _lib.call("dimo_timenet_embed_fwd", G, M, L, _lib.ptr(pts), _lib.ptr(times), _lib.ptr(latents),
                  _lib.ptr(cat), CAT, s)
        acts = []
        x_ptr, x_ld, x_k = cat.data_ptr(), CAT, E
        for i in range(DEPTH):
            if i == SKIP_AFTER:
                y = None
                y_ptr, y_ld = cat.data_ptr() + 4 * E, CAT
            else:
                y = torch.empty(R, HIDDEN, **f32)
                y_ptr, y_ld = y.data_ptr(), HIDDEN
            _linear_fwd(R, x_k, HIDDEN, x_ptr, x_ld, Ws[i], bs[i], y_ptr, y_ld, True, s)
            acts.append(y)
            if i == SKIP_AFTER:
                x_ptr, x_ld, x_k = cat.data_ptr(), CAT, CAT
            else:
                x_ptr, x_ld, x_k = y_ptr, y_ld, HIDDEN
        h = acts[DEPTH - 1]
        hp = torch.empty(R, HIDDEN, **f32)
        hr = torch.empty(R, HIDDEN, **f32)
        dxyz = torch.empty(G, M, 3, **f32)
        dquat = torch.empty(G, M, 4, **f32)
        _linear_fwd(R, HIDDEN, HIDDEN, h.data_ptr(), HIDDEN, Ws[8], bs[8], hp.data_ptr(), HIDDEN, True, s)
        _linear_fwd(R, HIDDEN, 3, hp.data_ptr(), HIDDEN, Ws[9], bs[9], dxyz.data_ptr(), 3, False, s)
        _linear_fwd(R, HIDDEN, HIDDEN, h.data_ptr(), HIDDEN, Ws[10], bs[10], hr.data_ptr(), HIDDEN, True, s)
        _linear_fwd(R, HIDDEN, 4, hr.data_ptr(), HIDDEN, Ws[11], bs[11], dquat.data_ptr(), 4, False, s)
        ctx.save_for_backward(pts, times, latents, cat, hp, hr, *[a for a in acts if a is not None], *Ws)
        ctx.dims = (M, G, L, E, CAT)
        if DEBUG_CAPTURE is not None:
            DEBUG_CAPTURE.append([t.clone() for t in (cat, hp, hr, *[a for a in acts if a is not None])])
        return dxyz, dquat

    @staticmethod
    def backward(ctx, g_dxyz, g_dquat):
        M, G, L, E, CAT = ctx.dims
        R = G * M
        saved = ctx.saved_tensors
        pts, times, latents, cat, hp, hr = saved[:6]
        acts_list = list(saved[6:6 + DEPTH - 1])
        Ws = list(saved[6 + DEPTH - 1:])
        acts = acts_list[:SKIP_AFTER] + [None] + acts_list[SKIP_AFTER:]
        dev = pts.device
        f32 = dict(dtype=torch.float32, device=dev)
        s = _lib.stream()
        g_dxyz = g_dxyz.contiguous().float()
        g_dquat = g_dquat.contiguous().float()
        sink = ctx.sink
        if sink is not None:
            dWs, dbs = list(sink[0::2]), list(sink[1::2])
        else:
            dWs = [torch.zeros_like(W) for W in Ws]
            dbs = [torch.zeros(W.shape[0], **f32) for W in Ws]

        deferred = []   # tensor-core weight gradients: independent of each other -> ONE grouped launch at the end
        # W^T for the tensor-core data-gradient GEMMs: one grouped transpose launch into one buffer
        Wts = {}
        if USE_TC and TC_DGRAD:
            tl = [li for li, W in enumerate(Ws) if W.shape[0] % 4 == 0]
            tbuf = torch.empty(sum(Ws[li].numel() for li in tl), **f32)
            o = 0
            for li in tl:
                Wts[li] = tbuf[o:o + Ws[li].numel()]
                o += Ws[li].numel()
            nt = len(tl)
            _lib.call("dimo_transpose_grouped", nt, (ctypes.c_int * nt)(*[Ws[li].shape[0] for li in tl]),
                      (ctypes.c_int * nt)(*[Ws[li].shape[1] for li in tl]),
                      (ctypes.c_void_p * nt)(*[Ws[li].data_ptr() for li in tl]),
                      (ctypes.c_void_p * nt)(*[Wts[li].data_ptr() for li in tl]), s)

        def bwd_layer(li, K, No, dY_ptr, lddy, Y_ptr, ldy, X_ptr, ldx, dX_ptr, lddx, accumulate):
            if USE_TC and TC_WGRAD and No % 4 == 0 and K % 4 == 0 and lddy % 4 == 0 and ldx % 4 == 0 and \
                    (Y_ptr is None or ldy % 4 == 0):
                deferred.append((li, K, No, dY_ptr, lddy, Y_ptr, ldy if Y_ptr is not None else 0, X_ptr, ldx))
            else:
                _lib.call("dimo_linear_bwd_weight", R, K, No, dY_ptr, lddy, Y_ptr, ldy, X_ptr, ldx,
                          _lib.ptr(dWs[li]), _lib.ptr(dbs[li]), s)
            if dX_ptr is not None:
                if USE_TC and TC_DGRAD and No % 4 == 0 and lddy % 4 == 0 and (Y_ptr is None or ldy % 4 == 0):
                    # dX[R,K] = (dY * [Y>0]) [R,No] * W[No,K]: same kernel with the transposed weight as the
                    # "[N_out, K_red]" operand (reduction over No)
                    _lib.call("dimo_linear_tc", R, No, K, dY_ptr, lddy, Y_ptr, ldy if Y_ptr is not None else 0,
                              _lib.ptr(Wts[li]), None, dX_ptr, lddx, 0, int(accumulate), s)
                else:
                    _lib.call("dimo_linear_bwd_data", R, K, No, dY_ptr, lddy, Y_ptr, ldy, _lib.ptr(Ws[li]), dX_ptr,
                              lddx, int(accumulate), s)

        h = acts[DEPTH - 1]
        dhp = torch.empty(R, HIDDEN, **f32)
        dhr = torch.empty(R, HIDDEN, **f32)
        dh = torch.empty(R, HIDDEN, **f32)
        # heads (no ReLU on the outputs)
        bwd_layer(9, HIDDEN, 3, g_dxyz.data_ptr(), 3, None, 0, hp.data_ptr(), HIDDEN, dhp.data_ptr(), HIDDEN, False)
        bwd_layer(11, HIDDEN, 4, g_dquat.data_ptr(), 4, None, 0, hr.data_ptr(), HIDDEN, dhr.data_ptr(), HIDDEN, False)
        bwd_layer(8, HIDDEN, HIDDEN, dhp.data_ptr(), HIDDEN, hp.data_ptr(), HIDDEN, h.data_ptr(), HIDDEN,
                  dh.data_ptr(), HIDDEN, False)
        bwd_layer(10, HIDDEN, HIDDEN, dhr.data_ptr(), HIDDEN, hr.data_ptr(), HIDDEN, h.data_ptr(), HIDDEN,
                  dh.data_ptr(), HIDDEN, True)
        # trunk, layers 7..0 ; dcat collects the skip-concatenated gradient [dh0 | d(out of layer 4)].  Every layer's
        # incoming gradient keeps its own buffer (no ping-pong) because the weight gradients read them at the end.
        dcat = torch.empty(R, CAT, **f32)
        dY_ptr, lddy = dh.data_ptr(), HIDDEN
        grads_alive = [dh]
        for i in range(DEPTH - 1, -1, -1):
            if i == SKIP_AFTER:
                Y_ptr, ldy = cat.data_ptr() + 4 * E, CAT
            else:
                Y_ptr, ldy = acts[i].data_ptr(), HIDDEN
            if i == 0:
                X_ptr, ldx, K = cat.data_ptr(), CAT, E
                dX_ptr, lddx, accumulate = dcat.data_ptr(), CAT, True     # += the skip branch's dh0
            elif i == SKIP_AFTER + 1:
                X_ptr, ldx, K = cat.data_ptr(), CAT, CAT
                dX_ptr, lddx, accumulate = dcat.data_ptr(), CAT, False
            else:
                prev = acts[i - 1]
                if i - 1 == SKIP_AFTER:
                    X_ptr, ldx = cat.data_ptr() + 4 * E, CAT
                else:
                    X_ptr, ldx = prev.data_ptr(), HIDDEN
                K = HIDDEN
                buf = torch.empty(R, HIDDEN, **f32)
                grads_alive.append(buf)
                dX_ptr, lddx, accumulate = buf.data_ptr(), HIDDEN, False
            bwd_layer(i, K, HIDDEN, dY_ptr, lddy, Y_ptr, ldy, X_ptr, ldx, dX_ptr, lddx, accumulate)
            if i == SKIP_AFTER + 1:
                dY_ptr, lddy = dcat.data_ptr() + 4 * E, CAT
            else:
                dY_ptr, lddy = dX_ptr, lddx
        if deferred:
            n = len(deferred)
            ia = lambda vals: (ctypes.c_int * n)(*vals)
            la = lambda vals: (ctypes.c_int64 * n)(*vals)
            pa = lambda vals: (ctypes.c_void_p * n)(*vals)
            _lib.call("dimo_linear_wgrad_tc_grouped", n, R, ia([d[1] for d in deferred]), ia([d[2] for d in deferred]),
                      pa([d[3] for d in deferred]), la([d[4] for d in deferred]), pa([d[5] for d in deferred]),
                      la([d[6] for d in deferred]), pa([d[7] for d in deferred]), la([d[8] for d in deferred]),
                      pa([dWs[d[0]].data_ptr() for d in deferred]), pa([dbs[d[0]].data_ptr() for d in deferred]), s)
        need_pts, need_lat = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
        dpts = torch.zeros(M, 3, **f32) if need_pts else None
        dlat = torch.zeros(G, L, **f32) if need_lat else None
        if need_pts or need_lat:
            _lib.call("dimo_timenet_embed_bwd", G, M, L, _lib.ptr(cat), _lib.ptr(dcat), CAT,
                      _lib.ptr(dpts), _lib.ptr(dlat), s)
        if sink is not None:
            return (dpts, None, dlat, None, *([None] * (2 * len(Ws))))
        grads = []
        for W, b in zip(dWs, dbs):
            grads += [W, b]
        return (dpts, None, dlat, None, *grads)


def _xavier(m):
    # renderer/latent_gs_renderer.py:166-170 (the bias branch re-initialises the weight; bias keeps nn.Linear's default)
    if isinstance(m, nn.Linear):
        init.xavier_uniform_(m.weight, gain=1)


class TimeNet(nn.Module):
    """Same constructor, parameters and outputs as the reference TimeNet (:184-235)."""

    def __init__(self, D=8, W=256, skips=[4], latent_code_dim=32, device="cuda"):
        super().__init__()
        if D != DEPTH or W != HIDDEN or list(skips) != [SKIP_AFTER]:
            raise NotImplementedError("dimo_b200 TimeNet is specialised to D=8, W=256, skips=[4] (the reference's only use)")
        self.pts_ch, self.times_ch = PTS_FREQS, TIME_FREQS
        self.input_ch = 3 * 2 * PTS_FREQS + 2 * TIME_FREQS + latent_code_dim
        self.skips = skips
        self.deformnet = nn.ModuleList(
            [nn.Linear(self.input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)])
        self.pts_layers = nn.Sequential(nn.Linear(W, W), nn.ReLU(), nn.Linear(W, 3))
        self.rot_layers = nn.Sequential(nn.Linear(W, W), nn.ReLU(), nn.Linear(W, 4))
        self.device = device
        # True: backward accumulates weight/bias gradients straight into the parameters' preallocated .grad tensors
        # (set by trainstep.TrainStep, whose gradients are views of one flat buffer); autograd's own accumulation is
        # bypassed for these parameters, so torch.autograd.grad() would not see them -- off by default
        self.direct_grads = False
        self.deformnet.apply(_xavier); self.pts_layers.apply(_xavier); self.rot_layers.apply(_xavier)
        init.constant_(self.pts_layers[-1].weight, 0); init.constant_(self.pts_layers[-1].bias, 0)
        init.constant_(self.rot_layers[-1].weight, 0)
        self.rot_layers[-1].bias.data = torch.tensor([1., 0., 0., 0.])

    def flat_params(self):
        ps = []
        for l in self.deformnet:
            ps += [l.weight, l.bias]
        for l in (self.pts_layers[0], self.pts_layers[2], self.rot_layers[0], self.rot_layers[2]):
            ps += [l.weight, l.bias]
        return ps

    def forward_batched(self, pts, times, latents, pts_direct=None):
        """pts [M,3]; times [G]; latents [G,L] -> dxyz [G,M,3], dquat [G,M,4] (one launch set for all G).
        pts_direct: None, or a callback: with direct_grads on and `pts` a parameter with a preallocated .grad, d(pts) is
        accumulated straight into it and pts_direct([pts]) is called afterwards."""
        ps = self.flat_params()
        sink = None
        if self.direct_grads and torch.is_grad_enabled():
            sink = [p.grad for p in ps]
            if any(g is None or not g.is_contiguous() for g in sink):
                raise RuntimeError("TimeNet.direct_grads needs every parameter's .grad preallocated (FlatGradReducer)")
        L = latents.shape[1]
        if USE_TC and USE_CHAIN and (72 + L) % 4 == 0 and 72 + L <= 128:
            pts_sink = None
            if sink is not None and pts_direct is not None and _lib.grad_sink(pts) is not None and pts.dtype == torch.float32 \
                    and pts.is_contiguous():
                pts_sink = (pts.grad, lambda: pts_direct([pts]))
            return _TimeNetChainFn.apply(pts, times, latents, sink, pts_sink, *ps)
        return _TimeNetFn.apply(pts, times, latents, sink, *ps)

    def forward(self, pts, t, latent_code, nobatch=False, t_apply=False):
        """Reference call forms: (pts [M,3], float t, latent [L]) and the t_apply form
        (pts [1,M,3], t [T,M,1], latent [L]) used by arap_loss_v2 (:1081-1094)."""
        if t_apply:
            p = pts[0] if pts.dim() == 3 else pts
            times = t[:, 0, 0]
            lat = latent_code[None, :].expand(times.shape[0], -1)
            return self.forward_batched(p, times, lat)
        squeeze = pts.dim() == 2
        p = pts if squeeze else pts[0]
        times = torch.tensor([float(t)], dtype=torch.float32).to(p.device, non_blocking=True)
        dxyz, dquat = self.forward_batched(p, times, latent_code[None, :])
        return (dxyz[0], dquat[0]) if squeeze else (dxyz, dquat)

    def get_mlp_parameters(self):
        a, r = [], []
        for name, p in self.named_parameters():
            (r if name.split('.')[0] == "rot_layers" else a).append(p)
        return a, r


class _LBSFn(torch.autograd.Function):
    """(xyz [N,3], rot [N,4], c_xyz [M,3], c_radius_raw [M,1], dxyz [G,M,3], dquat [G,M,4], idx, dist, sink)
    -> means3D [G,N,3], rotations [G,N,4] (normalised).  renderer/latent_gs_renderer.py:1191-1219.
    `sink`: None, or (grad_xyz, grad_rot, grad_c_xyz, grad_c_radius, on_done): tensors (or None) the backward kernel
    accumulates into directly instead of fresh zero buffers (autograd then sees no gradient for those inputs)."""

    @staticmethod
    def forward(ctx, xyz, rot, c_xyz, c_radius_raw, dxyz, dquat, idx, dist, sink=None):
        c = lambda t: t.contiguous().float()
        xyz, rot, c_xyz, c_radius_raw, dxyz, dquat, dist = map(c, (xyz, rot, c_xyz, c_radius_raw, dxyz, dquat, dist))
        idx = idx.contiguous()
        assert idx.dtype == torch.int64
        G, M = dxyz.shape[0], c_xyz.shape[0]
        N, K = idx.shape
        means3D = torch.empty(G, N, 3, dtype=torch.float32, device=xyz.device)
        rotations = torch.empty(G, N, 4, dtype=torch.float32, device=xyz.device)
        _lib.call("dimo_lbs_fwd", G, N, M, K, _lib.ptr(xyz), _lib.ptr(rot), _lib.ptr(idx), _lib.ptr(dist),
                  _lib.ptr(c_xyz), _lib.ptr(c_radius_raw), _lib.ptr(dxyz), _lib.ptr(dquat), _lib.ptr(means3D),
                  _lib.ptr(rotations), _lib.stream())
        ctx.save_for_backward(xyz, rot, c_xyz, c_radius_raw, dxyz, dquat, idx, dist)
        ctx.sink = sink
        return means3D, rotations

    @staticmethod
    def backward(ctx, g_means3D, g_rot):
        xyz, rot, c_xyz, c_radius_raw, dxyz, dquat, idx, dist = ctx.saved_tensors
        G, M = dxyz.shape[0], c_xyz.shape[0]
        N, K = idx.shape
        dev = xyz.device
        sink = ctx.sink if (ctx.sink is not None and not _lib.deterministic()) else (None, None, None, None, None)
        z = lambda t: _lib.acc_zeros(t.shape, t.device)          # int64 accumulators in deterministic mode
        outs = [sk if sk is not None else z(t) for sk, t in zip(sink[:4], (xyz, rot, c_xyz, c_radius_raw))]
        # the two per-frame outputs share one zero fill
        both = _lib.acc_zeros((G * M * 7,), dev)
        d_dxyz, d_dquat = both[G * M * 4:].view(G, M, 3), both[:G * M * 4].view(G, M, 4)
        g_means3D = g_means3D.contiguous().float() if g_means3D is not None else torch.zeros(G, N, 3, device=dev)
        g_rot = g_rot.contiguous().float() if g_rot is not None else torch.zeros(G, N, 4, device=dev)
        _lib.call("dimo_lbs_bwd", G, N, M, K, _lib.ptr(xyz), _lib.ptr(rot), _lib.ptr(idx), _lib.ptr(dist),
                  _lib.ptr(c_xyz), _lib.ptr(c_radius_raw), _lib.ptr(dxyz), _lib.ptr(dquat), _lib.ptr(g_means3D),
                  _lib.ptr(g_rot), _lib.ptr(outs[0]), _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(outs[3]),
                  _lib.ptr(d_dxyz), _lib.ptr(d_dquat), _lib.stream())
        res = [None if sk is not None else _lib.acc_result(o) for sk, o in zip(sink[:4], outs)]
        if both.dtype == torch.int64:
            d_dxyz, d_dquat = _lib.acc_result(d_dxyz.contiguous()), _lib.acc_result(d_dquat.contiguous())
        if sink[4] is not None:
            sink[4]()
        return res[0], res[1], res[2], res[3], d_dxyz, d_dquat, None, None, None


class _GatherRowsFn(torch.autograd.Function):
    """table.index_select(0, index) whose backward adds the row gradients straight into `sink` (table.grad): one
    index_add launch instead of zeros + index_add + AccumulateGrad add."""

    @staticmethod
    def forward(ctx, table, index, sink, on_done):
        ctx.save_for_backward(index)
        ctx.sink, ctx.on_done = sink, on_done
        return table.index_select(0, index)

    @staticmethod
    def backward(ctx, g):
        (index,) = ctx.saved_tensors
        ctx.sink.index_add_(0, index, g)
        if ctx.on_done is not None:
            ctx.on_done()
        return None, None, None, None


def gather_rows(table, index, on_done=None):
    """table[index] (rows); gradients go straight into table.grad when it is preallocated (_lib.grad_sink)."""
    sink = _lib.grad_sink(table) if torch.is_grad_enabled() else None
    if sink is None:
        return table.index_select(0, index)
    return _GatherRowsFn.apply(table, index, sink, on_done)


def lbs_deform(xyz, rot, c_xyz, c_radius_raw, dxyz, dquat, neighbor_indices, neighbor_dists, direct_grads=False,
               on_done=None):
    """Batched stage-s2 skinning; dxyz/dquat may be [M,*] (one frame) or [G,M,*].  direct_grads: the backward kernel
    accumulates the gradients of the four parameters into their preallocated .grad (see _lib.grad_sink) and then calls
    on_done(list of those parameters)."""
    single = dxyz.dim() == 2
    if single:
        dxyz, dquat = dxyz[None], dquat[None]
    sink = None
    if direct_grads and torch.is_grad_enabled():
        ps = (xyz, rot, c_xyz, c_radius_raw)
        sk = [_lib.grad_sink(p) for p in ps]
        if any(t is not None for t in sk):
            done = [p for p, t in zip(ps, sk) if t is not None]
            sink = (*sk, (lambda: on_done(done)) if on_done is not None else None)
    m, r = _LBSFn.apply(xyz, rot, c_xyz, c_radius_raw, dxyz, dquat, neighbor_indices, neighbor_dists, sink)
    return (m[0], r[0]) if single else (m, r)
