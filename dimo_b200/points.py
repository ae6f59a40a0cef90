"""Host side of csrc/points.cu: farthest point sampling, ball query, one-directional chamfer distance -- the three
point-set ops the reference takes from pytorch3d / chamferdist around its step (GUI.FPS main_train_dimo.py:511-515;
ARAP connectivity utils/deform_utils.py:128; key-point chamfer term main_train_dimo.py:298-299)."""
import torch

from . import _lib


def sample_farthest_points(points, K, start=0):
    """points [B,N,3] (CUDA fp32) -> (selected [B,K,3], idx [B,K] int64), pytorch3d.ops.sample_farthest_points order of
    results; the first pick is index `start` (pytorch3d's default random_start_point=False: 0)."""
    if points.dim() != 3 or points.shape[-1] != 3:
        raise ValueError("points must be [B,N,3]")
    pts = points.detach().contiguous().float()
    B, N, _ = pts.shape
    K = int(K)
    idx = torch.empty(B, K, dtype=torch.int64, device=pts.device)
    scratch = torch.empty(B, N, dtype=torch.float32, device=pts.device)
    _lib.call("dimo_fps", B, N, K, int(start), _lib.ptr(pts), _lib.ptr(scratch), _lib.ptr(idx), _lib.stream())
    sel = torch.gather(points, 1, idx[..., None].expand(-1, -1, 3))
    return sel, idx


def ball_query(p1, p2, K=500, radius=0.2, return_nn=True):
    """p1 [B,P1,3], p2 [B,P2,3] -> (dists [B,P1,K] squared, idx [B,P1,K] int64 with -1 padding, nn [B,P1,K,3] or None):
    the first K points of p2 in index order inside the ball (pytorch3d.ops.ball_query)."""
    a = p1.detach().contiguous().float()
    b = p2.detach().contiguous().float()
    B, P1, _ = a.shape
    P2 = b.shape[1]
    idx = torch.empty(B, P1, K, dtype=torch.int64, device=a.device)
    dists = torch.empty(B, P1, K, dtype=torch.float32, device=a.device)
    _lib.call("dimo_ball_query", B, P1, P2, int(K), float(radius), _lib.ptr(a), _lib.ptr(b), _lib.ptr(idx),
              _lib.ptr(dists), _lib.stream())
    nn = None
    if return_nn:
        safe = idx.clamp_min(0)
        nn = torch.gather(p2[:, None].expand(-1, P1, -1, -1), 2, safe[..., None].expand(-1, -1, -1, 3))
        nn = nn * (idx >= 0)[..., None]
    return dists, idx, nn


class _Chamfer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, tgt):
        s = src.contiguous().float()
        t = tgt.contiguous().float()
        N, M = s.shape[0], t.shape[0]
        d2 = torch.empty(N, dtype=torch.float32, device=s.device)
        nn = torch.empty(N, dtype=torch.int32, device=s.device)
        total = torch.zeros((), dtype=torch.float32, device=s.device)
        _lib.call("dimo_chamfer_fwd", N, M, _lib.ptr(s), _lib.ptr(t), _lib.ptr(d2), _lib.ptr(nn), _lib.ptr(total), None,
                  0.0, _lib.stream())
        ctx.save_for_backward(s, t, nn)
        return total

    @staticmethod
    def backward(ctx, g):
        s, t, nn = ctx.saved_tensors
        g = g.contiguous().float()
        d_src = torch.empty_like(s) if ctx.needs_input_grad[0] else None
        d_tgt = torch.zeros_like(t) if ctx.needs_input_grad[1] else None
        _lib.call("dimo_chamfer_bwd", s.shape[0], _lib.ptr(s), _lib.ptr(t), _lib.ptr(nn), _lib.ptr(g), 1.0,
                  _lib.ptr(d_src), _lib.ptr(d_tgt), _lib.stream())
        return d_src, d_tgt


def chamfer_forward(source, target):
    """chamferdist.ChamferDistance()(source[1,N,3], target[1,M,3]) with its defaults: sum over the source points of the
    squared distance to the nearest target point.  Accepts [N,3] or [1,N,3]."""
    s = source[0] if source.dim() == 3 else source
    t = target[0] if target.dim() == 3 else target
    if source.dim() == 3 and source.shape[0] != 1:
        raise NotImplementedError("chamfer_forward: batch size 1 (the DIMO call site)")
    return _Chamfer.apply(s, t)


class ChamferDistance(torch.nn.Module):
    """Call-compatible with chamferdist.ChamferDistance for the reference's use (defaults only)."""

    def forward(self, source_cloud, target_cloud, bidirectional=False, reverse=False, batch_reduction="mean",
                point_reduction="sum"):
        if batch_reduction not in ("mean", "sum", None) or point_reduction != "sum":
            raise NotImplementedError("dimo_b200 ChamferDistance: point_reduction='sum' only (the DIMO call site)")
        if reverse:
            source_cloud, target_cloud = target_cloud, source_cloud
        out = chamfer_forward(source_cloud, target_cloud)
        if bidirectional:
            out = out + chamfer_forward(target_cloud, source_cloud)
        return out
