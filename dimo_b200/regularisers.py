"""Control-point regularisers of the real DIMO step: the ARAP energy over T time samples
(Renderer.arap_loss_v2 renderer/latent_gs_renderer.py:1081-1094 -> utils/deform_utils.py:115-150 connectivity,
:152-232 rotation fit + energy) and the key-point trajectory term (main_train_dimo.py:295-302).

Formulation: a neighbour TABLE nbr [M,K] (-1 padded) instead of the reference's (ii, jj, nn) edge triplets -- every
per-vertex quantity (edge vectors, the 3x3 covariance, the fitted rotation, the energy) is then local to one vertex
and one target frame.

Two implementations of the same arithmetic:
  * CUDA tensors: TWO kernel launches, `dimo_arap_connectivity` + `dimo_arap_energy` (csrc/arap.cu over
    csrc/arap_math.h: thread per vertex / per (frame, vertex), fp64 Jacobi rotation fit, gradient by atomics) --
    the reference spends ~40 launches per call on M = 512 points, once per motion of the batch;
  * any device, `fused=False`: dense [.., M, K, 3] torch expressions batched over all target frames (what the CPU
    tests drive, and the A/B partner of the kernels on the GPU).
Both equal the reference: tests/golden/arap.npz is produced by executing utils/deform_utils.py, and the kernels' source
is checked against it through a host build (tests/test_arap_math_cpu.py).
"""
import numpy as np
import torch

from . import _lib
from . import points as _points


def common_neighbour_table(nn_idx, K=10):
    """nn_idx [T,M,Kq] int64, -1 padded, each list ascending (ball-query order): per frame and vertex a neighbour list.
    Returns (nbr [M,K] int64 with -1 padding: the vertices listed for i in EVERY frame, ascending; count [M]).
    deform_utils.py:131-137 builds the same set through one_hot(...).any(2).all(0) over an [T,M,K,M+1] tensor and a
    top-k over the 0/1 row; here the T lists are intersected directly (O(T M K^2), no M x M intermediate)."""
    T, M, Kq = nn_idx.shape
    cand = nn_idx[0]                                                           # [M,Kq]
    present = (cand[None, :, :, None] == nn_idx[:, :, None, :]).any(dim=-1)    # [T,M,Kq]
    common = present.all(dim=0) & (cand >= 0)
    count = common.sum(dim=1)
    order = torch.argsort((~common).to(torch.int8), dim=1, stable=True)        # common slots first, order kept
    packed = torch.gather(cand, 1, order)
    if Kq < K:
        packed = torch.cat([packed, packed.new_full((M, K - Kq), -1)], dim=1)
    packed = packed[:, :K]
    slot = torch.arange(K, device=nn_idx.device)[None, :]
    nbr = torch.where(slot < count[:, None], packed, torch.full_like(packed, -1))
    return nbr, count


def connectivity_v2(points, K=10, radius=0.1, ball_query=None):
    """points [T,M,3] -> (ii, jj, nn, nbr): the reference's edge triplets (cal_connectivity_from_points_v2) and the
    neighbour table they come from.  The ball query asks for K+1 hits and drops the first one (the reference assumes
    it is the vertex itself, :128-129 -- kept)."""
    bq = ball_query or _points.ball_query
    _d, idx, _nn = bq(points, points, K=K + 1, radius=radius)
    nbr, count = common_neighbour_table(idx[:, :, 1:], K)
    M = nbr.shape[0]
    valid = nbr >= 0
    ii = torch.arange(M, device=nbr.device)[:, None].expand(M, K)[valid]
    nn = torch.arange(K, device=nbr.device)[None, :].expand(M, K)[valid]
    jj = nbr[valid]
    return ii, jj, nn, nbr


def _edges(p, nbr):
    """p [..,M,3], nbr [M,K] -> edge vectors p_i - p_nbr(i,k), zero in the padded slots: [..,M,K,3]."""
    valid = (nbr >= 0)[..., None]
    q = p[..., nbr.clamp_min(0), :]
    return (p[..., :, None, :] - q) * valid


@torch.no_grad()
def fit_rotations(e_src, e_tgt, weight):
    """Per vertex the rotation that best maps the source edge fan onto the target one (Kabsch via a 3x3 SVD with the
    reflection fix, deform_utils.py:152-196).  e_src [M,K,3], e_tgt [F,M,K,3], weight [M,K] -> R [F,M,3,3].
    Vertices whose fan is bit-identical in at least one coordinate keep R = I (the reference zeroes their covariance,
    :175-176)."""
    S = torch.einsum("mka,mk,fmkb->fmab", e_src, weight, e_tgt)
    same = (e_src[None] == e_tgt).all(dim=2).any(dim=-1)                       # [F,M]
    S = torch.where(same[..., None, None], torch.zeros_like(S), S)
    U, sig, Vh = torch.linalg.svd(S)
    V = Vh.transpose(-1, -2)
    R = V @ U.transpose(-1, -2)
    flip = torch.det(R) <= 0
    if bool(flip.any()):
        col = torch.argmin(sig, dim=-1)                                         # [F,M]
        sign = torch.ones_like(sig)
        sign.scatter_(-1, col[..., None], -1.0)
        sign = torch.where(flip[..., None], sign, torch.ones_like(sign))
        R = V @ (U * sign[..., None, :]).transpose(-1, -2)
    return R


def arap_energy(nodes, nbr, weight=None, sample_num=512, sample_idx=None):
    """nodes [T,M,3] (frame 0 = source), nbr [M,K].  sum over frames t >= 1, vertices i and slots k of
    w_ik |(p^t_i - p^t_j) - R^t_i (p^0_i - p^0_j)|^2, rotations fitted without gradient (cal_arap_error,
    deform_utils.py:198-232).  More than `sample_num` vertices: a random sample WITH replacement of sample_num of them
    (np.random.choice, :213) -- pass `sample_idx` to fix it."""
    T, M, _ = nodes.shape
    if weight is None:
        weight = (nbr >= 0).to(nodes.dtype)
    if sample_idx is None and M > sample_num:
        sample_idx = torch.from_numpy(np.random.choice(M, sample_num)).long().to(nodes.device)
    e = _edges(nodes, nbr)                                                       # [T,M,K,3]
    e_src, e_tgt = e[0], e[1:]
    R = fit_rotations(e_src.detach(), e_tgt.detach(), weight)
    rigid = torch.einsum("fmab,mkb->fmka", R, e_src)
    per_vertex = (weight[None] * (e_tgt - rigid).square().sum(-1)).sum(-1)       # [T-1,M]
    if sample_idx is not None:
        per_vertex = per_vertex[:, sample_idx]
    return per_vertex.sum()


class Connectivity:
    """Result of the fused connectivity kernel.  Unpacks like the reference's `(ii, jj, nn, _)` tuple
    (`loss, conns = renderer.arap_loss_v2(...)`), but the edge triplets are only materialised when asked for: boolean
    indexing needs the edge count on the host, and the training loop never looks at them."""

    def __init__(self, nbr, count=None):
        self.nbr, self.count = nbr, count

    def triplets(self):
        M, K = self.nbr.shape
        valid = self.nbr >= 0
        ii = torch.arange(M, device=self.nbr.device)[:, None].expand(M, K)[valid]
        nn = torch.arange(K, device=self.nbr.device)[None, :].expand(M, K)[valid]
        return ii, self.nbr[valid], nn

    def __iter__(self):
        ii, jj, nn = self.triplets()
        return iter((ii, jj, nn, self.nbr))

    def __len__(self):
        return 4


def connectivity_fused(points, K=10, radius=0.1):
    """points [T,M,3] CUDA -> Connectivity (one launch, no host sync)."""
    p = points.detach().contiguous().float()
    T, M, _ = p.shape
    nbr = torch.empty(M, K, dtype=torch.int64, device=p.device)
    count = torch.empty(M, dtype=torch.int32, device=p.device)
    _lib.call("dimo_arap_connectivity", T, M, int(K), float(radius), _lib.ptr(p), _lib.ptr(nbr), _lib.ptr(count),
              _lib.stream())
    return Connectivity(nbr, count)


class _ArapEnergy(torch.autograd.Function):
    """energy (scalar) of nodes [T,M,3] over a neighbour table; the kernel returns the gradient with the forward."""

    @staticmethod
    def forward(ctx, nodes, nbr, mult):
        x = nodes.contiguous().float()
        T, M, _ = x.shape
        energy = torch.empty((), dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        _lib.call("dimo_arap_energy", T, M, nbr.shape[1], _lib.ptr(x), _lib.ptr(nbr), _lib.ptr(mult), _lib.ptr(energy),
                  _lib.ptr(grad), _lib.stream())
        ctx.save_for_backward(grad)
        return energy

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def arap_energy_fused(nodes, nbr, sample_num=512, sample_idx=None):
    """CUDA counterpart of arap_energy (unit edge weights).  Vertex sampling with replacement becomes a per-vertex
    multiplicity."""
    M = nodes.shape[1]
    if sample_idx is None and M > sample_num:
        sample_idx = torch.from_numpy(np.random.choice(M, sample_num)).long()
    mult = None
    if sample_idx is not None:
        mult = torch.bincount(sample_idx.to(nodes.device), minlength=M).float()
    return _ArapEnergy.apply(nodes, nbr.contiguous(), mult)


def arap_loss_points(means3D_t, K=10, radius=0.1, ball_query=None, fused=None):
    """means3D_t [T,M,3] -> (error, (ii, jj, nn, nbr)) -- the tail of Renderer.arap_loss_v2 (:1090-1094).
    fused: None = kernels on CUDA tensors (when no ball_query override is given), torch formulation otherwise."""
    if fused is None:
        fused = means3D_t.is_cuda and ball_query is None
    if fused:
        conn = connectivity_fused(means3D_t, K=K, radius=radius)
        return arap_energy_fused(means3D_t, conn.nbr), conn
    ii, jj, nn, nbr = connectivity_v2(means3D_t.detach(), K=K, radius=radius, ball_query=ball_query)
    return arap_energy(means3D_t, nbr), (ii, jj, nn, nbr)


def keypoint_trajectory_loss(cpts, cpts_ori, chamfer=True, lambda_ga1=10.0, lambda_ga2=10000.0):
    """main_train_dimo.py:295-302: the deformed key points of stage s2 stay near their stage-s1 trajectory."""
    cpts_ori = cpts_ori.detach()
    if chamfer:
        return lambda_ga1 * _points.chamfer_forward(cpts[None], cpts_ori[None])
    return lambda_ga2 * (cpts - cpts_ori).abs().mean()
