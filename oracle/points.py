"""Oracle: point-set ops around the step (FPS, ball query, chamfer) and the ARAP energy.
TEST INFRASTRUCTURE (see oracle/__init__.py).

* ``arap_connectivity_v2`` / ``arap_error`` restate utils/deform_utils.py:115-150 and :152-232 in the reference's own
  shape (edge triplets ii/jj/nn, one rotation fit per frame) and are PINNED: tests/golden/arap.npz is produced by
  tests/golden/make_golden_model.py executing the reference functions (with pytorch3d.ops.ball_query served by
  ``ball_query`` below).  The product (dimo_b200/regularisers.py) uses a different formulation (neighbour table,
  all frames batched) and is checked against both.
* PARITY UNPINNED: ``fps`` (pytorch3d.ops.sample_farthest_points), ``ball_query`` (pytorch3d.ops.ball_query) and
  ``chamfer_forward`` (chamferdist.ChamferDistance, unpinned in requirements.txt) -- pip packages absent from
  /root/reference and from this image.  Published behaviour restated: FPS starts at index 0, keeps per point the
  squared distance to the nearest picked point and picks the arg-max (first maximum); ball_query returns the first K
  points in index order with squared distance < radius^2 (idx padded with -1, dists with 0); ChamferDistance()(a, b)
  with default arguments is sum_i min_j |a_i - b_j|^2.

Squared distances are the fp32 sequence (dx*dx + dy*dy) + dz*dz, separately rounded (csrc/points.cu is compiled with
-fmad=false), so indices are bit-exact.
"""
import numpy as np
import torch


def _sqdist(q, r):
    dx = q[:, None, 0] - r[None, :, 0]
    dy = q[:, None, 1] - r[None, :, 1]
    dz = q[:, None, 2] - r[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def fps(points, K, start=0):
    """points [N,3] fp32 tensor -> idx [K] int64 (main_train_dimo.py:511-515 uses idxs[0] of the batched call)."""
    p = points.detach().cpu().float().numpy()
    N = p.shape[0]
    mind = np.full(N, np.inf, dtype=np.float32)
    out = np.empty(K, dtype=np.int64)
    sel = start
    out[0] = sel
    for k in range(1, K):
        d = p - p[sel]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        mind = np.minimum(mind, d2.astype(np.float32))
        sel = int(np.argmax(mind))                # first maximum
        out[k] = sel
    return torch.from_numpy(out)


def ball_query(p1, p2, K=500, radius=0.2, return_nn=True):
    """p1 [B,P1,3], p2 [B,P2,3] -> (dists [B,P1,K], idx [B,P1,K] int64, nn [B,P1,K,3] | None)."""
    B, P1, _ = p1.shape
    r2 = np.float32(radius) * np.float32(radius)
    dists = torch.zeros(B, P1, K, dtype=torch.float32)
    idx = torch.full((B, P1, K), -1, dtype=torch.int64)
    for b in range(B):
        d2 = _sqdist(p1[b].detach().float(), p2[b].detach().float())
        inside = d2 < float(r2)
        rank = torch.cumsum(inside.long(), dim=1) - 1                    # slot of every hit, index order
        take = inside & (rank < K)
        rows, cols = torch.nonzero(take, as_tuple=True)
        idx[b, rows, rank[rows, cols]] = cols
        dists[b, rows, rank[rows, cols]] = d2[rows, cols]
    nn = None
    if return_nn:
        nn = torch.zeros(B, P1, K, 3, dtype=p2.dtype)
        for b in range(B):
            m = idx[b] >= 0
            nn[b][m] = p2[b][idx[b][m]]
    return dists, idx, nn


def chamfer_forward(src, tgt):
    """[N,3], [M,3] -> sum_i min_j |src_i - tgt_j|^2, differentiable in both (gradient through the arg-min pairs)."""
    d2 = _sqdist(src.detach().float(), tgt.detach().float())
    nn = torch.argmin(d2, dim=1)
    diff = src - tgt[nn]
    return ((diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]).sum()


# ---------------------------------------------------------------------------------------------------------------
# ARAP (reference shape: edge triplets, per-frame loop)
# ---------------------------------------------------------------------------------------------------------------
def arap_connectivity_v2(points, K=10, radius=0.1):
    """utils/deform_utils.py:115-150.  points [T,Nv,3] -> ii, jj, nn (edges present in every frame's ball query)."""
    T, Nv, _ = points.shape
    _d, nn_idx, _n = ball_query(points, points, K=K + 1, radius=radius)
    nn_idx = nn_idx[:, :, 1:]                                                    # :129 drops the first hit
    hot = torch.nn.functional.one_hot(nn_idx + 1, num_classes=Nv + 1).to(torch.bool)      # :131
    member = hot.any(dim=2).all(dim=0).to(torch.float)                           # :132
    member[:, 0] = 0.0                                                           # the -1 padding class
    num = member.sum(dim=1).to(torch.uint8)
    _, cols = torch.topk(member, k=K, dim=1, largest=True)                        # :136
    cols = (cols - 1).abs()
    ii = torch.arange(Nv)[:, None].expand(Nv, K)
    nn = torch.arange(K)[None].expand(Nv, K)
    mask = torch.arange(K).expand_as(cols) < num[:, None]
    return ii[mask], cols[mask], nn[mask]


def _edge_matrix(verts, shape, ii, jj, nn):
    E = torch.zeros(shape, dtype=verts.dtype)
    E[ii, nn] = verts[ii] - verts[jj]
    return E


def arap_error(nodes, ii, jj, nn, K=10, sample_idx=None):
    """utils/deform_utils.py:198-232 with weight=None (unit weights on the listed edges).  nodes [Nt,Nv,3]."""
    Nt, Nv, _ = nodes.shape
    weight = torch.zeros(Nv, K, dtype=nodes.dtype)
    weight[ii, nn] = 1
    if sample_idx is None:
        sample_idx = torch.arange(Nv)
    src = _edge_matrix(nodes[0], (Nv, K, 3), ii, jj, nn)[sample_idx]
    w = weight[sample_idx]
    total = 0
    for t in range(1, Nt):
        tgt = _edge_matrix(nodes[t], (Nv, K, 3), ii, jj, nn)[sample_idx]
        with torch.no_grad():
            S = torch.bmm(src.permute(0, 2, 1), torch.bmm(torch.diag_embed(w), tgt))
            unchanged = torch.unique(torch.where((src == tgt).all(dim=1))[0])
            S[unchanged] = 0
            U, sig, W = torch.svd(S)
            R = torch.bmm(W, U.permute(0, 2, 1))
            bad = torch.nonzero(torch.det(R) <= 0, as_tuple=False).flatten()
            if len(bad) > 0:
                Um = U.clone()
                Um[bad, :, torch.argmin(sig[bad], dim=1)] *= -1
                R[bad] = torch.bmm(W[bad], Um[bad].permute(0, 2, 1))
        rigid = torch.bmm(R, src.permute(0, 2, 1)).permute(0, 2, 1)
        total = total + (w * (torch.norm(tgt - rigid, dim=2) ** 2)).sum()
    return total
