from collections import namedtuple

import torch

from dimo_b200 import points as _points

__all__ = ["sample_farthest_points", "ball_query", "knn_points"]

_KNN = namedtuple("KNN", "dists idx knn")


def sample_farthest_points(points, lengths=None, K=50, random_start_point=False):
    """(selected [B,K,3], idx [B,K]) -- GUI.FPS, main_train_dimo.py:511-515."""
    if lengths is not None or random_start_point or not isinstance(K, int):
        raise NotImplementedError("dimo_b200 pytorch3d shim: full clouds, fixed K, deterministic start (DIMO call site)")
    return _points.sample_farthest_points(points, K)


def ball_query(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, return_nn=True):
    """namedtuple (dists, idx, knn) -- utils/deform_utils.py:128 unpacks it positionally."""
    if lengths1 is not None or lengths2 is not None:
        raise NotImplementedError("dimo_b200 pytorch3d shim: ball_query over full clouds only")
    d, i, nn = _points.ball_query(p1, p2, K=K, radius=radius, return_nn=return_nn)
    return _KNN(dists=d, idx=i, knn=nn)


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False, return_sorted=True):
    """Squared distances + indices of the K nearest p2 points, ascending (the reference's ARAP v1 / geodesic helpers,
    utils/deform_utils.py:49,78 -- off the training path; K <= 8 through dimo_knn, larger K through torch.topk)."""
    if lengths1 is not None or lengths2 is not None or norm != 2:
        raise NotImplementedError("dimo_b200 pytorch3d shim: knn_points over full clouds, L2 only")
    from dimo_b200 import knn as _knn
    ds, ix = [], []
    for a, b in zip(p1, p2):
        if K <= 8:
            d, i = _knn.knn(b, a, K)
            d = d * d
        else:
            d2 = torch.cdist(a, b) ** 2
            d, i = torch.topk(d2, K, dim=1, largest=False)
        ds.append(d); ix.append(i)
    d, i = torch.stack(ds), torch.stack(ix)
    nn = None
    if return_nn:
        nn = torch.gather(p2[:, None].expand(-1, p1.shape[1], -1, -1), 2, i[..., None].expand(-1, -1, -1, 3))
    return _KNN(dists=d, idx=i, knn=nn)
