"""The oracles whose upstream sources are absent (KNN_CUDA, simple-knn, pytorch3d ball query, chamferdist: "parity
unpinned") cross-checked against INDEPENDENT library implementations of the same definitions -- scikit-learn's
NearestNeighbors and SciPy's cKDTree (float64 trees; comparisons allow for fp32 distance rounding at ties)."""
import numpy as np
import torch
from scipy.spatial import cKDTree
from sklearn.neighbors import NearestNeighbors

from oracle import knn as oknn
from oracle import points as op


def _cloud(n, seed, scale=1.0):
    return (torch.rand(n, 3, generator=torch.Generator().manual_seed(seed)) - 0.5) * scale


def test_knn_against_sklearn():
    ref, qry = _cloud(512, 0), _cloud(3000, 1)
    dist, idx = oknn.knn(ref, qry, 4)
    nn = NearestNeighbors(n_neighbors=4, algorithm="brute").fit(ref.numpy().astype(np.float64))
    d_ref, i_ref = nn.kneighbors(qry.numpy().astype(np.float64))
    assert np.abs(dist.numpy() - d_ref).max() <= 1e-6                      # Euclidean (unsquared), ascending
    same = (idx.numpy() == i_ref)
    assert same.mean() > 0.999                                              # only exact fp32 ties may order differently
    rows = np.where(~same.all(axis=1))[0]
    for r in rows:
        assert np.abs(np.sort(d_ref[r]) - np.sort(dist.numpy()[r])).max() <= 1e-6


def test_dist3nn_against_kdtree():
    pts = _cloud(2000, 2)
    got = oknn.dist3nn(pts).numpy()
    d, _ = cKDTree(pts.numpy().astype(np.float64)).query(pts.numpy().astype(np.float64), k=4)    # self + 3 neighbours
    want = (d[:, 1:] ** 2).mean(axis=1)
    assert np.abs(got - want).max() <= 1e-6 * want.max()


def test_ball_query_against_kdtree():
    p = _cloud(400, 3, scale=0.6)
    radius = 0.1
    _d, idx, _nn = op.ball_query(p[None], p[None], K=64, radius=radius)     # K large: nothing is truncated
    tree = cKDTree(p.numpy().astype(np.float64))
    d2 = ((p[:, None].double() - p[None].double()) ** 2).sum(-1).numpy()
    for i in range(p.shape[0]):
        got = [j for j in idx[0, i].tolist() if j >= 0]
        ball = sorted(tree.query_ball_point(p[i].numpy().astype(np.float64), radius))
        edge = {j for j in set(got) ^ set(ball) if abs(d2[i, j] - radius ** 2) <= 1e-7}   # on the surface: rounding decides
        assert set(got) ^ set(ball) == edge, i
        assert got == sorted(got)                                           # index order (pytorch3d), not distance order


def test_chamfer_against_kdtree():
    a, b = _cloud(700, 4), _cloud(512, 5)
    got = op.chamfer_forward(a, b).item()
    d, _ = cKDTree(b.numpy().astype(np.float64)).query(a.numpy().astype(np.float64), k=1)
    assert abs(got - float((d ** 2).sum())) <= 1e-5 * got
