"""Drop-in for `knn_cuda` (unlimblue/KNN_CUDA @ 619617b5): KNN(k, transpose_mode)(ref, query) ->
(dist, idx), call site main_train_dimo.py:502-509 / main_test_dimo.py:157-164."""
import torch
import torch.nn as nn

from dimo_b200 import knn as _knn

__all__ = ["KNN"]


class KNN(nn.Module):
    def __init__(self, k, transpose_mode=False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref, query):
        assert ref.size(0) == query.size(0), "ref.shape={} != query.shape={}".format(ref.shape, query.shape)
        with torch.no_grad():
            D, I = [], []
            for r, q in zip(ref, query):
                if not self._t:                 # [dim, n] layout -> [n, dim]
                    r, q = r.t(), q.t()
                d, i = _knn.knn(r, q, self.k)   # [nq, k]
                if not self._t:
                    d, i = d.t(), i.t()
                D.append(d); I.append(i)
            return torch.stack(D, dim=0), torch.stack(I, dim=0)
