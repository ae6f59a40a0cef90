"""The per-vertex ARAP arithmetic the CUDA kernels run (dimo_b200/csrc/arap_math.h), compiled for the host with g++ and
checked on the CPU against (a) the fixture produced by the reference's own functions (tests/golden/arap.npz) and
(b) the torch formulation in dimo_b200/regularisers.py.  Only the launch glue of csrc/arap.cu is left to the GPU tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from dimo_b200 import regularisers
from oracle import points as opoints

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("arap") / "arap_host.so")
    src = os.path.join(HERE, "cpu_harness", "arap_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src], check=True)
    L = ctypes.CDLL(so)
    vp = ctypes.c_void_p
    L.arap_energy_host.argtypes = [ctypes.c_int] * 3 + [vp] * 5
    L.arap_connectivity_host.argtypes = [ctypes.c_int] * 4 + [ctypes.c_float] + [vp] * 3
    L.arap_rotation_host.argtypes = [vp, vp]
    return L


def _energy(L, nodes, nbr, mult=None):
    nodes = np.ascontiguousarray(nodes, dtype=np.float32)
    nbr = np.ascontiguousarray(nbr, dtype=np.int64)
    T, M, _ = nodes.shape
    e = ctypes.c_double(0.0)
    grad = np.empty_like(nodes)
    m = None if mult is None else np.ascontiguousarray(mult, dtype=np.float32)
    L.arap_energy_host(T, M, nbr.shape[1], nodes.ctypes.data, nbr.ctypes.data, None if m is None else m.ctypes.data,
                       ctypes.addressof(e), grad.ctypes.data)
    return e.value, grad


def _connectivity(L, nodes, K=10, radius=0.1):
    nodes = np.ascontiguousarray(nodes, dtype=np.float32)
    T, M, _ = nodes.shape
    nbr = np.empty((M, K), dtype=np.int64)
    cnt = np.empty(M, dtype=np.int32)
    L.arap_connectivity_host(T, M, K + 1, K, radius, nodes.ctypes.data, nbr.ctypes.data, cnt.ctypes.data)
    return nbr, cnt


@pytest.mark.parametrize("tag", ["a", "b"])
def test_host_build_matches_reference_fixture(host, tag):
    gold = np.load(os.path.join(GOLD, "arap.npz"))
    nodes = gold[f"{tag}/nodes"]
    nbr, cnt = _connectivity(host, nodes)
    edges = sorted((i, int(j)) for i in range(nbr.shape[0]) for j in nbr[i] if j >= 0)
    assert edges == sorted(zip(gold[f"{tag}/ii"].tolist(), gold[f"{tag}/jj"].tolist()))
    assert cnt.sum() == len(edges)
    e, grad = _energy(host, nodes, nbr)
    want = float(gold[f"{tag}/error"])
    assert abs(e - want) <= 2e-5 * want
    g_ref = gold[f"{tag}/grad"]
    assert np.abs(grad - g_ref).max() <= 1e-4 * np.abs(g_ref).max()


def test_host_build_matches_torch_formulation(host):
    g = torch.Generator().manual_seed(8)
    for T, M, spread in ((8, 512, 0.45), (3, 40, 0.15), (2, 7, 0.05)):
        base = (torch.rand(M, 3, generator=g) - 0.5) * 2 * spread
        nodes = torch.stack([base + 0.004 * t * torch.randn(M, 3, generator=g) for t in range(T)])
        nodes[-1] = nodes[0]                                     # a frame identical to the source: R = I, zero energy
        nodes_t = nodes.clone().requires_grad_(True)
        err, (_ii, _jj, _nn, nbr_t) = regularisers.arap_loss_points(nodes_t, ball_query=opoints.ball_query)
        (gr,) = torch.autograd.grad(err, nodes_t)
        nbr, _cnt = _connectivity(host, nodes.numpy())
        assert np.array_equal(nbr, nbr_t.numpy())
        e, grad = _energy(host, nodes.numpy(), nbr)
        assert abs(e - err.item()) <= 2e-5 * max(err.item(), 1e-12)
        assert np.abs(grad - gr.numpy()).max() <= 1e-4 * max(float(gr.abs().max()), 1e-12)
        # sampling with replacement = per-vertex multiplicities
        idx = torch.randint(0, M, (M // 2,), generator=g)
        mult = torch.bincount(idx, minlength=M).float()
        e2, _ = _energy(host, nodes.numpy(), nbr, mult.numpy())
        want2 = regularisers.arap_energy(nodes, nbr_t, sample_idx=idx).item()
        assert abs(e2 - want2) <= 2e-5 * max(want2, 1e-12)


def test_rotation_fit_against_svd(host):
    """Kabsch rotation from a 3x3 covariance: full rank (det > 0 and det < 0), rank 2, rank 1, zero."""
    g = torch.Generator().manual_seed(9)

    def fit(S):
        S = np.ascontiguousarray(S, dtype=np.float64)
        R = np.empty((3, 3))
        host.arap_rotation_host(S.ctypes.data, R.ctypes.data)
        return R

    def kabsch(S):                       # deform_utils.py:179-191 in float64
        U, sig, Vt = np.linalg.svd(S)
        R = Vt.T @ U.T
        if np.linalg.det(R) <= 0:
            U[:, np.argmin(sig)] *= -1
            R = Vt.T @ U.T
        return R

    for k in range(200):
        S = torch.randn(3, 3, generator=g, dtype=torch.float64).numpy() * 10.0 ** ((k % 7) - 4)
        if k % 3 == 0:
            S[:, 0] *= -1                                        # mix of det signs
        R = fit(S)
        assert abs(np.linalg.det(R) - 1) < 1e-9 and np.abs(R @ R.T - np.eye(3)).max() < 1e-9
        assert np.abs(R - kabsch(S)).max() < 1e-7, k
    for k in range(50):                                           # rank 2: two edges
        a, b = torch.randn(2, 3, generator=g, dtype=torch.float64).numpy(), torch.randn(2, 3, generator=g,
                                                                                         dtype=torch.float64).numpy()
        S = a.T @ b
        R = fit(S)
        assert abs(np.linalg.det(R) - 1) < 1e-9
        assert np.abs(R - kabsch(S)).max() < 1e-6, k
    for k in range(20):                                           # rank 1: R maps the source direction onto the target's
        a, b = torch.randn(3, generator=g, dtype=torch.float64).numpy(), torch.randn(3, generator=g,
                                                                                     dtype=torch.float64).numpy()
        R = fit(np.outer(a, b))
        assert abs(np.linalg.det(R) - 1) < 1e-9
        assert np.abs(R @ (a / np.linalg.norm(a)) - b / np.linalg.norm(b)).max() < 1e-7
    assert np.array_equal(fit(np.zeros((3, 3))), np.eye(3))


def test_connectivity_host_build_vs_reference_shaped_oracle_random(host):
    """Random clouds, densities from 'almost no neighbours' to 'more than K in the ball' (lists truncated at K + 1 hits),
    T from 1 to 6: the kernels' connectivity source (host build) against the oracle written in the reference's own
    shape (one_hot / any / all / topk, oracle/points.py::arap_connectivity_v2)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True, database=None)
    @given(st.integers(1, 6), st.integers(12, 90), st.floats(0.08, 0.5), st.integers(0, 2 ** 31 - 1))
    def check(T, M, spread, seed):
        g = torch.Generator().manual_seed(seed)
        base = (torch.rand(M, 3, generator=g) - 0.5) * 2 * spread
        nodes = torch.stack([base + 0.01 * torch.randn(M, 3, generator=g) for _ in range(T)])
        nbr, cnt = _connectivity(host, nodes.numpy())
        ii, jj, _nn = opoints.arap_connectivity_v2(nodes)
        want = sorted(zip(ii.tolist(), jj.tolist()))
        got = sorted((i, int(j)) for i in range(M) for j in nbr[i] if j >= 0)
        assert got == want
        assert all(list(nbr[i][:cnt[i]]) == sorted(nbr[i][:cnt[i]]) and (nbr[i][cnt[i]:] == -1).all() for i in range(M))

    check()


def test_energy_host_build_vs_reference_shaped_oracle_random(host):
    """Energy and gradient of the kernels' source against the oracle in the reference's shape (edge triplets, one SVD
    per frame, torch.svd + reflection fix) on random clouds."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None, derandomize=True, database=None)
    @given(st.integers(2, 5), st.integers(12, 70), st.floats(0.1, 0.3), st.integers(0, 2 ** 31 - 1))
    def check(T, M, spread, seed):
        g = torch.Generator().manual_seed(seed)
        base = (torch.rand(M, 3, generator=g) - 0.5) * 2 * spread
        nodes = torch.stack([base + 0.01 * t * torch.randn(M, 3, generator=g) for t in range(T)]).requires_grad_(True)
        ii, jj, nn = opoints.arap_connectivity_v2(nodes.detach())
        want = opoints.arap_error(nodes, ii, jj, nn)
        nbr, _ = _connectivity(host, nodes.detach().numpy())
        e, grad = _energy(host, nodes.detach().numpy(), nbr)
        scale = max(want.item(), 1e-10)
        assert abs(e - want.item()) <= 1e-4 * scale
        if len(ii) and want.item() > 1e-8:
            (gw,) = torch.autograd.grad(want, nodes)
            assert np.abs(grad - gw.numpy()).max() <= 2e-3 * float(gw.abs().max())

    check()
