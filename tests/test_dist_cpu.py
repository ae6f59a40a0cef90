"""CPU, world_size 2, gloo: the N>1 host logic (motion sharding + single flat gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dimo_b200.dist import FlatGradReducer, shard_motions


def test_shard_motions_partition():
    for n, w in [(128, 8), (51, 4), (3, 8), (16, 1)]:
        seen = []
        for r in range(w):
            lo, hi = shard_motions(n, w, r)
            seen += list(range(lo, hi))
        assert seen == list(range(n))
        sizes = [shard_motions(n, w, r)[1] - shard_motions(n, w, r)[0] for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    shapes = [(100, 3), (100, 1, 3), (7,), (4, 32), (256, 104)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    g = torch.Generator().manual_seed(10 + rank)
    local = [torch.randn(s, generator=g) for s in shapes]
    red = FlatGradReducer(params, early=params[:2])          # gradients become views into red.flat
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in params)
    # "backward": autograd accumulates in place into the views; the early hook launches bucket 0 by itself
    lo, hi = shard_motions(4, world, rank)
    loss = 0
    for i, (p, w) in enumerate(zip(params, local)):
        if i == 3:      # latent-code-like parameter: only the rows of the motions this rank owns get gradient
            loss = loss + (p[lo:hi] * w[lo:hi]).sum()
        else:
            loss = loss + (p * w).sum()
    loss.backward()
    n = red.reduce()
    total = sum(p.numel() for p in params)
    assert n == red.flat.numel() and total <= n < total + 4 * len(params)      # one buffer (16-byte aligned tensors)
    assert all(o % 4 == 0 for o, _ in red.offsets) and red.n_early % 4 == 0
    for p in params:
        assert p.grad.untyped_storage().data_ptr() == red.flat.untyped_storage().data_ptr()
    q.put((rank, [p.grad.numpy().copy() for p in params], [l.numpy().copy() for l in local]))
    red.zero()
    assert float(red.flat.abs().max()) == 0.0 and float(params[0].grad.abs().max()) == 0.0
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, g0, l0), (r1, g1, l1) = [(r, [torch.tensor(x) for x in g], [torch.tensor(x) for x in l]) for r, g, l in res]
    for i, (a, b) in enumerate(zip(g0, g1)):
        assert torch.equal(a, b), "ranks disagree after the all-reduce"
        if i == 3:
            want = torch.zeros_like(a)
            for r, l in ((0, l0), (1, l1)):
                lo, hi = shard_motions(4, world, r)
                want[lo:hi] = l[3][lo:hi]
            assert torch.allclose(a, want)
        else:
            assert torch.allclose(a, l0[i] + l1[i])


# ---------------------------------------------------------------------------------------------------------------
# rank-identical densification (replicated Gaussians, per-rank statistics)
# ---------------------------------------------------------------------------------------------------------------
def _densify_model():
    import types
    from dimo_b200 import synthetic
    from dimo_b200.gaussian_model import GaussianModel
    import model_scenario as ms
    torch.manual_seed(3)
    g = GaussianModel(0, num_latent_code=2, device="cpu")
    g.load_state(synthetic.make_scene(500, n_ctrl=16, n_motions=2, seed=4))
    g.spatial_lr_scale = 1
    g.training_setup(ms.train_args(), optimizer="torch")
    gen = torch.Generator().manual_seed(1)
    for _ in range(2):                                    # replicated parameters: identical (all-reduced) gradients
        for grp in g.optimizer.param_groups:
            for p in grp["params"]:
                p.grad = 0.01 * torch.randn(p.shape, generator=gen)
        g.optimizer.step()
        g.optimizer.zero_grad()
    with torch.no_grad():
        g._scaling.data.copy_(torch.log(torch.rand(500, 3, generator=gen) * 0.08 + 0.005))
    return g, types


def _feed_stats(g, types, rank):
    gen = torch.Generator().manual_seed(100 + rank)       # every rank saw other frames
    n = g._xyz.shape[0]
    for _ in range(3):
        vs = types.SimpleNamespace(grad=0.03 * torch.randn(n, 3, generator=gen))
        vis = torch.rand(n, generator=gen) > 0.4
        radii = torch.randint(0, 3, (n,), generator=gen).float()
        g.max_radii2D[vis] = torch.max(g.max_radii2D[vis], radii[vis])
        g.add_densification_stats(vs, vis)


def _densify_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, types = _densify_model()
    _feed_stats(g, types, rank)
    torch.manual_seed(77 + rank)                          # ranks' own streams differ: the split must not depend on them
    g.densify_and_prune(0.02, min_opacity=0.1, extent=4, max_screen_size=1)
    with torch.no_grad():
        g._opacity.data[::9] = -8.0
    g.max_radii2D[rank::5] = 3.0                          # rank-local screen-size statistics
    g.prune(min_opacity=0.01, extent=4, max_screen_size=2)
    st = g.optimizer.state[g._xyz]
    q.put((rank, g._xyz.detach().numpy().copy(), g._scaling.detach().numpy().copy(), st["exp_avg"].numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_densify_and_prune_is_rank_identical():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_densify_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, x0, s0, m0), (_, x1, s1, m1) = res
    assert x0.shape == x1.shape and x0.shape[0] != 500
    assert (x0 == x1).all() and (s0 == s1).all() and (m0 == m1).all(), "ranks diverged"
    # ... and equal to ONE process that saw both ranks' frames (same split seed: rank 0's draw after seeding 77)
    g, types = _densify_model()
    _feed_stats(g, types, 0)
    _feed_stats(g, types, 1)
    torch.manual_seed(77)
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64))
    g.split_generator = torch.Generator().manual_seed(seed)
    g.densify_and_prune(0.02, min_opacity=0.1, extent=4, max_screen_size=1)
    with torch.no_grad():
        g._opacity.data[::9] = -8.0
    g.max_radii2D[0::5] = 3.0
    g.max_radii2D[1::5] = 3.0
    g.prune(min_opacity=0.01, extent=4, max_screen_size=2)
    assert g._xyz.shape[0] == x0.shape[0]
    assert (g._xyz.detach().numpy() == x0).all()


def _worker_direct(rank, world, port, q):
    """Gradient sinks (trainstep direct_grads): 'kernels' write the gradients straight into the .grad views, no
    AccumulateGrad node runs, the op reports through direct_written().  The early bucket must be launched exactly when
    its last parameter has been reported, the overflow word of the tail must ride in the late bucket, and the result
    must equal the sum over ranks."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = [(50, 3), (50, 4), (9, 3), (4, 8)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    red = FlatGradReducer(params, early=params[:2])
    g = torch.Generator().manual_seed(20 + rank)
    local = [torch.randn(s, generator=g) for s in shapes]
    launched = []
    orig = red._launch
    red._launch = lambda lo, hi: (launched.append((lo, hi)), orig(lo, hi))[1]
    with torch.no_grad():
        params[0].grad.add_(local[0])                 # "LBS backward": first early parameter
        red.direct_written([params[0]])
        assert launched == [] and red.dirty
        params[2].grad.add_(local[2])                 # a late parameter in between
        red.direct_written([params[2]])
        assert launched == []
        params[1].grad.add_(local[1])                 # last early parameter -> bucket 0 goes out now
        red.direct_written([params[1]])
        assert launched == [(0, red.n_early)]
        params[3].grad.add_(local[3])
        red.direct_written([params[3]])
        red.tail[0] = float(rank == 1)                # "this rank overflowed"
    red.reduce()
    assert launched[-1] == (red.n_early, red.n + 4) and len(launched) == 2
    q.put((rank, [p.grad.numpy().copy() for p in params], [l.numpy().copy() for l in local], float(red.tail[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_direct_written_bookkeeping_two_ranks():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_direct, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, g0, l0, t0), (_, g1, l1, t1) = res
    assert t0 == t1 == 1.0                            # number of ranks that overflowed, identical everywhere
    for a, b, x, y in zip(g0, g1, l0, l1):
        assert (a == b).all() and torch.allclose(torch.tensor(a), torch.tensor(x) + torch.tensor(y))
