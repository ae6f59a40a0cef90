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
