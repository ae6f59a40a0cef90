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
    for p, gr in zip(params, local):
        p.grad = gr.clone()
    # latent-code-like parameter: each rank only has gradient rows for the motions it owns
    lo, hi = shard_motions(4, world, rank)
    params[3].grad.zero_(); params[3].grad[lo:hi] = local[3][lo:hi]
    red = FlatGradReducer(params)
    n = red.reduce()
    assert n == sum(p.numel() for p in params)
    assert red.flat.numel() == n                      # one buffer, one collective
    q.put((rank, [p.grad.numpy().copy() for p in params], [l.numpy().copy() for l in local]))
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
