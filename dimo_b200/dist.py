"""Multi-GPU plumbing for the path: motion sharding + ONE flat gradient all-reduce per step.

The reference is single-GPU; frames (motion, t, view) are independent given the shared
parameters, so ranks own disjoint blocks of motions and the only exchange is the gradient sum at the
optimizer step.  torch.distributed (NCCL over NVLink/NVSwitch on the box, gloo in the CPU tests)
is the transport; there is no data-path collective inside any kernel.
"""
import torch


def shard_motions(n_motions, world, rank):
    """Block partition of motion (latent) indices: rank r owns [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(n_motions, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class FlatGradReducer:
    """Packs every gradient into one persistent fp32 buffer, all-reduces it once (SUM), unpacks.
    Gradients of latent codes owned by other ranks are zero locally, so SUM gives every rank the full update."""

    def __init__(self, params):
        self.params = [p for p in params if p.numel() > 0]
        self.flat = None

    def layout(self):
        o, out = 0, []
        for p in self.params:
            out.append((o, p.numel()))
            o += p.numel()
        return out, o

    def reduce(self, group=None):
        import torch.distributed as dist
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        lay, n = self.layout()
        dev = grads[0].device
        if self.flat is None or self.flat.numel() != n or self.flat.device != dev:
            self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        views = [self.flat[o:o + k].view_as(g) for (o, k), g in zip(lay, grads)]
        torch._foreach_copy_(views, grads)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        for p, v in zip(self.params, views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return n
