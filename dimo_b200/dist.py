"""Multi-GPU plumbing for the path: motion sharding + ONE flat gradient buffer all-reduced per step.

The reference is single-GPU; frames (motion, t, view) are independent given the shared
parameters, so ranks own disjoint blocks of motions and the only exchange is the gradient sum at the
optimizer step.  torch.distributed (NCCL over NVLink/NVSwitch on the box, gloo in the CPU tests)
is the transport; there is no data-path collective inside any kernel.

Design: every parameter's ``.grad`` is a VIEW into one persistent fp32 buffer, so autograd accumulates straight
into the communication buffer (no pack / unpack copies).  The buffer has two contiguous buckets:

  bucket 0  "early"  gradients that are final before the deformation-MLP backward runs (the per-Gaussian
            parameters: 14 floats per Gaussian) -- its all-reduce is launched from a post-accumulate hook and
            overlaps the TimeNet backward;
  bucket 1  everything else (control points, latent codes, TimeNet weights), reduced after backward.
"""
import torch

ALIGN = 4      # floats
TAIL = 4       # status floats behind the gradients (see FlatGradReducer)


def shard_motions(n_motions, world, rank):
    """Block partition of motion (latent) indices: rank r owns [lo, hi).  Sizes differ by at most one."""
    base, rem = divmod(n_motions, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


class FlatGradReducer:
    """Gradients of `params` live in one flat buffer (`.flat`); `reduce()` sums it over the process group.
    Gradients of latent codes owned by other ranks are zero locally, so SUM gives every rank the full update."""

    def __init__(self, params, early=(), group=None, attach=True):
        early_ids = {id(p) for p in early}
        self.early = [p for p in params if p.numel() > 0 and id(p) in early_ids]
        self.late = [p for p in params if p.numel() > 0 and id(p) not in early_ids]
        self.params = self.early + self.late
        self.group = group
        # every tensor starts on a 16-byte boundary (128-bit accesses in the fused optimizer, optim.FusedAdam);
        # the padding floats stay zero and ride along in the all-reduce
        al = lambda x: (x + ALIGN - 1) // ALIGN * ALIGN
        self.offsets, o = [], 0
        for i, p in enumerate(self.params):
            self.offsets.append((o, p.numel()))
            o = al(o + p.numel())
            if i == len(self.early) - 1:
                self.n_early = o
        if not self.early:
            self.n_early = 0
        self.n = o
        # TAIL floats behind the gradients ride along in the late bucket's all-reduce: [0] = "instance capacity
        # overflowed on this rank" (trainstep.TrainStep writes it every step; after the SUM it is the number of ranks
        # that dropped work, identical everywhere, and gates the optimizer update on every rank alike)
        self.flat = None
        self.flat_all = None
        self.tail = None
        self.dirty = False          # gradients accumulated since the optimizer last consumed them
        self._arrived = 0
        self._work = []
        self._hooks = []
        if attach and self.params:
            self.attach()

    # ------------------------------------------------------------------------------------------
    def attach(self):
        dev = self.params[0].device
        self.flat_all = torch.zeros(self.n + TAIL, dtype=torch.float32, device=dev)
        self.flat = self.flat_all[:self.n]
        self.tail = self.flat_all[self.n:]
        for p, (o, numel) in zip(self.params, self.offsets):
            v = self.flat[o:o + numel].view_as(p)
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v
        for p in self.early:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_early_grad))

    def detach(self):
        """Removes the post-accumulate hooks (call before the parameter set is re-laid out, e.g. after densify / prune:
        gaussian_model.GaussianModel._rebind builds a new reducer over the new tensors)."""
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def layout(self):
        return list(self.offsets), self.n

    def zero(self):
        self.flat.zero_()
        self.reset()

    def reset(self):
        """Per-step bookkeeping only (the fused optimizer clears the buffer itself)."""
        self._arrived = 0
        self._work = []

    def direct_written(self, params):
        """An op accumulated the (final) gradients of `params` straight into their .grad views -- no AccumulateGrad node
        ran for them, so no post-accumulate hook fires: the same bookkeeping, called by the op's backward."""
        early = {id(q) for q in self.early}
        for p in params:
            if id(p) in early:
                self._on_early_grad(p)
            else:
                self.dirty = True

    def _on_early_grad(self, _p):
        self.dirty = True
        self._arrived += 1
        if self._arrived == len(self.early) and self._distributed():
            self._launch(0, self.n_early)

    def _distributed(self):
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _launch(self, lo, hi):
        import torch.distributed as dist
        if hi > lo:
            self._work.append(dist.all_reduce(self.flat_all[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce(self):
        """Call after backward.  Launches what is still outstanding and makes the current stream wait for all of it."""
        if not self._distributed():
            return self.n
        early_done = self._arrived >= len(self.early) and len(self.early) > 0
        if not early_done:
            self._launch(0, self.n_early)
        self._launch(self.n_early, self.n + TAIL)
        for w in self._work:
            w.wait()
        self._work = []
        self._arrived = 0
        return self.n
