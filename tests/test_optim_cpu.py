"""CPU: the Adam oracle is pinned bit-exact against torch.optim.Adam (the class the reference instantiates,
renderer/latent_gs_renderer.py:475) with per-group learning rates that change between steps."""
import torch

from oracle import optim as oopt


def test_adam_oracle_equals_torch_adam():
    g = torch.Generator().manual_seed(0)
    shapes = [(50, 3), (50, 1), (7,), (4, 32), (16, 24)]
    lrs = [1.6e-4, 5e-2, 1e-3, 2.5e-3, 8e-4]
    ref = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    mine = [p.detach().clone() for p in ref]
    m = [torch.zeros_like(p) for p in mine]
    v = [torch.zeros_like(p) for p in mine]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ref, lrs)], lr=0.0, eps=1e-15)
    for step in range(1, 8):
        grads = [torch.randn(s, generator=g) * (10.0 ** (step % 3 - 1)) for s in shapes]
        if step == 4:                                   # a parameter whose gradient is exactly zero this step
            grads[3].zero_()
        if step == 5:                                   # update_learning_rate (:502-520) between steps
            lrs[0] *= 0.5
            opt.param_groups[0]["lr"] = lrs[0]
        for p, gr in zip(ref, grads):
            p.grad = gr.clone()
        opt.step()
        oopt.adam_step(mine, grads, m, v, step, lrs)
        for a, b in zip(ref, mine):
            assert torch.equal(a.detach(), b), f"step {step}"
