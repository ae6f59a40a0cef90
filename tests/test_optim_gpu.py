"""-m gpu: the one-launch optimizer (dimo_adam_step), the grouped transpose, the mask-loss reduction and the step
loss with the reference's per-frame MSE weights, each against the CPU oracle / torch through the C ABI."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g) for s in shapes], g


def test_fused_adam_matches_oracle(cuda):
    """7 steps, 5 groups with their own learning rates, one rate changed mid-run, one all-zero gradient, tensor sizes
    that are not multiples of 4 (padding), gradients cleared by the same launch."""
    from dimo_b200.dist import FlatGradReducer
    from dimo_b200.optim import FusedAdam
    from oracle import optim as oopt
    shapes = [(50, 3), (50, 1), (7,), (4, 32), (16, 24), (1001,)]
    group_of = [0, 1, 2, 3, 4, 4]
    lrs = [1.6e-4, 5e-2, 1e-3, 2.5e-3, 8e-4]
    init, g = _make(shapes, 0)
    dev_params = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    red = FlatGradReducer(dev_params, early=dev_params[:2])
    groups = [{"params": [p for p, gi in zip(dev_params, group_of) if gi == k], "lr": lrs[k], "name": f"g{k}"}
              for k in range(5)]
    opt = FusedAdam(groups, red, eps=1e-15)
    assert all(p.data_ptr() >= opt.flat.data_ptr() for p in dev_params), "parameters were not re-homed"
    ref = [t.clone() for t in init]
    m = [torch.zeros_like(t) for t in ref]
    v = [torch.zeros_like(t) for t in ref]
    for step in range(1, 8):
        grads = [torch.randn(s, generator=g) * (10.0 ** (step % 3 - 1)) for s in shapes]
        if step == 4:
            grads[3].zero_()
        if step == 5:
            lrs[0] *= 0.5
            opt.param_groups[0]["lr"] = lrs[0]
        for p, gr in zip(dev_params, grads):
            p.grad.copy_(gr)                     # .grad is a view of the flat buffer
        before = [p.detach().clone() for p in dev_params]
        opt.step()
        opt.zero_grad()
        oopt.adam_step(ref, grads, m, v, step, [lrs[gi] for gi in group_of])
        assert float(red.flat.abs().max()) == 0.0, "gradients not cleared by the fused step"
        for i, (p, r, b0) in enumerate(zip(dev_params, ref, before)):
            upd = (p.detach() - b0).cpu()
            want = r - b0.cpu()
            scale = max(want.abs().max().item(), 1e-30)
            # both sides round the new parameter to fp32 (|p| up to ~4: half an ulp each), hence the absolute slack
            assert (upd - want).abs().max().item() <= 1e-5 * scale + 1e-6, f"step {step} tensor {i}"
            assert torch.allclose(p.detach().cpu(), r, rtol=1e-6, atol=1e-7), f"step {step} tensor {i}"
        for (off, numel), mm, vv in zip(red.offsets, m, v):
            assert torch.allclose(opt.exp_avg[off:off + numel].cpu(), mm.reshape(-1), rtol=1e-5, atol=1e-12)
            assert torch.allclose(opt.exp_avg_sq[off:off + numel].cpu(), vv.reshape(-1), rtol=1e-5, atol=1e-12)
    assert int(opt.state[0]) == 7 and int(opt.state[1]) == 0
    sd = opt.state_dict()
    assert sd["step"] == 7 and torch.allclose(sd["exp_avg"][:150].cpu(), m[0].reshape(-1), rtol=1e-5, atol=1e-8)


def test_fused_adam_large_and_graph_replay(cuda):
    """2.1 M parameters (the c3 shape) stepped eagerly and through CUDA-graph replays: the device-resident step
    counter and learning rates make replays equal to eager launches."""
    from dimo_b200.dist import FlatGradReducer
    from dimo_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(3)
    init = [torch.randn(100000, 3, generator=g), torch.randn(100000, 14 - 3, generator=g), torch.randn(647431, generator=g)]
    res = []
    for use_graph in (False, True):
        ps = [torch.nn.Parameter(t.clone().cuda()) for t in init]
        red = FlatGradReducer(ps)
        opt = FusedAdam([{"params": ps[:2], "lr": 1e-3, "name": "a"}, {"params": ps[2:], "lr": 1e-4, "name": "b"}], red)
        gsrc = torch.randn(red.flat.numel(), generator=torch.Generator().manual_seed(9)).cuda()
        graph = None
        if use_graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            opt.sync_lrs()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                red.flat.copy_(gsrc)
                opt.step()
        for step in range(4):
            if step == 2:
                opt.param_groups[1]["lr"] = 5e-5
            if graph is not None:
                opt.sync_lrs()
                graph.replay()
            else:
                red.flat.copy_(gsrc)
                opt.step()
        torch.cuda.synchronize()
        assert int(opt.state[0]) == 4
        res.append([p.detach().clone() for p in ps])
    for a, b in zip(*res):
        assert torch.equal(a, b), "graph replay differs from eager launches"


def test_transpose_grouped(cuda):
    from dimo_b200 import _lib
    shapes = [(256, 104), (256, 256), (256, 360), (4, 256), (33, 7)]
    src = [torch.randn(s, device="cuda") for s in shapes]
    dst = [torch.empty(s[1], s[0], device="cuda") for s in shapes]
    n = len(shapes)
    _lib.call("dimo_transpose_grouped", n, (ctypes.c_int * n)(*[s[0] for s in shapes]),
              (ctypes.c_int * n)(*[s[1] for s in shapes]), (ctypes.c_void_p * n)(*[t.data_ptr() for t in src]),
              (ctypes.c_void_p * n)(*[t.data_ptr() for t in dst]), _lib.stream())
    for a, b in zip(src, dst):
        assert torch.equal(a.t().contiguous(), b)


@pytest.mark.parametrize("n,offset", [(4 * 64 * 64, 0), (1001, 0), (5000, 1), (3, 0)])
def test_sqdiff_sum(cuda, n, offset):
    from dimo_b200 import _lib
    g = torch.Generator().manual_seed(n)
    a = torch.rand(n + offset, generator=g); b = torch.rand(n + offset, generator=g)
    ad, bd = a.cuda()[offset:], b.cuda()[offset:]            # offset 1: pointers not 16-byte aligned
    out = torch.empty(1, device="cuda"); acc = torch.full((1,), 2.0, device="cuda")
    _lib.call("dimo_sqdiff_sum", n, ad.data_ptr(), bd.data_ptr(), _lib.ptr(out), _lib.ptr(acc), 0.5, _lib.stream())
    want = ((a[offset:].double() - b[offset:].double()) ** 2).sum().item()
    assert abs(float(out) - want) <= 1e-5 * want
    assert abs(float(acc) - (2.0 + 0.5 * want)) <= 1e-5 * (2.0 + 0.5 * want)


def test_step_loss_frame_weights_and_upstream_gradient(cuda):
    """step loss with the reference's 1 / 0.5 MSE weighting (main_train_dimo.py:333-336) and a non-unit upstream
    gradient, vs the oracle's losses + autograd."""
    from dimo_b200 import trainstep
    from oracle import loss as ol
    S, H, W, nm = 4, 48, 40, 2
    g = torch.Generator().manual_seed(4)
    img = torch.rand(S, 3, H, W, generator=g) * 1.4 - 0.2          # values outside [0,1]: the clamp matters
    alp = torch.rand(S, 1, H, W, generator=g)
    gt = torch.rand(S, 3, H, W, generator=g); mk = torch.rand(S, 1, H, W, generator=g)
    fw = torch.tensor([1.0, 0.5, 1.0, 0.5])
    lw = trainstep.StepLossWeights
    oi = img.clone().requires_grad_(True); oa = alp.clone().requires_grad_(True)
    ic = oi.clamp(0, 1)
    lo = 0
    for f in range(S):
        lo = lo + lw.lambda_mse * fw[f] * ol.mse_loss(ic[f], gt[f])
    for m in range(nm):
        sl = slice(2 * m, 2 * m + 2)
        lo = lo + lw.lambda_ssim * (1 - ol.ssim(ic[sl], gt[sl])) + lw.lambda_mask * ol.mse_loss(oa[sl], mk[sl])
    (0.37 * lo).backward()
    ci = img.cuda().requires_grad_(True); ca = alp.cuda().requires_grad_(True)
    lc = trainstep.step_loss(ci, ca, gt.cuda(), mk.cuda(), nm, frame_w=fw.cuda())
    (0.37 * lc).backward()
    import gpu_parity as gp
    assert abs(float(lc) - float(lo)) <= 1e-5 * abs(float(lo)), (float(lc), float(lo))
    assert gp.rel_err(ci.grad, oi.grad) < 1e-4
    assert gp.rel_err(ca.grad, oa.grad) < 1e-5


def test_timenet_direct_grads_equal_autograd(cuda):
    """TimeNet.direct_grads (kernels accumulate into preallocated .grad tensors) gives the same gradients as the
    autograd-accumulated path."""
    from dimo_b200.deform import TimeNet
    torch.manual_seed(0)
    net = TimeNet(latent_code_dim=32).cuda()
    with torch.no_grad():
        for lin in (net.pts_layers[-1], net.rot_layers[-1]):
            lin.weight.copy_(0.05 * torch.randn_like(lin.weight))
    pts = (torch.rand(200, 3, device="cuda") - 0.5).requires_grad_(True)
    times = torch.rand(3, device="cuda")
    lat = torch.randn(3, 32, device="cuda", requires_grad=True)
    wx = torch.randn(3, 200, 3, device="cuda"); wq = torch.randn(3, 200, 4, device="cuda")

    def run(direct):
        for p in list(net.parameters()) + [pts, lat]:
            p.grad = None
        net.direct_grads = direct
        if direct:
            for p in net.parameters():
                p.grad = torch.zeros_like(p)
        dx, dq = net.forward_batched(pts, times, lat)
        ((dx * wx).sum() + (dq * wq).sum()).backward()
        return [p.grad.clone() for p in net.flat_params()] + [pts.grad.clone(), lat.grad.clone()]

    a, b = run(False), run(True)
    net.direct_grads = False
    import gpu_parity as gp
    for x, y in zip(a, b):
        assert gp.rel_err(y, x) < 1e-5


@pytest.mark.gpu
def test_fused_adam_restores_torch_format_checkpoint(cuda):
    """A torch.optim.Adam state_dict (what the reference's capture() stores, renderer/latent_gs_renderer.py:296-315) is
    restored into the flat moments, training continues as torch would, and torch_state_dict() round-trips it."""
    from dimo_b200.dist import FlatGradReducer
    from dimo_b200.optim import FusedAdam
    shapes = [(40, 3), (40, 1), (9,), (5, 16)]
    names, lrs = ["xyz", "opacity", "c_radius", "timenet"], [1.6e-4, 5e-2, 1e-3, 8e-4]
    init, g = _make(shapes, 5)
    grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(6)]
    # reference run: torch.optim.Adam for 6 steps, checkpoint after 3
    tp = [torch.nn.Parameter(t.clone().cuda()) for t in init]
    topt = torch.optim.Adam([{"params": [p], "lr": lr, "name": n} for p, lr, n in zip(tp, lrs, names)], lr=0.0, eps=1e-15)
    ckpt = None
    for k in range(6):
        for p, gr in zip(tp, grads[k]):
            p.grad = gr.cuda()
        topt.step()
        if k == 2:
            import copy
            ckpt = (copy.deepcopy(topt.state_dict()), [p.detach().clone() for p in tp])
    # fused optimizer restored from the torch checkpoint, three more steps
    fp = [torch.nn.Parameter(t.clone()) for t in ckpt[1]]
    red = FlatGradReducer(fp)
    fopt = FusedAdam([{"params": [p], "lr": 0.0, "name": n} for p, n in zip(fp, names)], red, eps=1e-15)
    fopt.load_state_dict(ckpt[0])
    assert [gg["lr"] for gg in fopt.param_groups] == lrs and int(fopt.state[0]) == 3
    for k in range(3, 6):
        for p, gr in zip(fp, grads[k]):
            p.grad.copy_(gr)
        fopt.step(); fopt.zero_grad()
    for a, b in zip(fp, tp):
        assert torch.allclose(a.detach(), b.detach(), rtol=2e-6, atol=1e-7)
    # and back: the torch layout of the fused state loads into a fresh torch optimizer
    sd = fopt.torch_state_dict()
    t2 = torch.optim.Adam([{"params": [p], "lr": lr, "name": n} for p, lr, n in zip(tp, lrs, names)], lr=0.0, eps=1e-15)
    t2.load_state_dict(sd)
    for i, p in enumerate(tp):
        assert torch.allclose(t2.state[p]["exp_avg"], topt.state[p]["exp_avg"], rtol=1e-5, atol=1e-9)
        assert int(t2.state[p]["step"]) == 6
