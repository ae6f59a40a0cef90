#!/usr/bin/env python
"""Bring-up probe for the tcgen05 linear kernel (csrc/mlp_tc.cu): runs on the GPU box, prints for each knob setting
the error of random problems and where one-hot inputs land in the output (reveals descriptor/layout mistakes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dimo_b200 import _lib  # noqa: E402


def linear_tc(X, W, b=None, relu=False, mask=None, acc=None):
    R, K = X.shape
    No = W.shape[0]
    Y = torch.zeros(R, No, device="cuda") if acc is None else acc.clone()
    _lib.call("dimo_linear_tc", R, K, No, _lib.ptr(X), X.stride(0), _lib.ptr(mask), 0 if mask is None else mask.stride(0),
              _lib.ptr(W), _lib.ptr(b), _lib.ptr(Y), Y.stride(0), int(relu), int(acc is not None), _lib.stream())
    torch.cuda.synchronize()
    return Y


def main():
    torch.manual_seed(0)
    for swap in (0, 1):
        _lib.call("dimo_tc_debug_set", 0, swap)
        for single in (0, 1):
            _lib.call("dimo_tc_debug_set", 1, single)
            for (R, K, No) in [(128, 32, 128), (128, 64, 128), (4096, 256, 256), (300, 104, 256)]:
                X = torch.randn(R, K, device="cuda"); W = torch.randn(No, K, device="cuda") / K ** 0.5
                Y = linear_tc(X, W)
                ref = (X.double() @ W.double().t())
                err = ((Y.double() - ref).abs().max() / ref.abs().max()).item()
                print(f"swap={swap} single={single} R={R} K={K} No={No} rel_err={err:.3e} finite={bool(torch.isfinite(Y).all())}")
        # one-hot placement probe, K = 32
        _lib.call("dimo_tc_debug_set", 1, 1)
        for (r0, k0, n0) in [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (9, 5, 17), (100, 31, 127), (37, 12, 64)]:
            X = torch.zeros(128, 32, device="cuda"); W = torch.zeros(128, 32, device="cuda")
            X[r0, k0] = 1.0; W[n0, k0] = 2.0
            Y = linear_tc(X, W)
            nz = torch.nonzero(Y.abs() > 1e-3)
            print(f"swap={swap} onehot r0={r0} k0={k0} n0={n0} -> nonzeros={nz[:6].tolist()} vals={[round(Y[i,j].item(),3) for i,j in nz[:6].tolist()]}")
    _lib.call("dimo_tc_debug_set", 0, 0); _lib.call("dimo_tc_debug_set", 1, 0)
    # epilogue options with the default knobs
    X = torch.randn(1000, 256, device="cuda"); W = torch.randn(256, 256, device="cuda") / 16; b = torch.randn(256, device="cuda")
    Ym = torch.randn(1000, 256, device="cuda")
    Y = linear_tc(X, W, b, relu=True)
    ref = torch.relu(X.double() @ W.double().t() + b.double())
    print("bias+relu err", ((Y.double() - ref).abs().max() / ref.abs().max()).item())
    acc0 = torch.randn(1000, 256, device="cuda")
    Y = linear_tc(X, W, mask=Ym, acc=acc0)
    ref = acc0.double() + (X.double() * (Ym > 0)) @ W.double().t()
    print("mask+accumulate err", ((Y.double() - ref).abs().max() / ref.abs().max()).item())
    # exact call shapes of the TimeNet data-gradient (strided views into the [R,360] concat buffers)
    def call_tc(R, Kred, Nout, A, lda, mask, ldm, Wt, Y, ldy, acc):
        _lib.call("dimo_linear_tc", R, Kred, Nout, A, lda, mask, ldm, _lib.ptr(Wt), None, Y, ldy, 0, int(acc), _lib.stream())
        torch.cuda.synchronize()
    for R in (1536, 1000, 4096):
        E, CAT, Hd = 104, 360, 256
        dY = torch.randn(R, Hd, device="cuda"); Ym = torch.randn(R, Hd, device="cuda")
        W5 = torch.randn(Hd, CAT, device="cuda") / 16; W0 = torch.randn(Hd, E, device="cuda") / 16
        dcat = torch.full((R, CAT), 7.0, device="cuda")
        call_tc(R, Hd, CAT, dY.data_ptr(), Hd, Ym.data_ptr(), Hd, W5.t().contiguous(), dcat.data_ptr(), CAT, False)
        ref5 = (dY.double() * (Ym > 0)) @ W5.double()
        e5 = ((dcat.double() - ref5).abs().max() / ref5.abs().max()).item()
        dY0 = torch.randn(R, Hd, device="cuda"); Y0 = torch.randn(R, Hd, device="cuda")
        call_tc(R, Hd, E, dY0.data_ptr(), Hd, Y0.data_ptr(), Hd, W0.t().contiguous(), dcat.data_ptr(), CAT, True)
        ref = ref5.clone(); ref[:, :E] += (dY0.double() * (Y0 > 0)) @ W0.double()
        e0 = ((dcat.double() - ref).abs().max() / ref.abs().max()).item()
        # layer 4: dY and mask are column-offset views (ld 360) of the concat buffers
        catY = torch.randn(R, CAT, device="cuda"); W4 = torch.randn(Hd, Hd, device="cuda") / 16
        out = torch.zeros(R, Hd, device="cuda")
        call_tc(R, Hd, Hd, dcat.data_ptr() + 4 * E, CAT, catY.data_ptr() + 4 * E, CAT, W4.t().contiguous(), out.data_ptr(), Hd, False)
        ref4 = (dcat[:, E:].double() * (catY[:, E:] > 0)) @ W4.double()
        e4 = ((out.double() - ref4).abs().max() / ref4.abs().max()).item()
        print(f"dgrad shapes R={R}: layer5 {e5:.2e} layer0(acc) {e0:.2e} layer4(views) {e4:.2e}")
    # full TimeNet backward: tensor-core vs SIMT kernels on identical inputs
    import importlib
    from dimo_b200 import deform as dd
    def run_net(M, G, tc, fwd=True, dgrad=True, wgrad=True, sync=False):
        dd.USE_TC = tc; dd.TC_FWD = fwd; dd.TC_DGRAD = dgrad; dd.TC_WGRAD = wgrad
        _lib.SYNC_EVERY_CALL = sync
        torch.manual_seed(1)
        net = dd.TimeNet().cuda()
        with torch.no_grad():
            for q in net.parameters():
                q.copy_(torch.randn_like(q) * 0.05)
        pts = (torch.rand(M, 3, device="cuda") - 0.5).requires_grad_(True)
        lat = torch.randn(G, 32, device="cuda").requires_grad_(True)
        tms = torch.rand(G, device="cuda")
        dx, dq = net.forward_batched(pts, tms, lat)
        w1 = torch.randn_like(dx); w2 = torch.randn_like(dq)
        ((dx * w1).sum() + (dq * w2).sum()).backward()
        torch.cuda.synchronize()
        _lib.SYNC_EVERY_CALL = False
        names = ["dx", "dq", "dpts", "dlat"] + [n.replace("deformnet.", "d").replace("_layers.", "").replace("weight", "W").replace("bias", "b") for n, _ in net.named_parameters()]
        return names, [dx.detach(), dq.detach(), pts.grad, lat.grad] + [q.grad for q in net.parameters()]

    # activations / parameters after a tensor-core forward vs after a SIMT forward
    for (M, G) in [(512, 3)]:
        caps = {}
        for tc in (False, True):
            dd.USE_TC = tc; dd.TC_FWD = True
            dd.DEBUG_CAPTURE = []
            torch.manual_seed(1)
            net = dd.TimeNet().cuda()
            with torch.no_grad():
                for q in net.parameters():
                    q.copy_(torch.randn_like(q) * 0.05)
            p0 = [q.detach().clone() for q in net.parameters()]
            pts = (torch.rand(M, 3, device="cuda") - 0.5); lat = torch.randn(G, 32, device="cuda"); tms = torch.rand(G, device="cuda")
            with torch.no_grad():
                dx, dq = net.forward_batched(pts, tms, lat)
            torch.cuda.synchronize()
            pchg = max(float((a - b).abs().max()) for a, b in zip(p0, [q.detach() for q in net.parameters()]))
            caps[tc] = dd.DEBUG_CAPTURE[0]
            print(f"forward tc={tc}: params changed by {pchg:.1e}")
        dd.DEBUG_CAPTURE = None
        nm = ["cat", "hp", "hr"] + [f"act{i}" for i in (0, 1, 2, 3, 5, 6, 7)]
        for n, a, b in zip(nm, caps[True], caps[False]):
            d = (a - b).abs()
            flips = int(((a > 0) != (b > 0)).sum())
            print(f"  {n}: max abs diff {float(d.max()):.2e} (scale {float(b.abs().max()):.2e}) relu-mask flips {flips} of {a.numel()}"
                  f" rows with diff>1e-4: {int((d.max(dim=1).values > 1e-4).sum())}")
    for (M, G) in [(512, 3), (100, 1)]:
        names, ref = run_net(M, G, False)
        for label, kw in [("all", {}), ("all+sync", dict(sync=True)), ("fwd only", dict(dgrad=False, wgrad=False)),
                          ("dgrad only", dict(fwd=False, wgrad=False)), ("wgrad only", dict(fwd=False, dgrad=False))]:
            _, got = run_net(M, G, True, **kw)
            errs = {n: ((a.double() - b.double()).abs().max() / b.abs().max().clamp_min(1e-30)).item()
                    for n, a, b in zip(names, got, ref)}
            bad = " ".join(f"{k}:{v:.0e}" for k, v in errs.items() if v > 2e-5)
            print(f"TimeNet TC[{label}] vs SIMT M={M} G={G}: max {max(errs.values()):.1e} bad: {bad}")
    dd.TC_FWD = dd.TC_DGRAD = dd.TC_WGRAD = True
    dd.USE_TC = True
    # weight gradient (MN-major operands)
    for (R, K, No) in [(128, 128, 128), (4096, 256, 256), (4096, 360, 256), (4096, 104, 256), (1000, 256, 256)]:
        dY = torch.randn(R, No, device="cuda"); Yk = torch.randn(R, No, device="cuda"); Xk = torch.randn(R, K, device="cuda")
        dW = torch.zeros(No, K, device="cuda"); db = torch.zeros(No, device="cuda")
        _lib.call("dimo_linear_wgrad_tc", R, K, No, _lib.ptr(dY), No, _lib.ptr(Yk), No, _lib.ptr(Xk), K, _lib.ptr(dW),
                  _lib.ptr(db), _lib.stream())
        torch.cuda.synchronize()
        dYm = dY.double() * (Yk > 0)
        rW = dYm.t() @ Xk.double(); rb = dYm.sum(0)
        print(f"wgrad R={R} K={K} No={No} dW err {((dW.double()-rW).abs().max()/rW.abs().max()).item():.3e} "
              f"db err {((db.double()-rb).abs().max()/rb.abs().max()).item():.3e}")
    Wh = torch.randn(3, 256, device="cuda")
    Y = linear_tc(X, Wh)
    print("No=3 err", ((Y.double() - X.double() @ Wh.double().t()).abs().max()).item())


if __name__ == "__main__":
    main()
