#!/usr/bin/env python
"""Microbenchmarks of the rows widened from SURVEY.md 8f (CUDA events on the launching stream, warm-up, synchronise on
both sides; inputs larger than L2 or an explicit L2 flush where the kernel is HBM-bound).  One JSON line per kernel:
    python tools/widen_bench.py > gpurun_out/widen_bench.jsonl
"""
import json
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PEAK = 6544.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)                    # > L2: evicts the previous iteration's lines
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def section(name):
    def deco(f):
        try:
            f()
        except Exception as e:
            emit(kernel=name, error=f"{type(e).__name__}: {e}", trace=traceback.format_exc()[-400:])
    return deco


def main():
    from dimo_b200 import points, regularisers as reg
    from dimo_b200.data import GroundTruthCache
    dev = "cuda"
    flush = torch.zeros(64 * 1024 * 1024, device=dev)          # 256 MB > 126 MB L2
    g = torch.Generator().manual_seed(0)

    @section("dimo_gt_fetch")
    def _():
        for res, dtype in ((512, torch.uint8), (256, torch.uint8), (512, torch.float32)):
            cache = GroundTruthCache(4, 8, 16, 512, dtype=dtype)        # 512 frames: 537 MB (u8) / 2.1 GB (f32)
            if dtype == torch.uint8:
                cache.store.copy_(torch.randint(0, 256, cache.store.shape, dtype=torch.uint8, device=dev))
            else:
                cache.store.uniform_()
            rng = np.random.default_rng(0)
            triples = [(int(rng.integers(4)), int(rng.integers(8)), int(rng.integers(16))) for _ in range(16)]
            out = cache.fetch(triples, res)
            ms = timed(lambda: cache.fetch(triples, res, out=out), flush=flush)
            scale = 512 / res
            taps = 1 if res == 512 else 4
            rd = 16 * 4 * res * res * min(taps, scale * scale) * cache.store.element_size()   # unique source bytes
            wr = 16 * 4 * res * res * 4
            emit(kernel="dimo_gt_fetch", store=str(dtype), frames=16, src=512, out=res, ms=ms,
                 algorithmic_bytes=rd + wr, achieved_gbs=(rd + wr) / ms / 1e6, peak_gbs=PEAK,
                 frac=(rd + wr) / ms / 1e6 / PEAK, note="includes the 64-byte slot-list upload issued by fetch()")

    @section("arap")
    def _():
        M, T = 512, 8
        base = (torch.rand(M, 3, generator=g) - 0.5) * 0.45
        nodes = torch.stack([base + 0.004 * t * torch.randn(M, 3, generator=g) for t in range(T)]).cuda()

        def fused():
            x = nodes.clone().requires_grad_(True)
            e, _ = reg.arap_loss_points(x, fused=True)
            e.backward()

        def unfused():
            x = nodes.clone().requires_grad_(True)
            e, _ = reg.arap_loss_points(x, fused=False)
            e.backward()

        emit(kernel="arap fwd+bwd (M=512, T=8)", fused_ms=timed(fused), torch_formulation_ms=timed(unfused),
             launches_fused="2 kernels + clone / scale glue", note="reference: ~40 launches + a Python loop over frames")

    @section("fps")
    def _():
        for N, K in ((512, 512), (30000, 512), (100000, 512), (500000, 512)):
            pts = torch.rand(1, N, 3, generator=g).cuda()
            ms = timed(lambda: points.sample_farthest_points(pts, K), iters=5, warmup=1)
            emit(kernel="dimo_fps", N=N, K=K, ms=ms, us_per_round=1e3 * ms / K)

    @section("chamfer / ball query")
    def _():
        a = (torch.randn(512, 3, generator=g) * 0.2).cuda().requires_grad_(True)
        b = (torch.randn(512, 3, generator=g) * 0.2).cuda()

        def ch():
            v = points.chamfer_forward(a[None], b[None])
            v.backward()

        emit(kernel="chamfer fwd+bwd (512 x 512)", ms=timed(ch))
        p = (torch.rand(8, 512, 3, generator=g) - 0.5).cuda() * 0.45
        emit(kernel="dimo_ball_query (8 x 512, K=11)", ms=timed(lambda: points.ball_query(p, p, K=11, radius=0.1,
                                                                                         return_nn=False)))

    @section("densify re-layout")
    def _():
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import model_scenario as ms_
        from dimo_b200 import synthetic
        from dimo_b200.renderer import Renderer
        r = Renderer(sh_degree=0, device="cuda", num_latent_code=16)
        gm = r.gaussians
        gm.load_state(synthetic.make_scene(100000, n_ctrl=512, n_motions=16, seed=0))
        gm.spatial_lr_scale = 1
        gm.training_setup(ms_.train_args(), optimizer="fused")
        n = gm._xyz.shape[0]
        gm.xyz_gradient_accum = torch.rand(n, 1, device=dev) * 0.04
        gm.denom = torch.ones(n, 1, device=dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gm.densify_and_prune(0.02, min_opacity=0.01, extent=4, max_screen_size=None)
        e1.record()
        torch.cuda.synchronize()
        emit(kernel="densify_and_prune (100k Gaussians, fused optimizer, one flat re-layout)", ms=e0.elapsed_time(e1),
             n_before=n, n_after=int(gm._xyz.shape[0]))


if __name__ == "__main__":
    main()
