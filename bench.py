#!/usr/bin/env python
"""bench.py -- train frames/s of the DIMO deform -> raster -> loss step on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host CPU (oracle port)
  python bench.py --workload c2|c3|c4|c5                   # BASELINE.json configs[1..4]; c3 is the headline (default)

One "step" = one optimisation step over S frames per GPU: find_knn -> TimeNet over the unique (motion,t)
pairs -> LBS -> batched rasterisation -> {MSE, SSIM, mask-MSE} -> backward -> [all-reduce] -> Adam
(c4: forward only -- 4-D inference, S rendered frames per step).
One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train frames/s (deform+raster+SSIM fwd+bwd)"
METRIC_INFER = "render frames/s (deform+raster fwd, 4-D inference)"

WORKLOADS = {
    # bm x bv x bf = motions x views x frames per step and GPU (main_train_dimo.py:266-270 with batch_size 2)
    "c3": dict(mode="train", N=100_000, M=512, H=512, W=512, bm=4, bv=2, bf=2, motions_per_gpu=16, frames=32, views=8,
               desc="c3 shard: 100k synthetic Gaussians, 512 control points, 512x512, 16 motions x 32 frames per GPU"),
    "c2": dict(mode="train", N=30_000, M=512, H=512, W=512, bm=4, bv=2, bf=2, motions_per_gpu=51, frames=20, views=9,
               desc="c2 shape: 30k synthetic Gaussians, 512 control points, 512x512, 51 motions x 20 frames"),
    "c4": dict(mode="inference", N=30_000, M=512, H=1024, W=1024, bm=1, bv=4, bf=4, motions_per_gpu=1, frames=32,
               views=120,
               desc="c4: 4-D inference, 30k synthetic Gaussians, orbit 120 views x 32 frames at 1024x1024, (view, frame) "
                    "pairs sharded over the GPUs, forward only"),
    "c5": dict(mode="train", N=500_000, M=512, H=800, W=800, bm=4, bv=2, bf=2, motions_per_gpu=32, frames=32, views=8,
               desc="c5 shard: 500k-Gaussian stress, 32 of 256 motions per GPU, 800x800, KNN + SSIM on"),
    "small": dict(mode="train", N=5_000, M=128, H=128, W=128, bm=2, bv=2, bf=2, motions_per_gpu=4, frames=8, views=4,
                  desc="small smoke workload"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--min-seconds", type=float, default=2.0,
                    help="the timed region repeats the K-step block until it is at least this long (sustained clocks); "
                         "the first block alone is reported as `burst`")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostics: do not poll nvidia-smi during the timed region")
    ap.add_argument("--step-events", action="store_true",
                    help="diagnostics: one CUDA event per step in the timed region, per-step statistics on stderr")
    ap.add_argument("--e2e-gt", default="u8", choices=["u8", "f32"],
                    help="ground truth crossing PCIe every step: u8 = the 8-bit samples it was decoded from, converted on "
                         "the device by dimo_gt_fetch (GroundTruthCache semantics); f32 = the reference's host floats")
    ap.add_argument("--e2e-read", default="lagged", choices=["item", "lagged", "async"],
                    help="how the host reads every step's loss.  item: loss.item() (the host blocks on the step it just "
                         "enqueued); lagged: non-blocking D2H into pinned memory + the host waits for and reads the "
                         "PREVIOUS step's value while the current step runs (what a training loop that logs the loss "
                         "does); async: non-blocking D2H, one sync at the end")
    ap.add_argument("--cpu-frames", type=int, default=1, help="frames in the bounded CPU sample")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the whole-step CUDA graph")
    ap.add_argument("--regularisers", action="store_true",
                    help="add the depth / normal smoothness terms of the real step (main_train_dimo.py:363-372); "
                         "not part of the BASELINE metric, reported for completeness")
    return ap.parse_args()


def workload_config(wl):
    """The `config` object of the JSON line: names the workload only, identical for both arms."""
    S = wl["bm"] * wl["bv"] * wl["bf"]
    return {"workload": wl["desc"], "mode": wl["mode"], "frames_per_step_per_gpu": S, "H": wl["H"], "W": wl["W"],
            "gaussians": wl["N"], "control_points": wl["M"],
            "loss": "MSE + SSIM + mask MSE" if wl["mode"] == "train" else "none (forward only)",
            "l2": "inputs larger than L2: the per-step working set (ground truth + splat / instance / gradient buffers, "
                  "> 250 MB at c3) exceeds the 126 MB L2, no explicit flush"}


# ------------------------------------------------------------------------------------------------
def build_model(wl, rank, device, seed=0):
    import torch
    from dimo_b200 import synthetic
    from dimo_b200.renderer import Renderer
    torch.manual_seed(seed)               # TimeNet's xavier init (replicated on every rank) draws from the global RNG
    sc = synthetic.make_scene(wl["N"], n_ctrl=wl["M"], n_motions=wl["motions_per_gpu"], seed=seed)
    # each rank owns its own block of motions (latent codes): different seed stream for the latents only
    g = torch.Generator().manual_seed(1000 + rank)
    sc["_latent_codes"] = torch.randn(wl["motions_per_gpu"], 32, generator=g)
    r = Renderer(sh_degree=0, white_background=True, num_latent_code=wl["motions_per_gpu"], add_normal=True,
                 device=device)
    r.gaussians.load_state(sc)
    # xavier everywhere, the two head layers scaled x0.01 instead of the zero/identity init
    tn = r.gaussians._timenet
    gi = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for lin in (tn.pts_layers[-1], tn.rot_layers[-1]):
            w = torch.empty_like(lin.weight, device="cpu")
            torch.nn.init.xavier_uniform_(w, generator=gi)
            lin.weight.copy_(0.01 * w)
        tn.rot_layers[-1].bias.copy_(torch.tensor([1.0, 0, 0, 0]))
    return r, sc


def step_schedule(wl, step, rank=0, world=1):
    """Deterministic stand-in for the reference's random.sample (main_train_dimo.py:266-270):
    bm motions x bv views x bf frames, motion-major.  Inference (c4): the (view, frame) grid is block-partitioned over
    the ranks and walked bv x bf pairs at a time."""
    if wl["mode"] == "inference":
        vblocks, fblocks = wl["views"] // wl["bv"], wl["frames"] // wl["bf"]
        per_rank = (vblocks * fblocks) // world
        blk = rank * per_rank + step % max(per_rank, 1)
        v0, f0 = (blk // fblocks) * wl["bv"], (blk % fblocks) * wl["bf"]
        return [(0, v0 + i, f0 + j) for i in range(wl["bv"]) for j in range(wl["bf"])]
    ms = [(step * wl["bm"] + i) % wl["motions_per_gpu"] for i in range(wl["bm"])]
    vs = [(step * wl["bv"] + i) % wl["views"] for i in range(wl["bv"])]
    fs = [(step * wl["bf"] + i) % wl["frames"] for i in range(wl["bf"])]
    return [(m, v, f) for m in ms for v in vs for f in fs]


class ClockSampler(threading.Thread):
    """SM clock, power draw and throttle reasons sampled DURING the timed region: NVML every 20 ms when the binding is
    importable (nvidia_ml_py), else one nvidia-smi query per ~0.1 s."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.power, self.reasons = [], [], set()
        self.max_mhz = None
        self.source = None
        self._halt = threading.Event()

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        # NVML enumerates physical GPUs: honour CUDA_VISIBLE_DEVICES when it is a plain index list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.index
        try:
            ids = [int(x) for x in vis.split(",") if x.strip() != ""]
            if ids:
                idx = ids[self.index]
        except ValueError:
            pass
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self.source = "nvml"
        while not self._halt.is_set():
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            try:
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for n, bit in bits.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.02)

    def _run_smi(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        self.source = "nvidia-smi"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0])); self.max_mhz = float(f[1])
                try:
                    self.power.append(float(f[2]))
                except ValueError:
                    pass
                for n, v in zip(self.NAMES, f[3:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.05)

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            if not self._halt.is_set():
                self._run_smi()

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s, pw = sorted(self.samples), sorted(self.power)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_mhz_min": s[0] if s else None, "sm_max_mhz": self.max_mhz,
                "power_w": pw[len(pw) // 2] if pw else None, "power_w_max": pw[-1] if pw else None,
                "reasons": sorted(self.reasons), "samples": len(s), "source": self.source}


# ------------------------------------------------------------------------------------------------
def cpu_baseline(wl, n_frames, threads=None):
    """Times the oracle (CPU restatement of the reference algorithm, autograd backward) on `n_frames` frames of the
    same workload: deform (TimeNet+LBS) -> raster -> MSE+SSIM+mask loss -> backward (inference workloads: forward
    only).  Returns frames/s."""
    import torch
    from dimo_b200 import synthetic
    from oracle import deform as od, raster as orast, camera as ocam, loss as oloss, knn as oknn
    # the oracle's per-tile blend is a stream of small tensor ops: beyond ~16 threads the intra-op pool only
    # adds contention (measured: 8 threads 40 s/frame, 128 threads 657 s/frame on the B200 host) -> cap at 16
    threads = threads or min(os.cpu_count(), 16)
    torch.set_num_threads(threads)
    train = wl["mode"] == "train"
    sc = synthetic.make_scene(wl["N"], n_ctrl=wl["M"], n_motions=wl["motions_per_gpu"], seed=0)
    params = od.timenet_init(32, seed=0, final_scale=0.01)
    leaves = {k: v.clone().requires_grad_(train) for k, v in sc.items()}
    H, W = wl["H"], wl["W"]
    g = torch.Generator().manual_seed(5)
    t0 = time.perf_counter()
    dist, idx = oknn.knn(sc["_c_xyz"], sc["_xyz"], 4)
    done = 0
    with torch.set_grad_enabled(train):
        for (m, v, f) in step_schedule(wl, 0)[:n_frames]:
            cam = ocam.orbit_cam(v, wl["views"], W, H)
            t = f / wl["frames"]
            dxyz, dquat = od.timenet_forward(params, leaves["_c_xyz"], t, leaves["_latent_codes"][m])
            means, rots = od.lbs_deform(leaves["_xyz"], leaves["_rotation"], leaves["_c_xyz"],
                                        torch.exp(leaves["_c_radius"]), dxyz, dquat, idx, dist)
            out = orast.rasterize(means, torch.exp(leaves["_scaling"]), rots, torch.sigmoid(leaves["_opacity"]),
                                  cam.world_view_transform, cam.full_proj_transform, cam.camera_center, cam.tanfovx,
                                  cam.tanfovy, W, H, torch.ones(3),
                                  shs=torch.cat([leaves["_features_dc"], leaves["_features_rest"]], 1), sh_degree=0)
            img = out["image"].clamp(0, 1)[None]
            if train:
                gt = torch.rand(1, 3, H, W, generator=g); mk = torch.rand(1, 1, H, W, generator=g)
                loss = 5000.0 * oloss.mse_loss(img, gt) + 500.0 * (1 - oloss.ssim(img, gt)) + \
                    500.0 * oloss.mse_loss(out["alpha"][None], mk)
                loss.backward()
            done += 1
    dt = time.perf_counter() - t0
    return done / dt, dt, threads


def cpu_sample_text(args, wl, dt=None):
    what = "deform+raster+loss fwd+bwd" if wl["mode"] == "train" else "deform+raster fwd"
    S = wl["bm"] * wl["bv"] * wl["bf"]
    return f"{args.cpu_frames} frame(s) of the {args.workload} workload (of {S} per step)" + \
        (f", {dt:.1f} s" if dt is not None else "") + f", {what}, oracle port (PyTorch CPU" + \
        (" + autograd)" if wl["mode"] == "train" else ")")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    budget_s = 150.0                      # keep the whole arm within a few minutes
    t_start = time.perf_counter()
    for _ in range(max(1, args.steps)):
        fps, dt, threads = cpu_baseline(wl, args.cpu_frames)
        vals.append((fps, dt))
        if time.perf_counter() - t_start + dt > budget_s:
            break
    fps = sorted(v[0] for v in vals)[len(vals) // 2]
    line = {"impl": "reference", "metric": METRIC if wl["mode"] == "train" else METRIC_INFER, "value": fps,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(wl),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": cpu_sample_text(args, wl) + f"; median of {len(vals)} sample(s); one 'step' of "
                                                                   "this arm = that sample"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ge.ensure_built()
    from dimo_b200 import _lib, trainstep
    from dimo_b200.camera import orbit_minicam

    train = wl["mode"] == "train"
    r, _ = build_model(wl, rank, dev)
    if train:
        ts = trainstep.TrainStep(r, lr=1e-5, world=world, graph=not args.no_graph, probe_steps=6, capacity_margin=1.12,
                                 regularisers=args.regularisers)
    else:
        ts = trainstep.RenderStep(r, graph=not args.no_graph, probe_steps=6, capacity_margin=1.25)
    H, W = wl["H"], wl["W"]
    S = wl["bm"] * wl["bv"] * wl["bf"]
    cams_all = [orbit_minicam(v, wl["views"], W, H, device=dev) for v in range(wl["views"])]

    # Ground truth of a pool of steps: 8-bit samples (R, G, B, mask) like the decoded PNGs of the reference
    # (utils/load_utils.py:56-83), in pinned host memory for e2e and, converted by dimo_gt_fetch, resident in HBM as the
    # floats byte / 255 for `value`.
    pool = 4
    slots = torch.arange(S, dtype=torch.int32, device=dev)
    gt_u8_host, gt_dev, mk_dev, gt_f32_host = [], [], [], []
    if train:
        g = torch.Generator().manual_seed(1234 + rank)
        for _ in range(pool):
            u8 = torch.randint(0, 256, (S, 4, H, W), dtype=torch.uint8, generator=g).pin_memory()
            gt_u8_host.append(u8)
            rgb = torch.empty(S, 3, H, W, device=dev); msk = torch.empty(S, 1, H, W, device=dev)
            _lib.call("dimo_gt_fetch", S, H, W, H, W, 1, _lib.ptr(u8.to(dev)), _lib.ptr(slots), _lib.ptr(rgb),
                      _lib.ptr(msk), _lib.stream())
            gt_dev.append(rgb); mk_dev.append(msk)
            if args.e2e_gt == "f32":
                gt_f32_host.append((rgb.cpu().pin_memory(), msk.cpu().pin_memory()))
        torch.cuda.synchronize()

    # e2e: the step's ground truth travels host(pinned) -> device EVERY step on a copy stream, double-buffered, issued
    # after the previous step's launches so it overlaps that step's GPU work (the reference uploads per render on the
    # default stream, main_train_dimo.py:283-284); u8: converted on the device by one dimo_gt_fetch launch.
    copy_stream = torch.cuda.Stream(device=dev)
    if train:
        stage_u8 = [torch.empty(S, 4, H, W, dtype=torch.uint8, device=dev) for _ in range(2)]
        gt_buf = [torch.empty(S, 3, H, W, device=dev) for _ in range(2)]
        mk_buf = [torch.empty(S, 1, H, W, device=dev) for _ in range(2)]
    else:
        img_u8_host = [torch.empty(S, 3, H, W, dtype=torch.uint8).pin_memory() for _ in range(2)]
    ready_ev, free_ev, staged = [None, None], [None, None], {}
    loss_host = torch.zeros(64).pin_memory()
    loss_events = {}
    d2h_events = [None, None]

    def stage(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            if free_ev[slot] is not None:
                copy_stream.wait_event(free_ev[slot])
            if args.e2e_gt == "u8":
                stage_u8[slot].copy_(gt_u8_host[i % pool], non_blocking=True)
            else:
                gt_buf[slot].copy_(gt_f32_host[i % pool][0], non_blocking=True)
                mk_buf[slot].copy_(gt_f32_host[i % pool][1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        ready_ev[slot] = ev
        staged[i] = slot

    def one_step(i, e2e):
        frames = step_schedule(wl, i, rank, world)
        cams = [cams_all[v] for (_, v, _) in frames]
        times = [f / wl["frames"] for (_, _, f) in frames]
        lat = [m for (m, _, _) in frames]
        if not train:
            img = ts.run(cams, times, lat)
            if e2e:
                # what main_test_dimo.py:243-260 does per frame (.detach().cpu() -> uint8), batched: quantise on the
                # device, one D2H copy of the step's S frames into pinned memory, host waits one step behind
                slot = i % 2
                if d2h_events[slot] is not None:
                    d2h_events[slot].synchronize()
                u8 = (img * 255.0).to(torch.uint8)
                img_u8_host[slot].copy_(u8, non_blocking=True)
                ev = torch.cuda.Event(); ev.record()
                d2h_events[slot] = ev
            return None
        if e2e:
            if i not in staged:
                stage(i)
            slot = staged.pop(i)
            torch.cuda.current_stream().wait_event(ready_ev[slot])
            if args.e2e_gt == "u8":
                _lib.call("dimo_gt_fetch", S, H, W, H, W, 1, _lib.ptr(stage_u8[slot]), _lib.ptr(slots),
                          _lib.ptr(gt_buf[slot]), _lib.ptr(mk_buf[slot]), _lib.stream())
            loss = ts.run(cams, times, lat, gt_buf[slot], mk_buf[slot], wl["bm"])
            ev = torch.cuda.Event()
            ev.record()
            free_ev[slot] = ev
            # issued AFTER this step's work is enqueued: the copy engine then runs it under the step's
            # remaining GPU work instead of in front of the step's own small uploads
            stage(i + 1)
            if args.e2e_read == "item":
                return loss.item()      # device -> host read of the step's result, host waits for it
            slot_l = i % loss_host.numel()
            loss_host[slot_l].copy_(loss.detach(), non_blocking=True)
            if args.e2e_read == "lagged":
                ev = torch.cuda.Event()
                ev.record()
                prev = loss_events.pop(i - 1, None)
                loss_events[i] = (ev, slot_l)
                if prev is not None:    # the previous step's loss has landed (or we wait for it) -> host value
                    prev[0].synchronize()
                    return float(loss_host[prev[1]])
            return None
        return ts.run(cams, times, lat, gt_dev[i % pool], mk_dev[i % pool], wl["bm"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    counter = [0]
    host_ms = [0.0]

    def timed(e2e, steps, warmup, profile=False):
        """`warmup` untimed steps, then EXACTLY `steps` timed steps between barrier + synchronize, CUDA events on the
        launching stream, max over ranks.  Returns ms."""
        for _ in range(warmup):
            one_step(counter[0], e2e); counter[0] += 1
        barrier()
        if profile:
            _lib.PROFILE.reset(enabled=True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        marks = []
        for _ in range(steps):
            one_step(counter[0], e2e); counter[0] += 1
            if args.step_events:
                ev = torch.cuda.Event(enable_timing=True); ev.record(); marks.append((ev, time.perf_counter()))
        host_ms[0] = (time.perf_counter() - t_host) * 1000.0 / max(steps, 1)     # host time to ISSUE one step
        if e2e and loss_events:          # lagged reads: the last step's loss is read before the clock stops
            for ev, slot_l in loss_events.values():
                ev.synchronize()
                float(loss_host[slot_l])
            loss_events.clear()
        if e2e and not train:
            for ev in d2h_events:
                if ev is not None:
                    ev.synchronize()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if args.step_events and len(marks) > 2:
            series = [marks[k][0].elapsed_time(marks[k + 1][0]) for k in range(len(marks) - 1)]
            n10 = max(1, len(series) // 10)
            print("[step-events] mean gpu ms/step per tenth of the region: " +
                  " ".join(f"{sum(series[k:k + n10]) / len(series[k:k + n10]):.3f}" for k in range(0, len(series), n10)),
                  file=sys.stderr)
            d = sorted(series)
            h = sorted((marks[k + 1][1] - marks[k][1]) * 1000.0 for k in range(len(marks) - 1))
            q = lambda a, f: a[min(len(a) - 1, int(f * len(a)))]
            print(f"[step-events] n={len(d)} gpu ms/step p10 {q(d, .1):.3f} p50 {q(d, .5):.3f} p90 {q(d, .9):.3f} max {d[-1]:.3f} | "
                  f"host ms/step p10 {q(h, .1):.3f} p50 {q(h, .5):.3f} p90 {q(h, .9):.3f} max {h[-1]:.3f}", file=sys.stderr)
        if profile:
            _lib.PROFILE.enabled = False
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def work_stats():
        """How much list the blend kernels walk in the CURRENT state of the model (the optimizer changes the scene while
        the clock runs, so a step late in the region is not the same work as the first one): one untimed render."""
        frames = step_schedule(wl, 0, rank, world)
        with torch.no_grad():
            if getattr(r.gaussians, "neighbor_indices", None) is None:
                r.gaussians.find_knn(4)        # (never replace a table a captured graph may be reading)
            out = r.render_batch([cams_all[v] for (_, v, _) in frames], [f / wl["frames"] for (_, _, f) in frames],
                                 [m for (m, _, _) in frames], stage="s2", with_visibility=False, depth_normal=False,
                                 with_cpts=False)
        st = out["raster_state"]
        rg = st.ranges.view(-1, 2).long()
        length = float((rg[:, 1] - rg[:, 0]).sum())
        gx, gy = (W + 15) // 16, (H + 15) // 16
        pad = torch.zeros(S, gy * 16, gx * 16, device=dev)
        pad[:, :H, :W] = st.n_contrib.view(S, H, W).float()
        walked = float(pad.view(S, gy, 16, gx, 16).amax(dim=(2, 4)).sum())
        return {"instances": int(length), "walked_instances": int(walked),
                "walked_fraction": walked / max(length, 1.0)}

    sampler = ClockSampler(local) if (rank == 0 and not args.no_clocks) else None
    work_first = work_stats()
    # graph mode needs probe steps (eager, to learn the instance capacity) + the capture itself before the timed region
    extra_warm = (ts.probe_steps + 1) if ts.use_graph else 0
    ms_burst = timed(False, args.steps, args.warmup + extra_warm)          # the first K-step block on a cool GPU
    # sustained: repeat the K-step block back to back until the timed region lasts >= --min-seconds (one event pair
    # around ALL of it); `value` is taken from this region, the single block above is reported as `burst`
    blocks = max(1, int(math.ceil(1.05 * args.min_seconds * 1000.0 / max(ms_burst, 1e-3))))   # 5 % margin: a warm GPU can be faster than the burst block
    if sampler:
        sampler.start()
    ms = timed(False, args.steps * blocks, 0)
    host_issue_ms = host_ms[0]
    clocks = sampler.stop() if sampler else None
    work_last = work_stats()
    timed_steps = args.steps * blocks
    value = world * S * timed_steps / (ms / 1000.0)

    e2e = None
    if not args.no_e2e:
        staged.clear()
        ms_e = timed(True, timed_steps, 2)
        cam_bytes = S * 40 * 4
        if train:
            gt_bytes = S * 4 * H * W * (1 if args.e2e_gt == "u8" else 4)
            e2e = {"value": world * S * timed_steps / (ms_e / 1000.0), "unit": "frames/s",
                   "h2d_bytes_per_step": gt_bytes + cam_bytes, "d2h_bytes_per_step": 4,
                   "mode": f"ground truth ({args.e2e_gt}) from pinned host memory every step on a copy stream "
                           f"(double-buffered), loss read {args.e2e_read}", "timed_steps": timed_steps}
        else:
            e2e = {"value": world * S * timed_steps / (ms_e / 1000.0), "unit": "frames/s",
                   "h2d_bytes_per_step": cam_bytes, "d2h_bytes_per_step": S * 3 * H * W,
                   "mode": "cameras up, the step's S rendered frames down as uint8 (main_test_dimo.py:243-260), host one "
                           "step behind", "timed_steps": timed_steps}
    count_seen, capacity, overflow = ts.overflowed()
    graph_used = ts.use_graph and ts.graph is not None and getattr(ts, "graph_error", None) is None

    # per-kernel durations: the same step launched eagerly with CUDA events around every C-ABI call
    # (a captured graph cannot be timed per kernel); identical kernels, identical inputs
    ts.use_graph = False
    prof_steps = min(args.steps, 5)
    ms_prof = timed(False, prof_steps, 1, profile=True)
    prof = _lib.PROFILE.summary()
    for rec in prof.values():
        rec["ms_per_step"] = rec["ms"] / prof_steps

    def finish():
        """Multi-rank teardown.  A captured CUDA graph keeps references into the NCCL communicator and
        destroy_process_group() can block on it, so: drain the device, meet at a barrier, flush, and leave."""
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return

    # roofline for the dominant kernel group of the step (measured live above with CUDA events per C-ABI call)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    roof = None
    ms_k = lambda name: prof.get(name, {}).get("ms_per_step", 0.0)
    if prof:
        top = max(((k, v) for k, v in prof.items() if not k.startswith("py:")), key=lambda kv: kv[1]["ms"])
        name, rec = top
        st = dict(_lib.PROFILE.extra)
        st["walked"] = work_last["walked_instances"]       # the profiled pass runs on the model as the timed region left it
        st["R"] = work_last["instances"]
        alg = algorithmic_bytes(name, wl, S, st)
        dur_s = rec["ms"] / rec["calls"] / 1000.0
        km = kernel_metrics(args.workload, name)
        roof = {"kernel": name, "bound": "hbm", "achieved": alg / dur_s / 1e9 if alg else None, "peak": hbm_peak,
                "unit": "GB/s", "frac": (alg / dur_s / 1e9 / hbm_peak) if alg else None,
                "traffic": km.get("dram_bytes"), "traffic_source": km.get("source"),
                "peak_source": peak_src, "avg_launch_ms": rec["ms"] / rec["calls"],
                "share_of_step": rec["ms_per_step"] / (ms_prof / prof_steps),
                "timed_in": f"eager profiled pass of {prof_steps} steps ({ms_prof / prof_steps:.3f} ms/step) right after the timed region",
                # what actually binds the blend kernels: FP32 FMA + ALU + MUFU issue slots; from the
                # committed `ncu --set full` capture of the shipped kernel
                "issue": ({"bound": "fp32_issue", "pipe_fma_pct": km.get("pipe_fma"), "pipe_alu_pct": km.get("pipe_alu"),
                           "pipe_xu_pct": km.get("pipe_xu"), "issue_active_pct": km.get("issue_active"),
                           "source": km.get("source")} if km else None),
                "breakdown_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                "instances_R_per_step": st.get("R"), "walked_instances_per_step": st.get("walked"),
                "pairs_note": "R = tile instances built per step, walked = instances the blend kernels read before every "
                              "pixel of the tile has its last contributor; both of the model state after the timed region"}

    cpu = None
    if not args.no_cpu_baseline:
        fps, dt, threads = cpu_baseline(wl, args.cpu_frames)
        cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": cpu_sample_text(args, wl, dt)}

    t_fwd = ms_k("dimo_raster_preprocess") + ms_k("dimo_raster_bin") + ms_k("dimo_raster_blend_fwd")
    t_bwd = ms_k("dimo_raster_blend_bwd") + ms_k("dimo_raster_preprocess_bwd")
    mpix = S * H * W / 1e6
    line = {"metric": METRIC if train else METRIC_INFER, "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / timed_steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(wl),
            "timing": {"timed_steps": timed_steps, "timed_region_s": ms / 1000.0, "blocks_of_steps": blocks,
                       "host_issue_ms_per_step": host_issue_ms,
                       "work_drift": {"before_warmup": work_first, "after_timed_region": work_last,
                                      "note": "tile-list entries the blend kernels walk per step (largest last-contributor "
                                              "index per tile, summed): the optimizer keeps changing the scene while the "
                                              "clock runs, so `burst` (first block) and `value` (whole region) time "
                                              "different amounts of rasteriser work"},
                       "burst": {"steps": args.steps, "ms_per_step": ms_burst / args.steps,
                                 "value": world * S * args.steps / (ms_burst / 1000.0)}},
            "impl_detail": {
                "parallelism": f"motion-sharded dp{world}, one flat NCCL all-reduce/step" if (world > 1 and train) else
                               (f"(view, frame) pairs sharded over {world} GPUs, no collective" if world > 1 else "single GPU"),
                "execution": ("whole step replayed as one CUDA graph (rasteriser in capacity mode: "
                              f"{capacity} instance slots, max count seen {count_seen}, overflow={overflow})")
                if graph_used else ("eager launches" + (f" (graph capture failed: {ts.graph_error})" if getattr(ts, "graph_error", None) else "")),
                "optimizer": "dimo_adam_step: one launch over the flat parameter/gradient buffers, zero_grad folded in, "
                             "gated on the all-reduced overflow word" if train else None,
                "regularisers": bool(args.regularisers)},
            "raster": {"MPix_per_s_step": value * H * W / 1e6,
                       "MPix_per_s_fwd_kernels": mpix / (t_fwd / 1000.0) if t_fwd > 0 else None,
                       "MPix_per_s_fwd_bwd_kernels": mpix / ((t_fwd + t_bwd) / 1000.0) if (t_fwd > 0 and t_bwd > 0) else None,
                       "note": "step: whole-step throughput x pixels per frame; *_kernels: pixels of one step / summed "
                               "durations of the rasteriser's own calls (preprocess + bin + blend_fwd [+ blend_bwd + "
                               "preprocess_bwd]) in the eager profiled pass"},
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": _lib.PROFILE.kernel_launches_per_step(prof_steps)}
    if overflow:
        line["invalid"] = "instance capacity overflow during the timed region"
    print(json.dumps(line), flush=True)
    finish()


def kernel_metrics(workload, name):
    """dram bytes per launch and pipe utilisations of `name` from the committed ncu capture of the SHIPPED kernels
    (profiles/r2_kernel_metrics.json, written by tools/ncu_metrics_json.py from an `ncu --set full` report)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_kernel_metrics.json"))).get(workload, {}).get(name, {})
    except Exception:
        return {}


def algorithmic_bytes(name, wl, S, st):
    """Algorithmic HBM bytes per launch (DESIGN.md 'Kernels').  R = tile instances built by the binning
    stage, Wk = the instances the blend kernels actually walk (a tile stops at its deepest last contributor), both of
    the model's state when the per-kernel pass ran; P = pixels, BN = frames x Gaussians."""
    P = wl["H"] * wl["W"] * S
    R = st.get("R") or 0
    Wk = st.get("walked") or R
    BN = wl["N"] * S
    table = {
        # per walked instance: 4 B list word + 64 B blend record; per pixel: colour 12 + alpha 4 + final_T 4 + n_contrib 4
        "dimo_raster_blend_fwd": 68 * Wk + 24 * P,
        # per walked instance: the same 68 B in + 36 B of gradient fields reduced into the record table; per pixel 24 B in
        "dimo_raster_blend_bwd": 68 * Wk + 36 * Wk + 24 * P,
        # in: means3D 12 + scales 12 + rotation 16 + opacity 4 + SH 12; out: record 64 + radius 4 + count 4 + rect 8 + key 4
        "dimo_raster_preprocess": 56 * BN + 84 * BN,
        # depth sort: 4 passes x (key + index, read + write) = 64 B per splat; histogram 8 B; scatter 12 B per splat
        # + one 4 B list word per instance
        "dimo_raster_bin": (64 + 8 + 12) * BN + 4 * R,
        # in: gradient record 64 + the forward's inputs 56 + radius 4; out: d(means3D) 12 + d(rotation) 16 per (frame,
        # Gaussian); the shared parameters' gradients leave as [N, *] sums
        "dimo_raster_preprocess_bwd": 124 * BN + 28 * BN,
        "dimo_ssim_fwd": 8 * 3 * P + 36 * P,
        "dimo_ssim_bwd": 60 * P + 12 * P,
        "dimo_adam_step": 32 * st.get("n_params", 0),
    }
    return table.get(name)


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
