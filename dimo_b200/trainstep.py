"""One optimisation step of the DIMO inner loop on the B200 fast path.

Workload definition = the hot loop of GUI.train_step (main_train_dimo.py:253-417) restricted to the
north_star terms: find_knn -> S x [deform -> rasterise] -> {MSE per frame, SSIM per motion, mask MSE on
alpha} -> backward -> (one gradient all-reduce) -> Adam.  All S frames of the step go through ONE
launch set (Renderer.render_batch).  LPIPS and the ARAP/chamfer/smoothness regularisers are
SURVEY.md 8f "next" rows and are not part of this step.

Two execution modes, same kernels, same results:
  eager   every kernel is launched from Python each step; the rasteriser reads the instance count back once
          per step to size its buffers (the only host sync on the path);
  graph   the rasteriser runs in capacity mode (no read-back, overflow flag on the device) and the WHOLE step --
          forward, loss, backward, all-reduce, Adam -- is captured once in a CUDA graph and replayed; only the
          step's inputs (cameras, times, latent indices, ground truth) are copied into static buffers.
          A replay whose instance count exceeded the captured capacity (on ANY rank) can never reach the parameters:
          its overflow word is all-reduced with the gradients and makes dimo_adam_step discard the step; run() polls
          that word every `poll_every` replays (asynchronous 16-byte read), warns, grows the capacity and re-captures.
"""
import warnings


import torch

from . import _lib
from . import loss as _loss
from .dist import FlatGradReducer
from .optim import FusedAdam
from .renderer import Renderer


class StepLossWeights:
    # configs/train_config.yaml:34-38
    lambda_mse = 5000.0
    lambda_ssim = 500.0
    lambda_mask = 500.0
    lambda_smooth = 100.0        # edge-aware depth smoothness (add_depth, after depth_reg_start_iter), :49-51
    lambda_bilateral = 0.05      # bilateral normal smoothness (add_normal, after normal_reg_start_iter), :53-55


class _StepLoss(torch.autograd.Function):
    """loss = l_mse * sum_f w_f MSE(clamp(img_f), gt_f) + l_ssim * sum_m (1 - SSIM(clamp(img_m), gt_m))
            + l_mask * sum_m MSE(alpha_m, mask_m)
    over S frames grouped in n_motions equal groups (main_train_dimo.py:328-351).  w_f is the reference's 1 / 0.5
    weighting of reference / non-reference (view, frame) pairs (:333-336; `frame_w` [S] on the device, None = all 1).
    Launches: one fused SSIM+MSE forward over the images, one squared-difference reduction over the masks (both add
    their weighted sums straight into the loss scalar), two gradient kernels that read the upstream gradient as a
    device scalar.  No host sync, no elementwise glue."""

    @staticmethod
    def forward(ctx, image, alpha, gt, mask, frame_w, n_motions, l_mse, l_ssim, l_mask):
        S, _, H, W = image.shape
        dev = image.device
        image = image.contiguous(); alpha = alpha.contiguous(); gt = gt.contiguous(); mask = mask.contiguous()
        sums = torch.empty(4, dtype=torch.float32, device=dev)        # [ssim, l1, mse] of the images | mask sq. sum
        dm = torch.empty(3, S, 3, H, W, dtype=torch.float32, device=dev)
        n_img = float(3 * H * W)
        per_motion = S // n_motions
        # sum_f w_f MSE_f = sums[2] / n_img ; sum_m (1 - ssim_m) = n_motions - sums[0] / (per_motion * n_img)
        w_ssim, w_mse, w_mask = -l_ssim / (per_motion * n_img), l_mse / n_img, l_mask / (per_motion * H * W)
        loss = torch.full((), l_ssim * n_motions, dtype=torch.float32, device=dev)
        s = _lib.stream()
        _lib.call("dimo_ssim_fwd", S, 3, H, W, 1, _lib.ptr(image), _lib.ptr(gt), _lib.ptr(sums), _lib.ptr(dm),
                  _lib.ptr(frame_w), _lib.ptr(loss), w_ssim, 0.0, w_mse, s)
        _lib.call("dimo_sqdiff_sum", S * H * W, _lib.ptr(alpha), _lib.ptr(mask), sums.data_ptr() + 12, _lib.ptr(loss),
                  w_mask, s)
        ctx.save_for_backward(image, alpha, gt, mask, dm, frame_w)
        ctx.w = (w_ssim, w_mse, w_mask)
        return loss

    @staticmethod
    def backward(ctx, g):
        image, alpha, gt, mask, dm, frame_w = ctx.saved_tensors
        S, _, H, W = image.shape
        w_ssim, w_mse, w_mask = ctx.w
        s = _lib.stream()
        g = g.contiguous().float()
        d_img = torch.empty_like(image)
        d_alpha = torch.empty_like(alpha)
        _lib.call("dimo_ssim_bwd", S, 3, H, W, 1, _lib.ptr(image), _lib.ptr(gt), _lib.ptr(dm), w_ssim, 0.0, w_mse,
                  _lib.ptr(frame_w), _lib.ptr(g), _lib.ptr(d_img), s)
        _lib.call("dimo_ssim_bwd", S, 1, H, W, 0, _lib.ptr(alpha), _lib.ptr(mask), None, 0.0, 0.0, w_mask,
                  None, _lib.ptr(g), _lib.ptr(d_alpha), s)
        return d_img, d_alpha, None, None, None, None, None, None, None


def step_loss(image, alpha, gt, mask, n_motions, weights=StepLossWeights, frame_w=None):
    return _StepLoss.apply(image, alpha, gt, mask, frame_w, n_motions, weights.lambda_mse, weights.lambda_ssim,
                           weights.lambda_mask)


class TrainStep:
    """Holds the model + optimizer and runs steps.  `world` > 1: the flat gradient buffer is summed over the
    process group once per step (the only exchange on the path; frames are sharded by motion)."""

    def __init__(self, renderer: Renderer, lr=1e-4, world=1, stage="s2", graph=False, probe_steps=3,
                 capacity_margin=1.25, optimizer="fused", regularisers=False, poll_every=16):
        """lr: one float for every group, or {group name: lr} with the reference's group names
        (renderer/latent_gs_renderer.py:460-473).  optimizer: "fused" (dimo_adam_step, one launch incl. zero_grad) or
        "torch" (torch.optim.Adam(fused=True), kept for A/B runs)."""
        self.r = renderer
        self.g = renderer.gaussians
        self.stage = stage
        self.world = world
        g = self.g
        if optimizer == "fused" and g.optimizer is not None and g.reducer is not None:
            # GaussianModel.training_setup already built the flat buffers (reference flow: training_setup, then steps)
            self.reducer, self.opt = g.reducer, g.optimizer
            self.params = list(self.reducer.params)
        else:
            self.params = [p for p in g.parameters() if p.numel() > 0]
            # per-Gaussian parameters: their gradients are final once the LBS backward has run, i.e. before the
            # TimeNet backward -> bucket 0 of the flat buffer, all-reduced while the MLP backward executes
            early = [g._xyz, g._features_dc, g._features_rest, g._opacity, g._scaling, g._rotation]
            self.reducer = FlatGradReducer(self.params, early=early)
            groups = g.param_groups(lr)
            if optimizer == "fused":
                self.opt = FusedAdam(groups, self.reducer, eps=1e-15)
                g.reducer, g._optimizer_kind = self.reducer, "fused"
            else:
                self.opt = torch.optim.Adam([{"params": [p for p in gr["params"] if p.numel() > 0], "lr": gr["lr"],
                                              "name": gr["name"]} for gr in groups if any(p.numel() for p in gr["params"])],
                                            lr=0.0, eps=1e-15, fused=True, capturable=bool(graph))
                g.reducer, g._optimizer_kind = None, "torch"
        # densify / prune / reset_opacity re-lay the flat buffers out (GaussianModel._rebind): follow them
        g.on_relayout = self._on_relayout
        self.fused_opt = optimizer == "fused"
        g.optimizer = self.opt
        # TimeNet's weight gradients are accumulated by the kernels straight into the flat buffer
        g._timenet.direct_grads = True
        # ... and so are the gradients of the Gaussian / control-point / latent parameters (renderer.render_batch)
        renderer.direct_grads = self.fused_opt
        self.frame_w = None           # optional [S] device tensor: per-frame MSE weights (main_train_dimo.py:333-336)
        # depth / normal smoothness terms of the real step (main_train_dimo.py:363-372); off in the north-star step
        self.regularisers = bool(regularisers)
        # graph mode state
        self.use_graph = bool(graph)
        self.probe_steps = int(probe_steps)
        self.capacity_margin = float(capacity_margin)
        self.graph = None
        self.capacity = None
        self._seen = 0
        self._max_R = 0
        self._static = None
        self.graph_error = None
        self.poll_every = max(1, int(poll_every))
        self._replays = 0
        self._poll = None             # (event, pinned [4] floats) of the read in flight
        self.skipped_steps = 0        # replays discarded because of an overflow (as far as polled)
        self.recaptures = 0
        self._bind_skip_flag()

    def _bind_skip_flag(self):
        if self.fused_opt:
            self.opt.skip_flag = self.reducer.tail[0:1]

    def _on_relayout(self, model):
        """The Gaussian count (or a parameter tensor) changed: new flat buffers, new instance counts -- a captured graph
        is stale, so the next run() probes the instance count again and re-captures."""
        if model.reducer is None:
            raise RuntimeError("TrainStep follows re-layouts of the fused optimizer only (optimizer='torch' is an A/B "
                               "vehicle with a fixed parameter set)")
        self.reducer, self.opt = model.reducer, model.optimizer
        self.params = list(self.reducer.params)
        model._timenet.direct_grads = True
        self._bind_skip_flag()
        self._poll = None
        self.graph = None
        self._static = None
        self.capacity = None
        self._seen = 0
        self._max_R = 0

    def _timed(self, name, fn):
        if not _lib.PROFILE.enabled or torch.cuda.is_current_stream_capturing():
            return fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        _lib.PROFILE.events.append((name, e0, e1))
        return out

    # ------------------------------------------------------------------------------------------
    def _body(self, prep, gt, mask, n_motions, optimize=True, capacity=None, overflow_acc=None):
        if self.stage >= "s2":
            self.g.find_knn(4)
        # the loss kernel clamps the render to [0,1] on load (and masks the gradient), so skip the separate clamp pass
        out = self.r.render_batch(prepared=prep, stage=self.stage, clamp=False, capacity=capacity,
                                  with_visibility=False, depth_normal=self.regularisers, with_cpts=False)
        st = out["raster_state"]
        if capacity is None:
            self._max_R = max(self._max_R, st.R)
        elif overflow_acc is not None:
            overflow_acc.copy_(torch.maximum(overflow_acc, st.count_overflow))
            if self.fused_opt:       # this step's flag -> the all-reduced status word that gates the optimizer
                self.reducer.tail[0:1].copy_(st.count_overflow[1:2])
        loss = step_loss(out["image_raw"], out["alpha"], gt, mask, n_motions, frame_w=self.frame_w)
        if self.regularisers:
            loss = loss + _loss.smoothness_losses(out["image_raw"], out["depth"], out["normal"], groups=n_motions,
                                                  lambda_smooth=StepLossWeights.lambda_smooth,
                                                  lambda_bilateral=StepLossWeights.lambda_bilateral, clamp01=True)
        loss.backward()                                   # gradients accumulate straight into reducer.flat
        if self.world > 1:
            self._timed("py:allreduce_wait", self.reducer.reduce)
        if optimize and self.fused_opt:
            self.opt.step()                               # dimo_adam_step: update + gradient clear in one launch
            self.opt.zero_grad()
        else:
            if optimize:
                self._timed("py:adam", self.opt.step)
            self._timed("py:zero_grad", self.reducer.zero)
        return loss

    def _capture(self, prep, gt, mask, n_motions, optimize):
        dev = gt.device
        self.capacity = int(self._max_R * self.capacity_margin) + 1024
        self._static = {"prep": self.r.clone_prep(prep),
                        "gt": gt.clone(), "mask": mask.clone(),
                        "overflow": torch.zeros(2, dtype=torch.int32, device=dev), "n_motions": n_motions,
                        "optimize": optimize}
        st = self._static
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up in capacity mode (allocator, lazy inits);
            for _ in range(2):                             # no optimizer update: the step sequence stays unchanged
                self._body(st["prep"], st["gt"], st["mask"], n_motions, False, self.capacity, st["overflow"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            st["loss"] = self._body(st["prep"], st["gt"], st["mask"], n_motions, optimize, self.capacity, st["overflow"])
        self.graph = graph

    def _poll_overflow(self):
        """Every `poll_every` replays: read the (all-reduced, rank-identical) overflow word and this rank's largest
        instance count without blocking, and act on the previous read once it has landed.  Returns True when an
        overflow was seen and the captured graph was dropped (larger capacity, re-capture on the next run())."""
        st = self._static
        if self._replays % self.poll_every != 0:
            return False
        # the previous read was issued poll_every replays ago: waiting for it costs nothing, and acting at a fixed
        # replay index keeps all ranks in lock-step (a re-capture issues extra collectives)
        if self._poll is not None:
            self._poll[0].synchronize()
            host = self._poll[1]
            self._poll = None
            if float(host[0]) != 0.0:
                n_bad, need = int(host[0]), int(host[1])
                self.skipped_steps += 1
                warnings.warn(f"dimo_b200 TrainStep: the rasteriser's instance capacity ({self.capacity}) overflowed on "
                              f"{n_bad} rank(s) (largest count on this rank {need}); the optimizer discarded the "
                              "affected steps.  Growing the capacity and re-capturing the graph.")
                # every rank re-captures (the flag is identical everywhere); each grows by what IT saw, at least 25 %
                self._max_R = max(self._max_R, need, int(self.capacity * 1.25 / self.capacity_margin))
                self.graph = None
                self._static = None
                self.recaptures += 1
                self.reducer.tail.zero_()
                return True
        if self._poll is None:
            if "poll_host" not in st:
                st["poll_host"] = torch.zeros(4, dtype=torch.float32).pin_memory()
                st["poll_dev"] = torch.zeros(4, dtype=torch.float32, device=st["gt"].device)
            flag = self.reducer.tail[0:1] if self.fused_opt else st["overflow"][1:2]
            st["poll_dev"][0:1].copy_(flag)
            st["poll_dev"][1:2].copy_(st["overflow"][0:1])
            st["poll_host"].copy_(st["poll_dev"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._poll = (ev, st["poll_host"])
        return False

    def overflowed(self):
        """graph mode: (max instance count seen, capacity, overflow flag) -- one small device read."""
        if self._static is None:
            return self._max_R, None, False
        cnt, flag = self._static["overflow"].tolist()
        return cnt, self.capacity, bool(flag)

    # ------------------------------------------------------------------------------------------
    def run(self, cameras, times, latent_indices, gt, mask, n_motions, optimize=True):
        """cameras/times/latent_indices: length-S lists ordered motion-major; gt [S,3,H,W], mask [S,1,H,W] on device.
        Returns the (device) loss tensor."""
        if not self.use_graph or self.graph_error is not None:
            prep = self.r.prepare_step(cameras, times, latent_indices)
            return self._body(prep, gt, mask, n_motions, optimize)
        if self.graph is None:
            prep = self.r.prepare_step(cameras, times, latent_indices)
            if self._seen < self.probe_steps and self.recaptures == 0:    # eager probe steps: learn the instance count
                self._seen += 1
                return self._body(prep, gt, mask, n_motions, optimize)
            gen = torch.cuda.default_generators[gt.device.index if gt.device.index is not None
                                                else torch.cuda.current_device()]
            try:
                rng_backup = gen.clone_state()             # a failed capture leaves the generator in capture mode
            except Exception:
                rng_backup = None
            try:
                self._capture(prep, gt, mask, n_motions, optimize)
            except Exception as e:                         # stay correct: fall back to eager and say so
                self.graph_error = f"{type(e).__name__}: {e}"
                self.graph = None
                self._static = None
                torch.cuda.synchronize()
                if rng_backup is not None:
                    try:
                        gen.graphsafe_set_state(rng_backup.graphsafe_get_state())
                    except Exception:
                        pass
                warnings.warn("dimo_b200 TrainStep: CUDA-graph capture failed, running eagerly (" + self.graph_error +
                              ").  A common cause: tensors of an earlier backward (a render() result, a loss) are "
                              "still alive, so their AccumulateGrad nodes stay bound to the stream they were created on.")
                return self._body(prep, gt, mask, n_motions, optimize)
        if self._poll_overflow():                          # capacity grown: this step runs eagerly, the next one re-captures
            prep = self.r.prepare_step(cameras, times, latent_indices)
            return self._body(prep, gt, mask, n_motions, optimize)
        st = self._static
        self._replays += 1
        self.r.prepare_step(cameras, times, latent_indices, out=st["prep"])
        st["gt"].copy_(gt, non_blocking=True)
        st["mask"].copy_(mask, non_blocking=True)
        if self.fused_opt:
            self.opt.sync_lrs()                            # learning-rate changes reach the replay through device memory
        self.graph.replay()
        return st["loss"]


class RenderStep:
    """Forward-only launch set for 4-D inference (main_test_dimo.py:199-365: find_knn once, then per (view, frame)
    `render()` without backward): all S frames of a call go through ONE batched launch set; in graph mode the set is
    captured once (rasteriser in capacity mode) and replayed.  Returns the clamped images [S,3,H,W]."""

    def __init__(self, renderer: Renderer, stage="s2", graph=True, probe_steps=3, capacity_margin=1.25,
                 depth_normal=False):
        """depth_normal=False: images (+ alpha) only, as the inference loops consume them (main_test_dimo.py:243-260)."""
        self.depth_normal = bool(depth_normal)
        self.r, self.g, self.stage = renderer, renderer.gaussians, stage
        self.use_graph, self.probe_steps, self.capacity_margin = bool(graph), int(probe_steps), float(capacity_margin)
        self.graph, self._static, self.capacity = None, None, None
        self._seen, self._max_R = 0, 0
        if stage >= "s2":
            with torch.no_grad():
                self.g.find_knn(4)                 # once per run (main_test_dimo.py:213)

    def _body(self, prep, capacity=None):
        with torch.no_grad():
            out = self.r.render_batch(prepared=prep, stage=self.stage, clamp=True, capacity=capacity,
                                      with_visibility=False, depth_normal=self.depth_normal, with_cpts=False)
        st = out["raster_state"]
        if capacity is None:
            self._max_R = max(self._max_R, st.R)
        return out["image"], out["depth"], out["alpha"], st

    def run(self, cameras, times, latent_indices):
        if not self.use_graph:
            return self._body(self.r.prepare_step(cameras, times, latent_indices))[0]
        if self.graph is None:
            prep = self.r.prepare_step(cameras, times, latent_indices)
            if self._seen < self.probe_steps:
                self._seen += 1
                return self._body(prep)[0]
            self.capacity = int(self._max_R * self.capacity_margin) + 1024
            g = self.r.gaussians
            self._static = {"prep": self.r.clone_prep(prep),
                            "overflow": torch.zeros(2, dtype=torch.int32, device=prep["cams"].device),
                            # the captured launches read the model's KNN table by address: keep it alive even if the
                            # model is given a new one (the graph then renders with the table it was captured with)
                            "knn": (getattr(g, "neighbor_indices", None), getattr(g, "neighbor_dists", None))}
            stt = self._static
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._body(stt["prep"], self.capacity)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                img, depth, alpha, st = self._body(stt["prep"], self.capacity)
                stt["overflow"].copy_(torch.maximum(stt["overflow"], st.count_overflow))
                stt["image"] = img
            self.graph = graph
        stt = self._static
        self.r.prepare_step(cameras, times, latent_indices, out=stt["prep"])
        self.graph.replay()
        return stt["image"]

    def overflowed(self):
        if self._static is None:
            return self._max_R, None, False
        cnt, flag = self._static["overflow"].tolist()
        return cnt, self.capacity, bool(flag)
