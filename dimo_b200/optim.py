"""Optimizer of the loop: torch.optim.Adam over GaussianModel.training_setup's parameter groups
(renderer/latent_gs_renderer.py:453-476; stepped at main_train_dimo.py:416-417) as ONE kernel launch.

All parameters are re-homed as views of one flat fp32 buffer whose layout equals that of the flat gradient
buffer (`dist.FlatGradReducer`), so `dimo_adam_step` walks (param, grad, exp_avg, exp_avg_sq) with 128-bit
accesses and clears the gradient in the same pass (optimizer.zero_grad() folded in).  Learning rates are per
parameter group, live in a small device array and may change between steps (`update_learning_rate`, :502-520)
without re-capturing a CUDA graph; the step counter is device-resident for the same reason.

Same interface subset as torch.optim.Adam that the reference touches: `param_groups` (list of dicts with
"params", "lr", "name"), `step()`, `zero_grad()`, `state_dict()` / `load_state_dict()`.
"""
import ctypes

import torch

from . import _lib


class FusedAdam:
    def __init__(self, param_groups, reducer, betas=(0.9, 0.999), eps=1e-15, lr=0.0, fold_zero_grad=True):
        """param_groups: list of {"params": [...], "lr": float, "name": str} (or a flat list of tensors -> one group);
        reducer: the FlatGradReducer that owns the flat gradient buffer (defines the layout)."""
        if param_groups and not isinstance(param_groups[0], dict):
            param_groups = [{"params": list(param_groups), "lr": lr, "name": "all"}]
        self.param_groups = []
        for g in param_groups:
            ps = [p for p in g["params"] if p.numel() > 0]
            self.param_groups.append({"params": ps, "lr": float(g.get("lr", lr)), "name": g.get("name", "")})
        self.reducer = reducer
        self.betas = (float(betas[0]), float(betas[1]))
        self.eps = float(eps)
        self.fold_zero_grad = bool(fold_zero_grad)
        group_of = {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                group_of[id(p)] = gi
        dev = reducer.flat.device
        n = reducer.flat.numel()
        assert n % 4 == 0, "flat buffer must be padded to a multiple of 4 floats"
        # parameters become views of one flat buffer laid out like the gradient buffer
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        begins, groups = [], []
        for p, (off, numel) in zip(reducer.params, reducer.offsets):
            if id(p) not in group_of:
                raise ValueError("every parameter of the flat gradient buffer needs an optimizer group")
            view = self.flat[off:off + numel].view_as(p)
            view.copy_(p.data)
            p.data = view
            gi = group_of[id(p)]
            if groups and groups[-1] == gi:
                continue                      # same group as the previous tensor: extend the segment
            begins.append(off)
            groups.append(gi)
        begins.append(n)
        self._seg_group = groups
        # segments hold their group DICT: the reference pops groups from `param_groups` (main_train_dimo.py:489-493,
        # the shared-radius group "r" at the start of stage s2), which must not shift the other segments' rates
        self._seg_dict = [self.param_groups[g] for g in groups]
        self._seg_begin = (ctypes.c_int64 * len(begins))(*begins)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)     # [0] step count, [1] ticket
        # learning rates travel through a small ring of pinned staging buffers (asynchronous copies; a slot is
        # reused only after its previous copy has completed)
        self._lr_ring = None                                            # pinned, allocated on first use
        self._lr_events = [None] * 4
        self._lr_slot = 0
        self._lr_dev = torch.zeros(len(groups), dtype=torch.float32, device=dev)
        self._lr_sent = None
        # optional device float: non-zero -> dimo_adam_step discards the step's gradients (TrainStep points it at the
        # all-reduced capacity-overflow word of the flat buffer's tail)
        self.skip_flag = None
        self.sync_lrs()

    # ------------------------------------------------------------------------------------------
    def sync_lrs(self):
        """Uploads the groups' learning rates if they changed (call outside a graph capture / before a replay)."""
        cur = tuple(float(d["lr"]) for d in self._seg_dict)
        if cur != self._lr_sent:
            if self._lr_ring is None:
                self._lr_ring = [torch.zeros(len(cur), dtype=torch.float32).pin_memory() for _ in range(4)]
            k = self._lr_slot
            self._lr_slot = (k + 1) % len(self._lr_ring)
            if self._lr_events[k] is not None:
                self._lr_events[k].synchronize()
            host = self._lr_ring[k]
            for i, v in enumerate(cur):
                host[i] = v
            self._lr_dev.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._lr_events[k] = ev
            self._lr_sent = cur

    def step(self):
        self.reducer.dirty = False
        if not torch.cuda.is_current_stream_capturing():
            self.sync_lrs()
        _lib.call("dimo_adam_step", self.flat.numel(), _lib.ptr(self.flat), _lib.ptr(self.reducer.flat),
                  _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), len(self._seg_group), self._seg_begin,
                  _lib.ptr(self._lr_dev), self.betas[0], self.betas[1], self.eps, int(self.fold_zero_grad),
                  _lib.ptr(self.state), _lib.ptr(self.skip_flag), _lib.stream())

    # -- moment access for the surgery in gaussian_model.GaussianModel (densify / prune / reset_opacity) --------
    def _span(self, p):
        for q, (off, numel) in zip(self.reducer.params, self.reducer.offsets):
            if q is p:
                return off, numel
        raise KeyError("parameter is not part of this optimizer's flat buffer")

    def moments(self, p):
        """(exp_avg, exp_avg_sq) views shaped like `p`."""
        off, numel = self._span(p)
        return self.exp_avg[off:off + numel].view_as(p), self.exp_avg_sq[off:off + numel].view_as(p)

    def load_moments(self, by_id):
        """by_id: {id(param): (exp_avg, exp_avg_sq)}; parameters that are not listed keep zero moments."""
        for q, (off, numel) in zip(self.reducer.params, self.reducer.offsets):
            mv = by_id.get(id(q))
            if mv is not None:
                self.exp_avg[off:off + numel].copy_(mv[0].reshape(-1))
                self.exp_avg_sq[off:off + numel].copy_(mv[1].reshape(-1))

    def zero_grad(self, set_to_none=False):
        """No kernel when the clear is folded into step(); the reducer's bookkeeping is reset either way.  A backward
        whose gradients no step() consumed (a skipped / probe step; seen through the reducer's post-accumulate hooks
        on the per-Gaussian parameters) IS cleared, as torch.optim.Adam.zero_grad would."""
        if self.fold_zero_grad:
            if getattr(self.reducer, "dirty", False):
                self.reducer.zero()
                self.reducer.dirty = False
            self.reducer.reset()
        else:
            self.reducer.zero()

    # ------------------------------------------------------------------------------------------
    def state_dict(self):
        return {"step": int(self.state[0].item()), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lrs": [g["lr"] for g in self.param_groups], "names": [g["name"] for g in self.param_groups]}

    def load_state_dict(self, sd):
        """Accepts this class's own layout (state_dict above) and torch.optim.Adam's ({"state", "param_groups"}: a
        checkpoint written by the reference, `capture()` at renderer/latent_gs_renderer.py:296-315, or by the
        optimizer='torch' path).  torch keeps one step counter per parameter, this optimizer one for all: the largest
        is taken (they are equal whenever every parameter received a gradient in every step, as on this path)."""
        if "state" in sd and "param_groups" in sd:
            by_name = {g.get("name", k): g for k, g in enumerate(sd["param_groups"])}
            mom, step = {}, 0
            for k, g in enumerate(self.param_groups):
                tg = by_name.get(g["name"], sd["param_groups"][k] if k < len(sd["param_groups"]) else None)
                if tg is None:
                    continue
                g["lr"] = float(tg.get("lr", g["lr"]))
                live = [p for p in g["params"] if p.numel() > 0]
                for p, idx in zip(live, tg["params"]):
                    st = sd["state"].get(idx)
                    if st is None:
                        continue
                    if tuple(st["exp_avg"].shape) != tuple(p.shape):
                        raise ValueError(f"optimizer state of group '{g['name']}' has shape {tuple(st['exp_avg'].shape)}, "
                                         f"the parameter {tuple(p.shape)}")
                    mom[id(p)] = (st["exp_avg"].to(self.flat.device, torch.float32),
                                  st["exp_avg_sq"].to(self.flat.device, torch.float32))
                    step = max(step, int(float(st["step"])))
            self.exp_avg.zero_(); self.exp_avg_sq.zero_()
            self.load_moments(mom)
            self.state.zero_(); self.state[0] = step
            self.sync_lrs()
            return
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.state.zero_(); self.state[0] = int(sd["step"])
        for g, lr in zip(self.param_groups, sd["lrs"]):
            g["lr"] = float(lr)
        self.sync_lrs()

    def torch_state_dict(self):
        """The same state in torch.optim.Adam's layout (interchangeable checkpoints): parameters numbered group by
        group, every parameter with the global step."""
        state, groups, k = {}, [], 0
        step = torch.tensor(float(int(self.state[0].item())))
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                if p.numel() == 0:
                    continue
                m, v = self.moments(p)
                state[k] = {"step": step.clone(), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
                ids.append(k); k += 1
            groups.append({"lr": g["lr"], "name": g["name"], "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0,
                           "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                           "differentiable": False, "fused": None, "decoupled_weight_decay": False, "params": ids})
        return {"state": state, "param_groups": groups}
