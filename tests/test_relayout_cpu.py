"""Host logic of the flat re-layout (GaussianModel._rebind on the fused optimizer, TrainStep following it) on the CPU.
The kernel launch of FusedAdam.step and its pinned learning-rate upload are replaced by a torch restatement of the same
update (oracle/optim.py) -- everything else (flat parameter / gradient / moment buffers, segment tables, moment
carry-over, hook re-attachment) is the product code.  The real kernel runs the same life in tests/test_widen_gpu.py."""
import os
import sys
import types

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import model_scenario as ms  # noqa: E402
from dimo_b200 import optim as doptim  # noqa: E402
from dimo_b200 import synthetic  # noqa: E402
from dimo_b200.gaussian_model import GaussianModel  # noqa: E402
from oracle import optim as oopt  # noqa: E402


class HostFusedAdam(doptim.FusedAdam):
    """FusedAdam with the device pieces (dimo_adam_step, pinned lr ring) served on the host."""

    def sync_lrs(self):
        self._lr_sent = tuple(float(d["lr"]) for d in self._seg_dict)

    def step(self):
        self.sync_lrs()
        n = len(self._seg_dict)
        step = int(self.state[0]) + 1
        for k in range(n):
            lo, hi = self._seg_begin[k], self._seg_begin[k + 1]
            p, g = self.flat[lo:hi], self.reducer.flat[lo:hi]
            oopt.adam_step([p], [g], [self.exp_avg[lo:hi]], [self.exp_avg_sq[lo:hi]], step, [self._lr_sent[k]],
                           beta1=self.betas[0], beta2=self.betas[1], eps=self.eps)
        self.state[0] = step
        if self.fold_zero_grad:
            self.reducer.flat.zero_()


@pytest.fixture()
def host_fused(monkeypatch):
    monkeypatch.setattr(doptim, "FusedAdam", HostFusedAdam)


def _model(kind, n=600):
    torch.manual_seed(7)                   # TimeNet's xavier init: both optimizer kinds start from the same weights
    g = GaussianModel(0, num_latent_code=2, device="cpu")
    g.load_state(synthetic.make_scene(n, n_ctrl=16, n_motions=2, seed=4))
    g.spatial_lr_scale = 1
    g.training_setup(ms.train_args(), optimizer=kind)
    return g


def _life(kind):
    g = _model(kind)
    gen = torch.Generator().manual_seed(0)

    def step(k=1):
        for _ in range(k):
            for grp in g.optimizer.param_groups:
                for p in grp["params"]:
                    gr = 0.01 * torch.randn(p.shape, generator=gen)
                    if p.grad is None:
                        p.grad = gr
                    else:
                        p.grad.copy_(gr)
            g.optimizer.step()
            g.optimizer.zero_grad()

    step(3)
    yield "steps", g
    n = g._xyz.shape[0]
    for _ in range(3):
        vs = types.SimpleNamespace(grad=0.03 * torch.randn(n, 3, generator=gen))
        vis = torch.rand(n, generator=gen) > 0.25
        radii = torch.randint(0, 3, (n,), generator=gen).float()
        g.max_radii2D[vis] = torch.max(g.max_radii2D[vis], radii[vis])
        g.add_densification_stats(vs, vis)
    with torch.no_grad():
        g._scaling.data.copy_(torch.log(torch.rand(n, 3, generator=gen) * 0.08 + 0.005))
    torch.manual_seed(31)
    g.densify_and_prune(0.02, min_opacity=0.1, extent=4, max_screen_size=1)
    yield "densified", g
    step(2)
    yield "steps2", g
    with torch.no_grad():
        g._opacity.data[::5] = -7.0
    g.prune(min_opacity=0.01, extent=4)
    yield "pruned", g
    g.prune_points(torch.randperm(g._xyz.shape[0], generator=gen)[:100])
    yield "index_pruned", g
    step(1)
    g.reset_opacity()
    yield "reset", g
    step(2)
    yield "steps3", g


def _state(g, kind):
    out = {}
    for name in ("_xyz", "_features_dc", "_opacity", "_scaling", "_rotation", "_c_xyz", "_c_radius", "_latent_codes"):
        p = getattr(g, name)
        out[name] = p.detach().clone()
        if kind == "fused":
            m, v = g.optimizer.moments(p)
        else:
            st = g.optimizer.state[p]
            m, v = st["exp_avg"], st["exp_avg_sq"]
        out[name + "/m"], out[name + "/v"] = m.detach().clone(), v.detach().clone()
    out["timenet"] = g._timenet.deformnet[3].weight.detach().clone()
    for name in ("max_radii2D", "xyz_gradient_accum", "denom"):
        out[name] = getattr(g, name).detach().clone()
    return out


def test_flat_relayout_follows_reference_surgery(host_fused):
    sizes, relayouts = [], []
    for (tag_f, gf), (tag_t, gt) in zip(_life("fused"), _life("torch")):
        gf.on_relayout = lambda m: relayouts.append(m._xyz.shape[0])
        a, b = _state(gf, "fused"), _state(gt, "torch")
        sizes.append(a["_xyz"].shape[0])
        for k in a:
            assert a[k].shape == b[k].shape, (tag_f, k)
            if a[k].numel():
                scale = float(b[k].abs().max()) + 1e-12
                assert float((a[k] - b[k]).abs().max()) <= 1e-5 * scale + 1e-9, (tag_f, k)
        flat, gflat = gf.optimizer.flat, gf.reducer.flat
        for p in gf.reducer.params:
            assert flat.data_ptr() <= p.data_ptr() < flat.data_ptr() + flat.numel() * 4, tag_f
            assert gflat.data_ptr() <= p.grad.data_ptr() < gflat.data_ptr() + gflat.numel() * 4, tag_f
        assert float(gflat.abs().max()) == 0.0
        # one segment table entry per group that owns tensors, offsets 16-byte aligned
        assert all(b % 4 == 0 for b in gf.optimizer._seg_begin)
    assert sizes[0] == 600 and sizes[1] != 600 and sizes[4] == 100
    assert int(gf.optimizer.state[0]) == 8
    # densify_and_prune (append, append, prune, prune), prune, index prune, reset: ONE re-layout each
    assert len(relayouts) == 4, relayouts


def test_popped_group_keeps_other_rates(host_fused):
    """prepare_train_s2 pops the shared-radius group from param_groups (main_train_dimo.py:487-493); the remaining
    segments must keep their own rates."""
    g = _model("fused")
    opt = g.optimizer
    names = [d["name"] for d in opt._seg_dict]
    rid = [i for i, grp in enumerate(opt.param_groups) if grp["name"] == "r"][0]
    opt.param_groups[rid]["lr"] = 0.0
    opt.param_groups.pop(rid)
    for grp in opt.param_groups:
        if grp["name"] == "c_xyz":
            grp["lr"] = 0.125
    opt.sync_lrs()
    assert opt._lr_sent[names.index("c_xyz")] == 0.125
    assert opt._lr_sent[names.index("xyz")] == pytest.approx(0.01)
    g.reset_opacity()                      # a re-layout after the pop: the popped group stays out, rates survive
    names2 = [d["name"] for d in g.optimizer._seg_dict]
    assert g.optimizer._lr_sent[names2.index("c_xyz")] == 0.125


def test_trainstep_follows_relayout(host_fused):
    from dimo_b200.renderer import Renderer
    from dimo_b200.trainstep import TrainStep
    r = Renderer(sh_degree=0, device="cpu", num_latent_code=2)
    g = r.gaussians
    g.load_state(synthetic.make_scene(300, n_ctrl=16, n_motions=2, seed=4))
    g.spatial_lr_scale = 1
    g.training_setup(ms.train_args(), optimizer="fused")
    ts = TrainStep(r, stage="s2", graph=True)
    assert ts.opt is g.optimizer and ts.reducer is g.reducer
    ts.graph, ts._seen, ts._max_R = object(), 3, 12345           # pretend a graph was captured
    with torch.no_grad():
        g._opacity.data[::3] = -9.0
    g.prune(min_opacity=0.01, extent=4)
    assert g._xyz.shape[0] == 200
    assert ts.opt is g.optimizer and ts.reducer is g.reducer and ts.graph is None and ts._seen == 0 and ts._max_R == 0
    assert all(p.grad is not None for p in ts.params)
