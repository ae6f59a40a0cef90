"""Host logic of the renderer-class surface on the CPU against fixtures produced by the
reference's own classes (tests/golden/make_golden_model.py): GaussianModel life cycle, ARAP energy, PLY / .pth files.
No kernels are involved (the model runs with optimizer="torch" and an injected 3-NN distance function)."""
import functools
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import model_scenario  # noqa: E402
from dimo_b200 import ply, regularisers  # noqa: E402
from dimo_b200.gaussian_model import GaussianModel, get_expon_lr_func  # noqa: E402
from dimo_b200.renderer import Renderer  # noqa: E402
from oracle import knn as oknn  # noqa: E402
from oracle import points as opoints  # noqa: E402

GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def scenario():
    cls = functools.partial(Renderer, device="cpu")
    return model_scenario.run(cls, init_kwargs={"dist3nn": lambda p: oknn.dist3nn(p.float())},
                              setup_kwargs={"optimizer": "torch"})


def test_lifecycle_matches_reference_classes(scenario):
    gold = np.load(os.path.join(GOLD, "model.npz"))
    assert sorted(gold.files) == sorted(scenario.keys())
    worst = 0.0
    for k in gold.files:
        a, b = gold[k], scenario[k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if a.dtype.kind in "US":
            assert list(a) == list(b), k
            continue
        if a.size == 0:
            continue
        if k.startswith("sched/") or k.endswith("/lrs") or k.endswith("group_sizes"):
            assert np.array_equal(a, b), k            # float64 schedule arithmetic / integers: identical
            continue
        err = float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
        scale = float(np.abs(a).max()) + 1e-12
        worst = max(worst, err / scale)
        assert err <= 1e-6 * scale + 1e-9, (k, err, scale)
    # row selection, append order and the split's random stream are the reference's: sizes agree exactly
    assert scenario["densified/_xyz"].shape[0] == 784 and scenario["fps_pruned/_xyz"].shape[0] == 40


def test_index_prune_keeps_mirrored_rows():
    """GUI.FPS hands prune_points an int64 index tensor; `~idx` selects row N-1-idx (main_train_dimo.py:511-515)."""
    g = GaussianModel(0, device="cpu")
    n = 20
    st = {"_xyz": torch.arange(n * 3).float().reshape(n, 3), "_features_dc": torch.zeros(n, 1, 3),
          "_features_rest": torch.zeros(n, 0, 3), "_scaling": torch.zeros(n, 3), "_rotation": torch.zeros(n, 4),
          "_opacity": torch.zeros(n, 1), "_c_xyz": torch.zeros(4, 3), "_c_radius": torch.zeros(4, 1)}
    g.load_state(st)
    g.spatial_lr_scale = 1
    g.training_setup(model_scenario.train_args(), optimizer="torch")
    idx = torch.tensor([0, 5, 19])
    g.prune_points(idx)
    assert g._xyz.shape[0] == 3
    assert torch.equal(g._xyz.detach()[:, 0], torch.tensor([19.0, 14.0, 0.0]) * 3)


def test_expon_lr_func_edges():
    assert get_expon_lr_func(1e-3, 1e-3)(123) == 1e-3
    assert get_expon_lr_func(0.0, 0.0)(5) == 0.0
    f = get_expon_lr_func(1e-2, 1e-4, max_steps=100)
    assert f(-1) == 0.0 and abs(f(0) - 1e-2) < 1e-16 and abs(f(100) - 1e-4) < 1e-16 and abs(f(1000) - 1e-4) < 1e-16
    assert abs(f(50) - 1e-3) < 1e-15


def test_cpu_model_refuses_fused_optimizer():
    g = GaussianModel(0, device="cpu")
    g.load_state({"_xyz": torch.zeros(4, 3), "_features_dc": torch.zeros(4, 1, 3), "_features_rest": torch.zeros(4, 0, 3),
                  "_scaling": torch.zeros(4, 3), "_rotation": torch.zeros(4, 4), "_opacity": torch.zeros(4, 1),
                  "_c_xyz": torch.zeros(2, 3), "_c_radius": torch.zeros(2, 1)})
    with pytest.raises(RuntimeError):
        g.training_setup(model_scenario.train_args())           # no silent CPU fallback for the product optimizer


# ------------------------------------------------------------------------------------------------------------------
# ARAP
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_arap_matches_reference_functions(tag):
    gold = np.load(os.path.join(GOLD, "arap.npz"))
    nodes = torch.from_numpy(gold[f"{tag}/nodes"]).requires_grad_(True)
    ref_edges = sorted(zip(gold[f"{tag}/ii"].tolist(), gold[f"{tag}/jj"].tolist()))
    # oracle restatement (reference-shaped) against the reference's output
    ii, jj, nn = opoints.arap_connectivity_v2(nodes.detach())
    assert sorted(zip(ii.tolist(), jj.tolist())) == ref_edges
    e_or = opoints.arap_error(nodes, ii, jj, nn)
    assert abs(e_or.item() - float(gold[f"{tag}/error"])) <= 1e-5 * float(gold[f"{tag}/error"])
    # product formulation (neighbour table, frames batched), ball query served by the oracle on the CPU
    err, (pi, pj, pn, nbr) = regularisers.arap_loss_points(nodes, ball_query=opoints.ball_query)
    assert sorted(zip(pi.tolist(), pj.tolist())) == ref_edges
    assert int((nbr >= 0).sum()) == len(ref_edges)
    assert abs(err.item() - float(gold[f"{tag}/error"])) <= 2e-5 * float(gold[f"{tag}/error"])
    (grad,) = torch.autograd.grad(err, nodes)
    g_ref = gold[f"{tag}/grad"]
    assert np.abs(grad.numpy() - g_ref).max() <= 1e-4 * np.abs(g_ref).max()


def test_arap_rigid_motion_has_zero_energy():
    g = torch.Generator().manual_seed(3)
    base = (torch.rand(80, 3, generator=g) - 0.5) * 0.4
    ang = 0.7
    Ry = torch.tensor([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], dtype=torch.float32)
    Rx = torch.tensor([[1, 0, 0], [0, np.cos(0.4), -np.sin(0.4)], [0, np.sin(0.4), np.cos(0.4)]], dtype=torch.float32)
    # a general axis: a rotation about a coordinate axis leaves that coordinate of every edge bit-identical, which the
    # reference treats as "vertex unchanged" and pins R = I (deform_utils.py:175-176)
    R = Ry @ Rx
    nodes = torch.stack([base, base @ R.T + torch.tensor([0.1, -0.2, 0.05]), base])
    err, (ii, _jj, _nn, _nbr) = regularisers.arap_loss_points(nodes, ball_query=opoints.ball_query)
    assert len(ii) > 50
    assert err.item() < 1e-9


def test_arap_sampling_with_replacement():
    g = torch.Generator().manual_seed(4)
    nodes = (torch.rand(3, 40, 3, generator=g) - 0.5) * 0.3
    _ii, _jj, _nn, nbr = regularisers.connectivity_v2(nodes, ball_query=opoints.ball_query)
    full = regularisers.arap_energy(nodes, nbr)
    idx = torch.tensor([0, 0, 5, 7, 7, 7])
    part = regularisers.arap_energy(nodes, nbr, sample_idx=idx)
    per = [regularisers.arap_energy(nodes, nbr, sample_idx=torch.tensor([i])).item() for i in range(40)]
    assert abs(sum(per) - full.item()) <= 1e-5 * full.item()
    assert abs(part.item() - (2 * per[0] + per[5] + 3 * per[7])) <= 1e-6


# ------------------------------------------------------------------------------------------------------------------
# oracle point ops: known answers (pytorch3d / chamferdist are absent: parity unpinned, semantics restated)
# ------------------------------------------------------------------------------------------------------------------
def test_oracle_fps_known_answer():
    pts = torch.tensor([[0.0, 0, 0], [1, 0, 0], [0.4, 0, 0], [3, 0, 0], [-2, 0, 0], [3, 0, 0]])
    assert opoints.fps(pts, 4).tolist() == [0, 3, 4, 1]          # ties (3 and 5 coincide) -> lower index
    idx = opoints.fps(torch.randn(200, 3, generator=torch.Generator().manual_seed(0)), 50)
    assert len(set(idx.tolist())) == 50 and idx[0] == 0


def test_oracle_ball_query_known_answer():
    p = torch.tensor([[[0.0, 0, 0], [0.05, 0, 0], [0.2, 0, 0], [0.0, 0.09, 0], [0.0, 0.1, 0]]])
    d, idx, nn = opoints.ball_query(p, p, K=3, radius=0.1)
    assert idx[0, 0].tolist() == [0, 1, 3]                        # index order, strict < radius: point 4 is outside
    assert idx[0, 2].tolist() == [2, -1, -1]
    assert torch.allclose(d[0, 0], torch.tensor([0.0, 0.0025, 0.0081]), atol=1e-7)
    assert torch.equal(nn[0, 2, 1], torch.zeros(3))
    d2, idx2, _ = opoints.ball_query(p, p, K=2, radius=0.1)
    assert idx2[0, 0].tolist() == [0, 1]                          # truncated at K, still index order


def test_oracle_chamfer_gradient():
    g = torch.Generator().manual_seed(1)
    a = torch.randn(30, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    b = torch.randn(20, 3, generator=g, dtype=torch.float64).requires_grad_(True)
    v = opoints.chamfer_forward(a, b)
    d2 = ((a[:, None] - b[None]) ** 2).sum(-1)
    assert abs(v.item() - d2.min(dim=1).values.sum().item()) < 1e-12
    assert torch.autograd.gradcheck(opoints.chamfer_forward, (a, b), eps=1e-7, atol=1e-6)


# ------------------------------------------------------------------------------------------------------------------
# files
# ------------------------------------------------------------------------------------------------------------------
def _tiny_model(n=9, m=4, sh_degree=0, seed=0, vae=False):
    gen = torch.Generator().manual_seed(seed)
    g = GaussianModel(sh_degree, num_latent_code=3, device="cpu", vae_latent=vae)
    rest = (sh_degree + 1) ** 2 - 1
    g.load_state({"_xyz": torch.randn(n, 3, generator=gen), "_features_dc": torch.randn(n, 1, 3, generator=gen),
                  "_features_rest": torch.randn(n, rest, 3, generator=gen), "_scaling": torch.randn(n, 3, generator=gen),
                  "_rotation": torch.randn(n, 4, generator=gen), "_opacity": torch.randn(n, 1, generator=gen),
                  "_c_xyz": torch.randn(m, 3, generator=gen), "_c_radius": torch.randn(m, 1, generator=gen)})
    return g


@pytest.mark.parametrize("sh_degree", [0, 2])
def test_ply_round_trip_and_layout(tmp_path, sh_degree):
    g = _tiny_model(sh_degree=sh_degree)
    p1, p2 = str(tmp_path / "s2" / "point_cloud.ply"), str(tmp_path / "s2" / "point_cloud_c.ply")
    g.save_ply(p1, p2)
    raw = open(p1, "rb").read()
    header, body = raw.split(b"end_header\n", 1)
    lines = header.decode().strip().split("\n")
    n_rest = 3 * ((sh_degree + 1) ** 2 - 1)
    names = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(n_rest)] \
        + ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 9"]
    assert lines[3:] == [f"property float {n}" for n in names]                 # renderer/latent_gs_renderer.py:517-529
    assert len(body) == 9 * 4 * len(names)
    row0 = np.frombuffer(body, dtype="<f4", count=len(names))
    assert np.array_equal(row0[:3], g._xyz.detach().numpy()[0]) and np.all(row0[3:6] == 0)
    # features are stored channel-major: transpose(1, 2).flatten (:545-546)
    assert np.array_equal(row0[9:9 + n_rest], g._features_rest.detach().transpose(1, 2).flatten(1).numpy()[0])
    h = _tiny_model(sh_degree=sh_degree, seed=9)
    h.load_ply(p1, p2)
    for k in ("_xyz", "_features_dc", "_features_rest", "_scaling", "_rotation", "_opacity", "_c_xyz", "_c_radius"):
        assert torch.equal(getattr(h, k).detach(), getattr(g, k).detach()), k
        assert getattr(h, k).requires_grad
    assert h.active_sh_degree == sh_degree


def test_ply_shared_radius_written_as_scale(tmp_path):
    g = _tiny_model()
    g._r = torch.nn.Parameter(torch.tensor([[-2.5]]))
    p1 = str(tmp_path / "pc.ply")
    g.save_ply(p1)
    v = ply.read_ply(p1).first
    assert np.all(v["scale_0"] == np.float32(-2.5)) and np.all(v["scale_2"] == np.float32(-2.5))     # :550-551


def test_ply_reader_ascii_and_big_endian(tmp_path):
    txt = "ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\nproperty float x\nproperty double y\n" \
          "property uchar z\nend_header\n1.5 2.25 7\n-3 4e-1 255\n"
    p = tmp_path / "a.ply"
    p.write_text(txt)
    v = ply.read_ply(str(p)).first
    assert v["x"].tolist() == [1.5, -3.0] and v["y"].tolist() == [2.25, 0.4] and v["z"].tolist() == [7, 255]
    be = b"ply\nformat binary_big_endian 1.0\nelement vertex 1\nproperty float x\nproperty int k\nend_header\n" \
        + np.array([2.5], ">f4").tobytes() + np.array([-7], ">i4").tobytes()
    q = tmp_path / "b.ply"
    q.write_bytes(be)
    w = ply.read_ply(str(q)).first
    assert w["x"][0] == 2.5 and w["k"][0] == -7
    with pytest.raises(ValueError):
        (tmp_path / "c.ply").write_bytes(be[:-3])
        ply.read_ply(str(tmp_path / "c.ply"))
    with pytest.raises(ValueError):
        (tmp_path / "d.ply").write_text("ply\nformat ascii 1.0\nelement face 1\nproperty list uchar int vertex_indices\n"
                                        "end_header\n3 0 1 2\n")
        ply.read_ply(str(tmp_path / "d.ply"))


def test_plyfile_shim_round_trip(tmp_path):
    import dimo_b200
    dimo_b200.install_shims()
    from plyfile import PlyData, PlyElement
    arr = np.empty(3, dtype=[("c_x", "f4"), ("c_y", "f4"), ("c_z", "f4"), ("c_radius", "f4")])
    arr[:] = [(1, 2, 3, 4), (5, 6, 7, 8), (9, 10, 11, 12)]
    path = str(tmp_path / "c.ply")
    PlyData([PlyElement.describe(arr, "vertex")]).write(path)
    back = PlyData.read(path)
    assert [p.name for p in back.elements[0].properties] == ["c_x", "c_y", "c_z", "c_radius"]
    assert np.asarray(back.elements[0]["c_radius"]).tolist() == [4.0, 8.0, 12.0]


@pytest.mark.parametrize("vae", [False, True])
def test_save_load_model_files(tmp_path, vae):
    g = _tiny_model(vae=vae)
    with torch.no_grad():
        for p in g.latent_parameters():
            p.copy_(torch.randn(p.shape))
        g._timenet.pts_layers[-1].weight.normal_()
    g.save_model(str(tmp_path), step=500)
    g.save_model(str(tmp_path))
    want = {"timenet.pth", "timenet_500.pth"} | ({"mu.pth", "log_var.pth", "mu_500.pth", "log_var_500.pth"} if vae
                                                  else {"latent_codes.pth", "latent_codes_500.pth"})
    assert set(os.listdir(tmp_path)) == want                                   # file names: :629-635 / gaussian twin
    h = _tiny_model(seed=5, vae=vae)
    h.load_model(str(tmp_path), step=500)
    for a, b in zip(g.latent_parameters(), h.latent_parameters()):
        assert torch.equal(a.detach(), b.detach()) and b.requires_grad
    for (ka, va), (kb, vb) in zip(g._timenet.state_dict().items(), h._timenet.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    # the reference's parameter names, so a released timenet.pth loads unchanged
    assert "deformnet.0.weight" in g._timenet.state_dict() and "rot_layers.2.bias" in g._timenet.state_dict()


def test_capture_restore(tmp_path):
    g = _tiny_model()
    g.spatial_lr_scale = 1
    g.training_setup(model_scenario.train_args(), optimizer="torch")
    gen = torch.Generator().manual_seed(0)
    model_scenario.fake_backward(g, gen, 0.01)
    g.optimizer.step()
    blob = g.capture()
    assert len(blob) == 17
    h = _tiny_model(seed=3)
    h.restore(blob, model_scenario.train_args(), optimizer="torch")
    assert torch.equal(h._xyz, g._xyz)
    st_g, st_h = g.optimizer.state[g._xyz], h.optimizer.state[h._xyz]
    assert torch.equal(st_g["exp_avg"], st_h["exp_avg"])


def test_vae_groups_and_reparameterize():
    r = Renderer(sh_degree=0, device="cpu", vae_latent=True, num_latent_code=4)
    g = r.gaussians
    assert g._mu.shape == (4, 32) and float(g._mu.abs().sum()) == 0.0 and float(g._log_var.abs().sum()) == 0.0
    names = [grp["name"] for grp in g.param_groups(0.0)]
    assert names[6:8] == ["latent_code_mu", "latent_code_log_var"] and "latent_code" not in names
    torch.manual_seed(0)
    z = r.reparameterize(torch.ones(1000, 32) * 2, torch.full((1000, 32), float(np.log(0.25))))
    assert abs(z.mean().item() - 2) < 0.02 and abs(z.std().item() - 0.5) < 0.02
