"""CPU: self-consistency and known-answer tests of the rasteriser oracle (parity unpinned upstream, so the
oracle is pinned analytically instead: closed forms, ordering, tile geometry, fp64 gradcheck)."""
import math

import pytest
import torch

from oracle import camera as ocam, raster as orast, sh as osh, knn as oknn


def _cam(W=64, H=64, view=0):
    return ocam.orbit_cam(view, 8, W, H)


def _raster(cam, xyz, scales, rot, op, rgb, W, H, bg=(0., 0., 0.), dtype=torch.float32, **kw):
    return orast.rasterize(xyz.to(dtype), scales.to(dtype), rot.to(dtype), op.to(dtype), cam.world_view_transform,
                           cam.full_proj_transform, cam.camera_center, cam.tanfovx, cam.tanfovy, W, H,
                           torch.tensor(bg, dtype=dtype), colors_precomp=rgb.to(dtype), **kw)


def test_single_isotropic_gaussian_closed_form():
    W = H = 64
    cam = _cam(W, H)
    s = 0.05
    xyz = torch.zeros(1, 3); scales = torch.full((1, 3), s); rot = torch.tensor([[1., 0, 0, 0]])
    op = torch.tensor([[0.8]]); rgb = torch.tensor([[0.2, 0.5, 0.9]])
    r = _raster(cam, xyz, scales, rot, op, rgb, W, H)
    # camera at distance 2 on +z looking at the origin: depth 2, centre pixel (W-1)/2
    fx = W / (2 * cam.tanfovx)
    var = (fx * s / 2.0) ** 2 + orast.DILATION
    assert abs(r["pre"]["cov2d"][0, 0].item() - var) < 1e-3 * var
    assert abs(r["pre"]["xy"][0, 0].item() - (W - 1) / 2) < 1e-3
    assert r["radii"][0].item() == math.ceil(3 * math.sqrt(var))
    px, py = 31, 31
    d2 = (px - (W - 1) / 2) ** 2 + (py - (H - 1) / 2) ** 2
    alpha = 0.8 * math.exp(-0.5 * d2 / var)
    assert abs(r["alpha"][0, py, px].item() - alpha) < 1e-5
    assert torch.allclose(r["image"][:, py, px], alpha * rgb[0], atol=1e-5)
    assert abs(r["depth"][0, py, px].item() - alpha * 2.0) < 1e-4
    assert r["n_contrib"][py, px].item() == 1


def test_front_to_back_order_and_background():
    W = H = 32
    cam = _cam(W, H)
    xyz = torch.tensor([[0., 0, 0.3], [0., 0, -0.3]])      # first is nearer to the camera (z=+2)
    scales = torch.full((2, 3), 0.2); rot = torch.tensor([[1., 0, 0, 0]] * 2)
    op = torch.tensor([[0.6], [0.7]]); rgb = torch.tensor([[1., 0, 0], [0, 1., 0]])
    bg = (0.1, 0.2, 0.3)
    r = _raster(cam, xyz, scales, rot, op, rgb, W, H, bg=bg)
    assert r["ids"][r["ranges"][0, 0]:r["ranges"][0, 1]].tolist()[0] == 0       # nearer first
    c = 15
    pre = r["pre"]
    def a(i):
        dx = pre["xy"][i, 0] - c; dy = pre["xy"][i, 1] - c
        p = -0.5 * (pre["conic"][i, 0] * dx * dx + pre["conic"][i, 2] * dy * dy) - pre["conic"][i, 1] * dx * dy
        return min(0.99, (pre["opacity"][i] * torch.exp(p)).item())
    a0, a1 = a(0), a(1)
    T = (1 - a0) * (1 - a1)
    want = torch.tensor([a0 * 1.0 + T * bg[0], (1 - a0) * a1 + T * bg[1], T * bg[2]])
    assert torch.allclose(r["image"][:, c, c], want, atol=1e-5)
    assert abs(r["alpha"][0, c, c].item() - (1 - T)) < 1e-6


def test_tile_rect_and_keys():
    W, H = 64, 48
    cam = _cam(W, H)
    xyz = torch.tensor([[0.0, 0.0, 0.0]]); scales = torch.full((1, 3), 0.02)
    r = _raster(cam, xyz, scales, torch.tensor([[1., 0, 0, 0]]), torch.tensor([[0.5]]), torch.ones(1, 3), W, H)
    x, y = r["pre"]["xy"][0].tolist(); rad = r["radii"][0].item()
    gx, gy = 4, 3
    x0 = min(gx, max(0, int((x - rad) / 16))); x1 = min(gx, max(0, int((x + rad + 15) / 16)))
    y0 = min(gy, max(0, int((y - rad) / 16))); y1 = min(gy, max(0, int((y + rad + 15) / 16)))
    assert r["pre"]["rect"][0].tolist() == [x0, y0, x1, y1]
    assert r["tiles_touched"][0].item() == (x1 - x0) * (y1 - y0) == r["keys"].numel()
    tiles = sorted(ty * gx + tx for ty in range(y0, y1) for tx in range(x0, x1))
    assert (r["keys"] >> 32).tolist() == tiles
    depth_bits = torch.tensor([2.0]).view(torch.int32).item()
    assert all((k & 0xFFFFFFFF) == depth_bits for k in r["keys"].tolist())


def test_culling():
    W = H = 32
    cam = _cam(W, H)
    xyz = torch.tensor([[0., 0, 1.9], [0., 0, 5.0], [50., 0, 0]])   # z_view = 0.1 (near-culled), behind, far off-screen
    r = _raster(cam, xyz, torch.full((3, 3), 0.01), torch.tensor([[1., 0, 0, 0]] * 3), torch.full((3, 1), 0.5),
                torch.ones(3, 3), W, H, bg=(1., 1., 1.))
    assert r["radii"].tolist() == [0, 0, 0] and r["keys"].numel() == 0
    assert torch.equal(r["image"], torch.ones(3, H, W)) and float(r["alpha"].abs().max()) == 0


def test_gradcheck_fp64():
    """autograd through the oracle agrees with finite differences (fp64), i.e. the gradient oracle is sound"""
    torch.manual_seed(0)
    W = H = 16
    cam = _cam(W, H)
    N = 6
    xyz = ((torch.rand(N, 3) - 0.5) * 0.4).double().requires_grad_(True)
    scales = (0.05 + 0.1 * torch.rand(N, 3)).double().requires_grad_(True)
    rot = torch.nn.functional.normalize(torch.randn(N, 4)).double().requires_grad_(True)
    op = (0.3 + 0.5 * torch.rand(N, 1)).double().requires_grad_(True)
    shs = torch.randn(N, 4, 3).double().requires_grad_(True)
    wts = torch.rand(8, H, W).double()

    def f(xyz, scales, rot, op, shs):
        r = orast.rasterize(xyz, scales, rot, op, cam.world_view_transform, cam.full_proj_transform,
                            cam.camera_center, cam.tanfovx, cam.tanfovy, W, H, torch.tensor([0.3, 0.6, 0.9]).double(),
                            shs=shs, sh_degree=1)
        out = torch.cat([r["image"], r["depth"], r["normal"], r["alpha"]], 0)
        return (out * wts).sum()

    assert torch.autograd.gradcheck(f, (xyz, scales, rot, op, shs), eps=1e-6, atol=1e-5, rtol=1e-3, nondet_tol=0)


def test_knn_oracle_basic():
    ref = torch.tensor([[0., 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [1, 0, 0]])
    q = torch.tensor([[0.9, 0, 0]])
    d, i = oknn.knn(ref, q, 4)
    assert i.tolist() == [[1, 4, 0, 2]]          # tie between 1 and 4 -> lower index first
    assert torch.allclose(d[0], torch.tensor([0.1, 0.1, 0.9, math.sqrt(0.81 + 4)]))
    pts = torch.tensor([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [5, 5, 5]])
    assert torch.allclose(oknn.dist3nn(pts)[0], torch.tensor(1.0))


def test_ewa_covariance_against_the_true_projection():
    """The screen-space covariance the oracle derives in closed form (J W Sigma W^T J^T + 0.3 I) against two derivations
    that never touch that formula: (a) the numerical Jacobian of the exact world -> pixel mapping given by the
    camera matrices, (b) the sample covariance of points drawn from the 3-D Gaussian and projected exactly."""
    import numpy as np
    W, H = 96, 64
    cam = _cam(W, H, view=3)
    g = torch.Generator().manual_seed(0)
    n = 12
    xyz = (torch.rand(n, 3, generator=g, dtype=torch.float64) - 0.5) * 0.6
    scales = torch.rand(n, 3, generator=g, dtype=torch.float64) * 0.02 + 0.004
    q = torch.nn.functional.normalize(torch.randn(n, 4, generator=g, dtype=torch.float64))
    pre = orast.preprocess(xyz, scales, q, torch.ones(n, dtype=torch.float64), cam.world_view_transform.double(),
                           cam.full_proj_transform.double(), cam.camera_center.double(), cam.tanfovx, cam.tanfovy, W, H,
                           colors_precomp=torch.zeros(n, 3, dtype=torch.float64))
    P = cam.full_proj_transform.double().numpy()

    def pixel(p):                                     # exact mapping: row vector times full projection, NDC -> pixel
        h = np.concatenate([p, np.ones_like(p[..., :1])], axis=-1) @ P
        ndc = h[..., :2] / h[..., 3:4]
        return ((ndc + 1.0) * np.array([W, H]) - 1.0) * 0.5

    w, x, y, z = q.numpy().T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                  np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                  np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], 1)
    S3 = R @ np.stack([np.diag(s ** 2) for s in scales.numpy()]) @ R.transpose(0, 2, 1)
    cov = pre["cov2d"].numpy()
    rng = np.random.default_rng(0)
    for i in range(n):
        p = xyz[i].numpy()
        assert np.abs(pixel(p) - pre["xy"][i].numpy()).max() < 1e-5        # the oracle divides by (w + 1e-7)
        eps = 1e-6
        J = np.stack([(pixel(p + eps * e) - pixel(p - eps * e)) / (2 * eps) for e in np.eye(3)], axis=1)   # [2,3]
        C = J @ S3[i] @ J.T
        got = np.array([[cov[i, 0] - orast.DILATION, cov[i, 1]], [cov[i, 1], cov[i, 2] - orast.DILATION]])
        assert np.abs(got - C).max() <= 1e-6 * np.abs(C).max(), i
        samples = rng.multivariate_normal(p, S3[i], size=200_000)
        Cs = np.cov(pixel(samples).T)
        assert np.abs(Cs - C).max() <= 0.03 * np.abs(C).max(), i              # Monte Carlo + second-order terms
    # conic = inverse of the dilated covariance; radius = ceil(3 sigma_max)
    for i in range(n):
        A = np.array([[cov[i, 0], cov[i, 1]], [cov[i, 1], cov[i, 2]]])
        inv = np.linalg.inv(A)
        assert np.allclose([inv[0, 0], inv[0, 1], inv[1, 1]], pre["conic"][i].numpy(), rtol=1e-9)
        # published rule: lambda_max = mid + sqrt(max(0.1, mid^2 - det)) -- the floor widens nearly isotropic footprints
        mid, det = 0.5 * (A[0, 0] + A[1, 1]), np.linalg.det(A)
        lam = mid + np.sqrt(max(orast.LAMBDA_FLOOR, mid * mid - det))
        assert lam >= np.linalg.eigvalsh(A).max() - 1e-12
        assert int(pre["radii"][i]) == int(np.ceil(orast.RADIUS_SIGMAS * np.sqrt(lam)))


def test_tile_pipeline_against_a_per_pixel_scalar_renderer():
    """Binning (tile rectangles -> keys -> stable sort -> ranges) + vectorised per-tile blending of the oracle against a
    second implementation with no tile lists at all: a scalar loop per pixel over ALL Gaussians in (depth, index) order,
    each admitted when the pixel's tile lies inside its rectangle, blended front to back with the four published
    thresholds.  Non-square image with partial tiles; overlapping, partly opaque Gaussians so that early termination,
    the alpha < 1/255 skip and the 0.99 clamp all fire."""
    import numpy as np
    W, H = 50, 37
    cam = _cam(W, H, view=2)
    g = torch.Generator().manual_seed(4)
    n = 90
    xyz = (torch.rand(n, 3, generator=g) - 0.5) * 0.7
    scales = torch.rand(n, 3, generator=g) * 0.12 + 0.01
    rot = torch.nn.functional.normalize(torch.randn(n, 4, generator=g))
    op = torch.rand(n, generator=g) * 0.98 + 0.02
    op[::5] = 1.0                                                  # opaque ones: clamp + fast saturation
    rgb = torch.rand(n, 3, generator=g)
    bg = (0.2, 0.5, 0.9)
    res = _raster(cam, xyz, scales, rot, op, rgb, W, H, bg=bg)
    pre = res["pre"]
    xy, con, depth = pre["xy"].numpy(), pre["conic"].numpy(), pre["depth"].numpy()
    rect, vis = pre["rect"].numpy(), (pre["tiles_touched"] > 0).numpy()
    order = sorted(range(n), key=lambda i: (np.float32(depth[i]).view(np.int32), i))   # key = float bits of the depth
    img = np.zeros((3, H, W)); dep = np.zeros((H, W)); alp = np.zeros((H, W)); ncon = np.zeros((H, W), dtype=np.int64)
    stats = dict(skipped=0, clamped=0, stopped=0)
    for py in range(H):
        for px in range(W):
            tx, ty = px // 16, py // 16
            T, seen, last = 1.0, 0, 0
            c = np.zeros(3); d = 0.0
            for i in order:
                if not vis[i] or not (rect[i, 0] <= tx < rect[i, 2] and rect[i, 1] <= ty < rect[i, 3]):
                    continue
                seen += 1
                dx, dy = xy[i, 0] - px, xy[i, 1] - py
                power = -0.5 * (con[i, 0] * dx * dx + con[i, 2] * dy * dy) - con[i, 1] * dx * dy
                if power > 0:
                    continue
                a = float(op[i]) * np.exp(power)
                if a > 0.99:
                    a = 0.99; stats["clamped"] += 1
                if a < 1.0 / 255.0:
                    stats["skipped"] += 1
                    continue
                if T * (1 - a) < 1e-4:
                    stats["stopped"] += 1
                    break
                c += rgb[i].numpy() * a * T; d += depth[i] * a * T
                T *= 1 - a
                last = seen
            img[:, py, px] = c + T * np.array(bg); dep[py, px] = d; alp[py, px] = 1 - T; ncon[py, px] = last
    assert min(stats.values()) > 0, stats                          # every threshold was exercised
    assert np.abs(res["image"].numpy() - img).max() <= 2e-5
    assert np.abs(res["depth"][0].numpy() - dep).max() <= 5e-5
    assert np.abs(res["alpha"][0].numpy() - alp).max() <= 2e-5
    assert (res["n_contrib"].numpy() == ncon).mean() > 0.999        # position of the last contributor in the tile list
