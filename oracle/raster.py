"""Oracle: tile-based differentiable 3D-Gaussian rasterisation (colour, depth, normal, alpha).

TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED.  The reference calls the third-party CUDA extensions ``diff_gauss``
(slothfulxtx/diff-gaussian-rasterization @ 726449a8, live path,
renderer/latent_gs_renderer.py:1133-1147,1256-1266) and ``diff_gaussian_rasterization``
(ashawkey @ 8829d14f, :1149-1163,1268-1277); neither source tree is present under
/root/reference and the reference holds no golden images.  This file restates the
published 3DGS tile-rasterisation algorithm (Kerbl et al. 2023, as extended with depth /
alpha / normal outputs by those forks).  Every constant is named below.

Integer-producing arithmetic (view transform, covariance, radius, pixel centre, tile
rectangle, depth key) is written as an explicit left-to-right sequence of separately
rounded fp32 operations; the CUDA preprocess translation unit is compiled with
-fmad=false and performs the same sequence, which is what makes radii / tiles_touched /
sort keys / tile ranges bit-exact between the two.

The per-pixel blend is vectorised per tile and differentiated by autograd; the clamp
alpha=min(0.99, o*G) uses a straight-through gradient as the published backward pass does
(it differentiates o*G and ignores the clamp).
"""
import math
import torch

from .sh import eval_sh_rgb

TILE = 16                    # BLOCK_X = BLOCK_Y = 16
NEAR_CULL_Z = 0.2            # in_frustum: p_view.z <= 0.2 -> culled
FOV_CLAMP = 1.3              # cov2D: t.xy/t.z clamped to +-1.3*tanfov
DILATION = 0.3               # low-pass added to the cov2D diagonal (pixels^2)
LAMBDA_FLOOR = 0.1           # max(0.1, mid^2 - det) inside the eigenvalue sqrt
RADIUS_SIGMAS = 3.0          # radius = ceil(3 * sqrt(lambda_max))
W_EPS = 1e-7                 # p_w = 1 / (p_hom.w + 1e-7)
ALPHA_MAX = 0.99             # alpha = min(0.99, opacity * G)
ALPHA_MIN = 1.0 / 255.0      # skip if alpha < 1/255
T_MIN = 1e-4                 # stop before T would fall below 1e-4
CLAMP_STRAIGHT_THROUGH = True


def _f32(x):
    return torch.tensor(x, dtype=torch.float32).item()


def _sqrt_rn(x):
    """Correctly rounded fp32 square root.  torch.sqrt on CPU (SLEEF) is off by 1 ulp for ~0.6 % of inputs, the
    CUDA sqrtf (default -prec-sqrt=true) is IEEE; sqrt in fp64 followed by rounding to fp32 is exact-rounded
    (53 >= 2*24+2 bits), so this is the same function the kernel computes."""
    if x.dtype == torch.float32:
        return x.double().sqrt().float()
    return torch.sqrt(x)


def _dot3(a0, a1, a2, b0, b1, b2):
    return (a0 * b0 + a1 * b1) + a2 * b2


def preprocess(means3D, scales, rotations, opacities, view, proj, campos, tanfovx, tanfovy,
               W, H, scale_modifier=1.0, shs=None, sh_degree=0, colors_precomp=None,
               means2D=None):
    """Per-Gaussian projection.  All inputs fp32 (or fp64 for gradcheck) CPU tensors.

    view / proj: the [4,4] row-vector-convention matrices the reference passes
    (world_view_transform, full_proj_transform: renderer/latent_gs_renderer.py:960-969).
    Returns a dict of per-Gaussian tensors; invisible Gaussians have radius 0.
    """
    dt = means3D.dtype
    px, py, pz = means3D[:, 0], means3D[:, 1], means3D[:, 2]
    V = view.to(dt)
    P = proj.to(dt)
    tvx = ((px * V[0, 0] + py * V[1, 0]) + pz * V[2, 0]) + V[3, 0]
    tvy = ((px * V[0, 1] + py * V[1, 1]) + pz * V[2, 1]) + V[3, 1]
    tvz = ((px * V[0, 2] + py * V[1, 2]) + pz * V[2, 2]) + V[3, 2]
    hx = ((px * P[0, 0] + py * P[1, 0]) + pz * P[2, 0]) + P[3, 0]
    hy = ((px * P[0, 1] + py * P[1, 1]) + pz * P[2, 1]) + P[3, 1]
    hw = ((px * P[0, 3] + py * P[1, 3]) + pz * P[2, 3]) + P[3, 3]
    p_w = 1.0 / (hw + W_EPS)
    ndc_x = hx * p_w
    ndc_y = hy * p_w
    if means2D is not None:                    # gradient sink (viewspace_points), NDC units
        ndc_x = ndc_x + means2D[:, 0]
        ndc_y = ndc_y + means2D[:, 1]

    # 3D covariance  Sigma = (R S)(R S)^T ; quaternion used as given (not re-normalised)
    s = scales * scale_modifier
    r, x, y, z = rotations[:, 0], rotations[:, 1], rotations[:, 2], rotations[:, 3]
    R = [[1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)],
         [2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)],
         [2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]]
    L = [[R[i][j] * s[:, j] for j in range(3)] for i in range(3)]
    S = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(i, 3):
            S[i][j] = _dot3(L[i][0], L[i][1], L[i][2], L[j][0], L[j][1], L[j][2])
            S[j][i] = S[i][j]

    # 2D covariance  (EWA): cov = J W Sigma W^T J^T + 0.3 I
    if dt == torch.float32:          # scalars rounded exactly as the fp32 host/device code does
        limx = _f32(_f32(FOV_CLAMP) * _f32(tanfovx))
        limy = _f32(_f32(FOV_CLAMP) * _f32(tanfovy))
        fx = _f32(float(W) / _f32(2.0 * _f32(tanfovx)))
        fy = _f32(float(H) / _f32(2.0 * _f32(tanfovy)))
    else:
        limx = FOV_CLAMP * tanfovx
        limy = FOV_CLAMP * tanfovy
        fx = W / (2.0 * tanfovx)
        fy = H / (2.0 * tanfovy)
    txtz = tvx / tvz
    tytz = tvy / tvz
    tx = torch.clamp(txtz, -limx, limx) * tvz
    ty = torch.clamp(tytz, -limy, limy) * tvz
    # NB: `python_scalar / tensor` is evaluated by torch as tensor.reciprocal() * scalar (two roundings); the kernel
    # performs ONE IEEE division, so the numerator is materialised as a tensor first.  (Found by the 1024x1024 test:
    # 1 radius in 30 000 differed by one pixel through the cancellation in mid^2 - det.)
    J00 = torch.full_like(tvz, fx) / tvz
    J02 = -(fx * tx) / (tvz * tvz)
    J11 = torch.full_like(tvz, fy) / tvz
    J12 = -(fy * ty) / (tvz * tvz)
    M0 = [J00 * V[k, 0] + J02 * V[k, 2] for k in range(3)]
    M1 = [J11 * V[k, 1] + J12 * V[k, 2] for k in range(3)]
    v0 = [_dot3(S[i][0], S[i][1], S[i][2], M0[0], M0[1], M0[2]) for i in range(3)]
    v1 = [_dot3(S[i][0], S[i][1], S[i][2], M1[0], M1[1], M1[2]) for i in range(3)]
    ca = _dot3(M0[0], M0[1], M0[2], v0[0], v0[1], v0[2]) + DILATION
    cb = _dot3(M0[0], M0[1], M0[2], v1[0], v1[1], v1[2])
    cc = _dot3(M1[0], M1[1], M1[2], v1[0], v1[1], v1[2]) + DILATION
    det = ca * cc - cb * cb
    det_inv = 1.0 / det
    conic_a = cc * det_inv
    conic_b = -cb * det_inv
    conic_c = ca * det_inv
    mid = 0.5 * (ca + cc)
    root = _sqrt_rn(torch.clamp_min(mid * mid - det, LAMBDA_FLOOR))
    lam = torch.maximum(mid + root, mid - root)
    radius_f = torch.ceil(RADIUS_SIGMAS * _sqrt_rn(lam))
    pix_x = ((ndc_x + 1.0) * W - 1.0) * 0.5
    pix_y = ((ndc_y + 1.0) * H - 1.0) * 0.5

    gx = (W + TILE - 1) // TILE
    gy = (H + TILE - 1) // TILE
    with torch.no_grad():
        finite = torch.isfinite(pix_x) & torch.isfinite(pix_y) & torch.isfinite(radius_f)
        rf = torch.where(finite, radius_f, torch.zeros_like(radius_f))
        fxp = torch.where(finite, pix_x, torch.zeros_like(pix_x))
        fyp = torch.where(finite, pix_y, torch.zeros_like(pix_y))

        def tile_coord(v, g):
            return torch.clamp(torch.trunc(v), 0.0, float(g)).to(torch.int32)
        rminx = tile_coord((fxp - rf) / TILE, gx)
        rminy = tile_coord((fyp - rf) / TILE, gy)
        rmaxx = tile_coord((fxp + rf + (TILE - 1)) / TILE, gx)
        rmaxy = tile_coord((fyp + rf + (TILE - 1)) / TILE, gy)
        tiles = (rmaxx - rminx) * (rmaxy - rminy)
        visible = (tvz > NEAR_CULL_Z) & (det != 0) & finite & (tiles > 0)
        radii = torch.where(visible, rf.to(torch.int32), torch.zeros_like(rminx))
        tiles = torch.where(visible, tiles, torch.zeros_like(tiles))

    # colour
    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        rgb = eval_sh_rgb(sh_degree, shs, means3D, campos.to(dt))

    # view-space normal: shortest-axis direction, oriented towards the camera, rotated by the
    # view matrix (in-tree definition: renderer/latent_gs_renderer.py:387-401 get_smallest_axis /
    # get_normal, and the commented block :1244-1247  local_normal = global_normal @ W[:3,:3]).
    with torch.no_grad():                      # first minimum on ties
        kmin = torch.where((s[:, 0] <= s[:, 1]) & (s[:, 0] <= s[:, 2]), 0,
                           torch.where(s[:, 1] <= s[:, 2], 1, 2))
    Rcols = torch.stack([torch.stack([R[0][k], R[1][k], R[2][k]], dim=-1) for k in range(3)], dim=1)  # N,k,3
    n_w = Rcols[torch.arange(Rcols.shape[0]), kmin]
    cp = campos.to(dt)
    dotp = _dot3(n_w[:, 0], n_w[:, 1], n_w[:, 2], cp[0] - px, cp[1] - py, cp[2] - pz)
    sign = torch.where(dotp < 0, -torch.ones_like(dotp), torch.ones_like(dotp)).detach()
    n_w = n_w * sign[:, None]
    normal = torch.stack([_dot3(n_w[:, 0], n_w[:, 1], n_w[:, 2], V[0, j], V[1, j], V[2, j]) for j in range(3)], dim=-1)

    return dict(xy=torch.stack([pix_x, pix_y], dim=-1), depth=tvz,
                conic=torch.stack([conic_a, conic_b, conic_c], dim=-1),
                opacity=opacities.reshape(-1), rgb=rgb, normal=normal,
                radii=radii, tiles_touched=tiles,
                rect=torch.stack([rminx, rminy, rmaxx, rmaxy], dim=-1),
                cov2d=torch.stack([ca, cb, cc], dim=-1), grid=(gx, gy))


def bin_tiles(pre, frame=0):
    """duplicateWithKeys + stable radix sort + identifyTileRanges.

    key = (tile_id << 32) | float_bits(depth); ties keep ascending Gaussian index (a stable
    sort over keys emitted in Gaussian order).  Returns keys [R] int64, ids [R] int64,
    ranges [tiles,2] int64.
    """
    gx, gy = pre["grid"]
    rect = pre["rect"]
    tiles = pre["tiles_touched"]
    vis = torch.nonzero(tiles > 0).flatten()
    depth_bits = pre["depth"].detach().to(torch.float32).view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys, ids = [], []
    for i in vis.tolist():
        x0, y0, x1, y1 = rect[i].tolist()
        ys = torch.arange(y0, y1, dtype=torch.int64)
        xs = torch.arange(x0, x1, dtype=torch.int64)
        t = (ys[:, None] * gx + xs[None, :]).flatten() + frame * gx * gy
        keys.append((t << 32) | depth_bits[i])
        ids.append(torch.full_like(t, i))
    if keys:
        keys = torch.cat(keys); ids = torch.cat(ids)
        order = torch.sort(keys, stable=True).indices
        keys = keys[order]; ids = ids[order]
    else:
        keys = torch.zeros(0, dtype=torch.int64); ids = torch.zeros(0, dtype=torch.int64)
    ntiles = gx * gy
    tile_of = (keys >> 32) - frame * ntiles
    counts = torch.bincount(tile_of, minlength=ntiles) if len(keys) else torch.zeros(ntiles, dtype=torch.int64)
    ends = torch.cumsum(counts, 0)
    starts = ends - counts
    return keys, ids, torch.stack([starts, ends], dim=-1)


def blend(pre, ids, ranges, W, H, bg, extra=None):
    """Front-to-back alpha blending per 16x16 tile.  Returns image [3,H,W], depth [1,H,W],
    normal [3,H,W], alpha [1,H,W], n_contrib [H,W] int32, final_T [H,W] (+ extra [E,H,W])."""
    dt = pre["xy"].dtype
    gx, gy = pre["grid"]
    feats = [pre["rgb"], pre["depth"][:, None], pre["normal"]]
    if extra is not None:
        feats.append(extra)
    feat = torch.cat(feats, dim=-1)                      # N, C
    C = feat.shape[1]
    out = torch.zeros(C, gy * TILE, gx * TILE, dtype=dt)
    Tfin = torch.ones(gy * TILE, gx * TILE, dtype=dt)
    ncontrib = torch.zeros(gy * TILE, gx * TILE, dtype=torch.int32)
    out_tiles, T_tiles = {}, {}
    jj, ii = torch.meshgrid(torch.arange(TILE), torch.arange(TILE), indexing="ij")
    for ty in range(gy):
        for tx in range(gx):
            t = ty * gx + tx
            s0, s1 = ranges[t].tolist()
            if s1 <= s0:
                continue
            g = ids[s0:s1]
            pxs = (tx * TILE + ii).flatten().to(dt)
            pys = (ty * TILE + jj).flatten().to(dt)
            inside = ((tx * TILE + ii).flatten() < W) & ((ty * TILE + jj).flatten() < H)
            xy = pre["xy"][g]
            con = pre["conic"][g]
            op = pre["opacity"][g]
            dx = xy[None, :, 0] - pxs[:, None]
            dy = xy[None, :, 1] - pys[:, None]
            power = -0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) - con[None, :, 1] * dx * dy
            G = torch.exp(power)
            a_raw = op[None, :] * G
            a_cl = torch.clamp(a_raw, max=ALPHA_MAX)
            alpha = a_raw + (a_cl - a_raw).detach() if CLAMP_STRAIGHT_THROUGH else a_cl
            with torch.no_grad():
                skip = (power > 0) | (alpha < ALPHA_MIN) | (~inside[:, None])
                a_ns = torch.where(skip, torch.zeros_like(alpha), alpha)
                T_after = torch.cumprod(1 - a_ns, dim=1)
                contrib = (~skip) & (T_after >= T_MIN)
                idx = torch.arange(1, g.shape[0] + 1, dtype=torch.int32)[None, :].expand_as(contrib)
                last = torch.where(contrib, idx, torch.zeros_like(idx)).max(dim=1).values
            a_eff = torch.where(contrib, alpha, torch.zeros_like(alpha))
            T_incl = torch.cumprod(1 - a_eff, dim=1)
            T_before = torch.cat([torch.ones_like(T_incl[:, :1]), T_incl[:, :-1]], dim=1)
            wgt = a_eff * T_before                                 # 256, L
            acc = wgt @ feat[g]                                    # 256, C
            Tf = T_incl[:, -1]
            ys = slice(ty * TILE, (ty + 1) * TILE)
            xs = slice(tx * TILE, (tx + 1) * TILE)
            out_tiles[(ty, tx)] = acc.t().reshape(C, TILE, TILE)
            T_tiles[(ty, tx)] = Tf.reshape(TILE, TILE)
            ncontrib[ys, xs] = last.reshape(TILE, TILE)
    # assemble with autograd-friendly cat
    rows = []
    Trows = []
    for ty in range(gy):
        row = [out_tiles.get((ty, tx), torch.zeros(C, TILE, TILE, dtype=dt)) for tx in range(gx)]
        Trow = [T_tiles.get((ty, tx), torch.ones(TILE, TILE, dtype=dt)) for tx in range(gx)]
        rows.append(torch.cat(row, dim=2))
        Trows.append(torch.cat(Trow, dim=1))
    out = torch.cat(rows, dim=1)[:, :H, :W]
    Tfin = torch.cat(Trows, dim=0)[:H, :W]
    image = out[0:3] + Tfin[None] * bg.to(dt)[:, None, None]
    depth = out[3:4]
    normal = out[4:7]
    alpha = (1 - Tfin)[None]
    res = dict(image=image, depth=depth, normal=normal, alpha=alpha,
               n_contrib=ncontrib[:H, :W], final_T=Tfin)
    if extra is not None:
        res["extra"] = out[7:]
    return res


def rasterize(means3D, scales, rotations, opacities, view, proj, campos, tanfovx, tanfovy, W, H, bg,
              scale_modifier=1.0, shs=None, sh_degree=0, colors_precomp=None, means2D=None,
              extra_attrs=None):
    """Full forward; differentiable w.r.t. every floating tensor input via autograd.

    Pass 1 (no grad, all N) fixes the integers; pass 2 recomputes the visible subset with
    autograd so culled Gaussians (possibly inf/NaN intermediates) cannot poison gradients.
    """
    N = means3D.shape[0]
    with torch.no_grad():
        pre_all = preprocess(means3D, scales, rotations, opacities, view, proj, campos, tanfovx,
                             tanfovy, W, H, scale_modifier, shs, sh_degree, colors_precomp, means2D)
        keys, ids, ranges = bin_tiles(pre_all)
    vis = torch.nonzero(pre_all["tiles_touched"] > 0).flatten()
    remap = torch.full((N,), -1, dtype=torch.int64)
    remap[vis] = torch.arange(vis.numel())
    sel = lambda t: None if t is None else t[vis]
    pre = preprocess(means3D[vis], scales[vis], rotations[vis], opacities[vis], view, proj, campos,
                     tanfovx, tanfovy, W, H, scale_modifier, sel(shs), sh_degree,
                     sel(colors_precomp), sel(means2D))
    res = blend(pre, remap[ids], ranges, W, H, bg, sel(extra_attrs))
    res.update(radii=pre_all["radii"], tiles_touched=pre_all["tiles_touched"], keys=keys, ids=ids,
               ranges=ranges, pre=pre_all)
    return res
