"""Shared CUDA-vs-oracle comparison helpers (used by the -m gpu tests and __graft_entry__.smoke)."""
import math

import torch

import dimo_b200
from dimo_b200 import raster as draster, synthetic
from dimo_b200.camera import orbit_minicam
from oracle import raster as oraster, camera as ocamera, deform as odeform

PIX_TOL = 1e-4      # north_star: within 1e-4 rel on pixel values and gradients
GRAD_TOL = 1e-4


def rel_err(a, ref):
    """max |a-ref| relative to the tensor's scale max|ref| (per-element relative error is meaningless for
    the near-zero entries every gradient tensor has)."""
    a = a.detach().double().cpu()
    ref = ref.detach().double().cpu()
    scale = max(ref.abs().max().item(), 1e-30)
    return (a - ref).abs().max().item() / scale


def outlier_frac(a, ref, tol):
    a = a.detach().double().cpu(); ref = ref.detach().double().cpu()
    scale = max(ref.abs().max().item(), 1e-30)
    return ((a - ref).abs() > tol * scale).double().mean().item()


def scene_inputs(N, seed=0, sh_coeffs=1, scale_boost=0.0):
    sc = synthetic.make_scene(N, n_ctrl=min(512, N), seed=seed, sh_coeffs=sh_coeffs)
    xyz = sc["_xyz"]
    scales = torch.exp(sc["_scaling"] + scale_boost)
    rot = torch.nn.functional.normalize(sc["_rotation"])
    op = torch.sigmoid(sc["_opacity"])
    shs = torch.cat([sc["_features_dc"], sc["_features_rest"]], dim=1)
    return xyz, scales, rot, op, shs


def loss_weights(H, W, seed=1):
    g = torch.Generator().manual_seed(seed)
    return dict(c=torch.rand(3, H, W, generator=g), d=torch.rand(1, H, W, generator=g) * 0.3,
                n=torch.rand(3, H, W, generator=g) - 0.5, a=torch.rand(1, H, W, generator=g))


def weighted_loss(img, depth, normal, alpha, w, use_dn=True):
    """use_dn=False: depth and normal stay out of the loss -> the CUDA backward runs its 9-field (DN = false) variant"""
    l = (img * w["c"]).sum() + (alpha * w["a"]).sum()
    if use_dn:
        l = l + (depth * w["d"]).sum() + (normal * w["n"]).sum()
    return l


def run_raster_pair(N, W, H, view=1, nviews=8, sh_degree=0, seed=0, scale_boost=0.0, bg=(1.0, 1.0, 1.0),
                    device="cuda", use_dn=True, backward=True):
    """Runs the oracle (CPU fp32 + autograd) and the CUDA path on identical inputs; returns both result dicts."""
    K = (sh_degree + 1) ** 2
    xyz, scales, rot, op, shs = scene_inputs(N, seed, K, scale_boost)
    bg_t = torch.tensor(bg, dtype=torch.float32)
    w = loss_weights(H, W)

    # ---- oracle ----
    ocam = ocamera.orbit_cam(view, nviews, W, H)
    leaves = [t.clone().requires_grad_(True) for t in (xyz, scales, rot, op, shs)]
    m2d = torch.zeros(N, 3, requires_grad=True)
    o = oraster.rasterize(leaves[0], leaves[1], leaves[2], leaves[3], ocam.world_view_transform,
                          ocam.full_proj_transform, ocam.camera_center, ocam.tanfovx, ocam.tanfovy, W, H, bg_t,
                          shs=leaves[4], sh_degree=sh_degree, means2D=m2d)
    o["grads"] = {}
    if backward:
        weighted_loss(o["image"], o["depth"], o["normal"], o["alpha"], w, use_dn).backward()
        o["grads"] = dict(means3D=leaves[0].grad, scales=leaves[1].grad, rotations=leaves[2].grad,
                          opacities=leaves[3].grad, shs=leaves[4].grad, means2D=m2d.grad)

    # ---- CUDA ----
    cam = orbit_minicam(view, nviews, W, H, device=device)
    assert torch.equal(cam.world_view_transform.cpu(), ocam.world_view_transform)
    assert torch.equal(cam.full_proj_transform.cpu(), ocam.full_proj_transform)
    cams = draster.pack_cameras(cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                                math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), bg_t.to(device))
    cl = [t.clone().to(device).requires_grad_(True) for t in (xyz, scales, rot, op, shs)]
    cm2d = torch.zeros(N, 3, device=device, requires_grad=True)
    state = []
    color, depth, normal, alpha, radii = draster.rasterize_batch(
        cams, cl[0], cl[1], cl[2], cl[3], W, H, shs=cl[4], sh_degree=sh_degree, means2D=cm2d, state_out=state)
    wd = {k: v.to(device) for k, v in w.items()}
    if backward:
        weighted_loss(color[0], depth[0], normal[0], alpha[0], wd, use_dn).backward()
    torch.cuda.synchronize()
    st = state[0]
    # the library sorts 32-bit tile ids (emitted front-to-back); rebuild the reference-shaped 64-bit
    # (tile << 32 | depth bits) keys from the sorted tile ids and the sorted splats' depths for comparison
    vals = st.record_ids(st.R)
    dbits = st.splats[vals, 10].contiguous().view(torch.int32).long() & 0xFFFFFFFF
    keys64 = (st.tile_keys(st.R) << 32) | dbits
    c = dict(image=color[0], depth=depth[0], normal=normal[0], alpha=alpha[0], radii=radii[0],
             tiles_touched=st.tiles_touched, keys=keys64, ids=vals,
             ranges=st.ranges, n_contrib=st.n_contrib[0], final_T=st.final_T[0], R=st.R,
             grads=dict(means3D=cl[0].grad, scales=cl[1].grad, rotations=cl[2].grad, opacities=cl[3].grad,
                        shs=cl[4].grad, means2D=cm2d.grad) if backward else {})
    return o, c


def compare_raster(o, c, verbose=True):
    """Returns (int_mismatches: dict name->count, float_errs: dict name->rel_err)."""
    ints = {}
    ints["radii"] = int((o["radii"] != c["radii"].cpu()).sum())
    ints["tiles_touched"] = int((o["tiles_touched"] != c["tiles_touched"].cpu()).sum())
    ints["R"] = abs(int(o["keys"].numel()) - int(c["R"]))
    if ints["R"] == 0:
        ints["keys"] = int((o["keys"] != c["keys"].cpu()).sum())
        ints["ids"] = int((o["ids"] != c["ids"].cpu().long()).sum())
        # empty tiles: the kernel leaves (0,0), the oracle's cumsum gives (k,k) -- both mean "no instances"
        orng, crng = o["ranges"], c["ranges"].cpu().long()
        o_empty, c_empty = orng[:, 1] == orng[:, 0], crng[:, 1] == crng[:, 0]
        ints["ranges"] = int((o_empty != c_empty).sum()) + int((orng[~o_empty] != crng[~o_empty]).sum())
    flo = {k: rel_err(c[k], o[k]) for k in ("image", "depth", "normal", "alpha", "final_T")}
    flo["n_contrib_mismatch_frac"] = (o["n_contrib"] != c["n_contrib"].cpu()).double().mean().item()
    gr = {k: rel_err(c["grads"][k], o["grads"][k]) for k in o["grads"]}
    if verbose:
        print("ints", ints)
        print("pix ", {k: f"{v:.2e}" for k, v in flo.items()})
        print("grad", {k: f"{v:.2e}" for k, v in gr.items()})
    return ints, flo, gr


def l2_err(a, ref):
    """||a - ref||_2 / ||ref||_2 over the whole tensor (second metric next to rel_err: insensitive to one large entry
    setting the scale, sensitive to many small entries being wrong)."""
    a = a.detach().double().cpu(); ref = ref.detach().double().cpu()
    return ((a - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def entry_outlier_frac(a, ref, rtol=1e-4, atol_scale=1e-6):
    """fraction of entries with |a - ref| > rtol * |ref| + atol_scale * max|ref| (per-entry relative error with an
    absolute floor for the near-zero entries)"""
    a = a.detach().double().cpu(); ref = ref.detach().double().cpu()
    scale = max(ref.abs().max().item(), 1e-30)
    return ((a - ref).abs() > rtol * ref.abs() + atol_scale * scale).double().mean().item()


def cuda_relu_masks(capture):
    """dimo_b200.deform.DEBUG_CAPTURE entry -> the ten activation patterns [R,256] (bool) in the oracle's ReLU order:
    deformnet.0..7, pts_layers.0, rot_layers.0"""
    cat, hp, hr, *acts = capture
    E = cat.shape[1] - odeform.HIDDEN
    trunk = acts[:odeform.SKIP_AFTER] + [cat[:, E:]] + acts[odeform.SKIP_AFTER:]
    return [(t > 0).cpu() for t in trunk + [hp, hr]]


def run_step_pair(N=1500, M=32, W=64, H=64, device="cuda", regularisers=False, force_masks=True):
    """One full deform -> raster -> loss step (4 frames: 2 motions x 1 view x 2 times) on the CUDA fast path and on the
    oracle, same seeded inputs.  Returns (loss_cuda, loss_oracle, grads_cuda, grads_oracle, mask_stats): gradients of
    EVERY parameter on the path (Gaussian attributes, control points, latents, all 24 TimeNet tensors).

    force_masks: the oracle's TimeNet is evaluated on the activation pattern the CUDA forward chose (exported ReLU
    masks), so a pre-activation that the two evaluations round to different sides of zero does not turn into a
    whole-row gradient difference; mask_stats reports how many of the R*2560 signs differ between the oracle's own
    pattern and the CUDA one and how far from zero (relative to the layer's scale) the oracle's pre-activation is at
    those places -- a genuine kink flip has |z| ~ 1e-6 * scale, anything larger is a bug."""
    from dimo_b200 import trainstep, deform as ddeform
    from dimo_b200.renderer import Renderer
    from oracle import loss as oloss, knn as oknn
    sc = synthetic.make_scene(N, n_ctrl=M, n_motions=2, seed=11)
    params = odeform.timenet_init(32, seed=5, final_scale=0.05)
    frames = [(0, 1, 0.25), (0, 1, 0.75), (1, 1, 0.25), (1, 1, 0.75)]      # (motion, view, t), motion-major
    g = torch.Generator().manual_seed(2)
    gt = torch.rand(4, 3, H, W, generator=g); mk = torch.rand(4, 1, H, W, generator=g)
    lw = trainstep.StepLossWeights

    # ---- CUDA fast path (eager TrainStep, no optimizer update) ----
    r = Renderer(sh_degree=0, num_latent_code=2, add_normal=True, device=device)
    r.gaussians.load_state(sc)
    tn = r.gaussians._timenet
    with torch.no_grad():
        for p_pair, (Wt, b) in zip(zip(tn.flat_params()[0::2], tn.flat_params()[1::2]), params):
            p_pair[0].copy_(Wt); p_pair[1].copy_(b)
    ts = trainstep.TrainStep(r, lr=0.0)
    cams = [orbit_minicam(v, 4, W, H, device=device) for (_, v, _) in frames]
    prep = r.prepare_step(cams, [t for (_, _, t) in frames], [m for (m, _, _) in frames])
    ts.g.find_knn(4)
    ddeform.DEBUG_CAPTURE = []
    try:
        out = r.render_batch(prepared=prep, stage="s2", clamp=False)
        capture = ddeform.DEBUG_CAPTURE[0]
    finally:
        ddeform.DEBUG_CAPTURE = None
    lc = trainstep.step_loss(out["image_raw"], out["alpha"], gt.to(device), mk.to(device), 2)
    if regularisers:
        from dimo_b200 import loss as dloss
        lc = lc + dloss.smoothness_losses(out["image_raw"], out["depth"], out["normal"], groups=2,
                                          lambda_smooth=lw.lambda_smooth, lambda_bilateral=lw.lambda_bilateral,
                                          clamp01=True)
    lc.backward()
    torch.cuda.synchronize()
    G = r.gaussians
    gc = {"xyz": G._xyz.grad, "features_dc": G._features_dc.grad, "opacity": G._opacity.grad,
          "scaling": G._scaling.grad, "rotation": G._rotation.grad, "c_xyz": G._c_xyz.grad,
          "c_radius": G._c_radius.grad, "latents": G._latent_codes.grad}
    for li, (pw, pb) in enumerate(zip(tn.flat_params()[0::2], tn.flat_params()[1::2])):
        gc[f"W{li}"] = pw.grad; gc[f"b{li}"] = pb.grad
    gc = {k: v.detach().clone() for k, v in gc.items()}
    masks = cuda_relu_masks(capture)                       # rows ordered [pair][control point]
    pair_of_frame = prep["pair_of_frame"]

    # ---- oracle ----
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    op = [(Wt.clone().requires_grad_(True), b.clone().requires_grad_(True)) for Wt, b in params]
    dist, idx = oknn.knn(sc["_c_xyz"], sc["_xyz"], 4)
    imgs, alphas, depths, normals = [], [], [], []
    n_signs = n_flip = 0
    worst_flip = 0.0
    for f, (m, v, t) in enumerate(frames):
        cam = ocamera.orbit_cam(v, 4, W, H)
        rows = slice(pair_of_frame[f] * M, (pair_of_frame[f] + 1) * M)
        own = []
        dxyz, dquat = odeform.timenet_forward(op, leaves["_c_xyz"], t, leaves["_latent_codes"][m],
                                              masks_in=[mm[rows] for mm in masks] if force_masks else None,
                                              masks_out=own)
        for (z, pos), mm in zip(own, masks):
            diff = pos != mm[rows]
            n_signs += pos.numel(); n_flip += int(diff.sum())
            if bool(diff.any()):
                worst_flip = max(worst_flip, float(z[diff].abs().max() / z.abs().max()))
        means, rots = odeform.lbs_deform(leaves["_xyz"], leaves["_rotation"], leaves["_c_xyz"],
                                         torch.exp(leaves["_c_radius"]), dxyz, dquat, idx, dist)
        o = oraster.rasterize(means, torch.exp(leaves["_scaling"]), rots, torch.sigmoid(leaves["_opacity"]),
                              cam.world_view_transform, cam.full_proj_transform, cam.camera_center, cam.tanfovx,
                              cam.tanfovy, W, H, torch.ones(3), shs=torch.cat([leaves["_features_dc"], leaves["_features_rest"]], 1))
        imgs.append(o["image"].clamp(0, 1)); alphas.append(o["alpha"])
        depths.append(o["depth"]); normals.append(o["normal"])
    img = torch.stack(imgs); alp = torch.stack(alphas)
    dep = torch.stack(depths); nrm = torch.stack(normals)
    lo = 0
    for f in range(4):
        lo = lo + lw.lambda_mse * oloss.mse_loss(img[f], gt[f])
    for m in range(2):
        sl = slice(2 * m, 2 * m + 2)
        lo = lo + lw.lambda_ssim * (1 - oloss.ssim(img[sl], gt[sl])) + lw.lambda_mask * oloss.mse_loss(alp[sl], mk[sl])
        if regularisers:       # main_train_dimo.py:363-372
            lo = lo + lw.lambda_smooth * oloss.edge_aware_smoothness(dep[sl], img[sl]) + \
                lw.lambda_bilateral * oloss.bilateral_normal_smoothness(nrm[sl], img[sl])
    lo.backward()
    go = {"xyz": leaves["_xyz"].grad, "features_dc": leaves["_features_dc"].grad, "opacity": leaves["_opacity"].grad,
          "scaling": leaves["_scaling"].grad, "rotation": leaves["_rotation"].grad, "c_xyz": leaves["_c_xyz"].grad,
          "c_radius": leaves["_c_radius"].grad, "latents": leaves["_latent_codes"].grad}
    for li, (Wt, b) in enumerate(op):
        go[f"W{li}"] = Wt.grad; go[f"b{li}"] = b.grad
    stats = {"signs": n_signs, "flips": n_flip, "worst_flip_rel_z": worst_flip}
    return float(lc), float(lo), gc, go, stats


def smoke():
    o, c = run_raster_pair(800, 64, 64)
    ints, flo, gr = compare_raster(o, c)
    assert all(v == 0 for v in ints.values()), ints
    assert all(flo[k] < PIX_TOL for k in ("image", "depth", "normal", "alpha")), flo
    assert all(v < 5 * GRAD_TOL for v in gr.values()), gr
    lc, lo, gc, go, stats = run_step_pair()
    print(f"full step: loss cuda {lc:.6f} oracle {lo:.6f}; ReLU pattern {stats}")
    assert abs(lc - lo) <= 1e-4 * abs(lo), (lc, lo)
