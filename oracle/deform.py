"""Oracle: latent-conditioned deformation (positional encoding, TimeNet MLP, LBS skinning).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pure PyTorch, device/dtype agnostic
(fp32 for parity with the kernels, fp64 for gradcheck).  Each function cites the
reference lines it restates; pinned by tests/golden/deform_*.npz which were produced
by executing those very lines of the reference in the build container.
"""
import math
import torch
import torch.nn.functional as F

PTS_FREQS = 10      # renderer/latent_gs_renderer.py:187  (self.pts_ch = 10)
TIME_FREQS = 6      # renderer/latent_gs_renderer.py:188  (self.times_ch = 6)
HIDDEN = 256        # renderer/latent_gs_renderer.py:185  (W=256)
DEPTH = 8           # renderer/latent_gs_renderer.py:185  (D=8)
SKIP_AFTER = 4      # renderer/latent_gs_renderer.py:185  (skips=[4])
LBS_EPS = 1e-7      # renderer/latent_gs_renderer.py:1192


def posenc(x, num_freqs):
    """src/pos_enc.py:6-54 with include_input=False, log_sampling=True.

    Output order: for k in 0..L-1: [sin(2^k x) (all dims), cos(2^k x) (all dims)].
    The reference multiplies by a float32 tensor element ``freq`` (x * freq), so the
    product is rounded to the input dtype before sin/cos.
    """
    outs = []
    for k in range(num_freqs):
        f = float(2.0 ** k)
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, dim=-1)


def timenet_layer_shapes(latent_dim=32):
    """(out, in) of the 12 Linear layers in canonical order
    deformnet.0..7, pts_layers.0, pts_layers.2, rot_layers.0, rot_layers.2
    (renderer/latent_gs_renderer.py:192-197)."""
    in_ch = 3 * 2 * PTS_FREQS + 1 * 2 * TIME_FREQS + latent_dim
    shapes = [(HIDDEN, in_ch)]
    for i in range(DEPTH - 1):
        shapes.append((HIDDEN, HIDDEN + in_ch) if i == SKIP_AFTER else (HIDDEN, HIDDEN))
    shapes += [(HIDDEN, HIDDEN), (3, HIDDEN), (HIDDEN, HIDDEN), (4, HIDDEN)]
    return shapes


def timenet_init(latent_dim=32, seed=0, dtype=torch.float32, final_scale=None):
    """Weights in canonical order as a list of (W[out,in], b[out]).

    Mirrors renderer/latent_gs_renderer.py:166-203: xavier-uniform weights (gain 1),
    default nn.Linear bias init left in place (``initialize_weights`` re-inits the
    *weight* twice and never the bias), last pts layer zero, last rot layer W=0,
    b=[1,0,0,0].  ``final_scale``: instead of the zero/identity heads,
    xavier heads scaled by ``final_scale`` so every gradient path is live.
    """
    g = torch.Generator().manual_seed(seed)
    params = []
    shapes = timenet_layer_shapes(latent_dim)
    for li, (o, i) in enumerate(shapes):
        bound = math.sqrt(6.0 / (i + o))
        W = (torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * bound
        bb = 1.0 / math.sqrt(i)
        b = (torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * bb
        params.append([W.to(dtype), b.to(dtype)])
    if final_scale is None:
        params[9][0].zero_(); params[9][1].zero_()
        params[11][0].zero_(); params[11][1] = torch.tensor([1.0, 0, 0, 0], dtype=dtype)
    else:
        params[9][0] *= final_scale; params[9][1] *= final_scale
        params[11][0] *= final_scale; params[11][1] *= final_scale
    return [(W, b) for W, b in params]


def timenet_forward(params, pts, t, latent, return_kink_distance=False, masks_in=None, masks_out=None):
    """TimeNet.forward, renderer/latent_gs_renderer.py:205-235, on flat rows.

    pts [R,3], t [R,1] (or python float), latent [R,L] (or [L]) -> (dxyz [R,3], dquat [R,4]).
    return_kink_distance: also return, per row, min |pre-activation| / max |pre-activation| over all ReLU
    inputs -- rows where it is ~1e-6 sit on a ReLU kink, where the gradient is discontinuous and two
    floating-point evaluations can legitimately disagree (used by the parity tests to mask such rows).
    masks_out: a list that receives, per ReLU layer (deformnet.0..7, pts_layers.0, rot_layers.0), the tuple
    (pre-activation z, z > 0).  masks_in: a list of ten boolean [R, 256] tensors -- the ReLU of layer i becomes
    z * masks_in[i], i.e. the network is evaluated on the linear piece selected by a GIVEN activation pattern (the one
    the CUDA forward chose); with the pattern fixed, the function is smooth and every row's gradient is comparable.
    """
    R = pts.shape[0]
    if not torch.is_tensor(t):
        # reference builds torch.tensor([t]) -> float32 regardless of double input (:219)
        t = torch.full((R, 1), float(torch.tensor([t]).item()), dtype=pts.dtype, device=pts.device)
    if latent.dim() == 1:
        latent = latent[None, :].expand(R, -1)
    h0 = torch.cat([posenc(pts, PTS_FREQS), posenc(t, TIME_FREQS), latent], dim=-1)
    h = h0
    kink = None

    def relu_layer(x, W, b):
        nonlocal kink
        z = F.linear(x, W, b)
        if return_kink_distance:
            with torch.no_grad():
                d = z.abs().min(dim=-1).values / z.abs().max().clamp_min(1e-30)
                kink = d if kink is None else torch.minimum(kink, d)
        if masks_out is not None:
            masks_out.append((z.detach(), z.detach() > 0))
        if masks_in is not None:
            m = masks_in[len(used)]
            used.append(1)
            return z * m.to(z.dtype)
        return F.relu(z)

    used = []
    for i in range(DEPTH):
        W, b = params[i]
        h = relu_layer(h, W, b)
        if i == SKIP_AFTER:
            h = torch.cat([h0, h], dim=-1)
    hp = relu_layer(h, *params[8])
    dxyz = F.linear(hp, *params[9])
    hr = relu_layer(h, *params[10])
    dquat = F.linear(hr, *params[11])
    if return_kink_distance:
        return dxyz, dquat, kink
    return dxyz, dquat


def build_rotation(q):
    """renderer/latent_gs_renderer.py:89-133 (build_rotation / build_rotation_3d):
    normalises q=(r,x,y,z) then the standard rotation matrix.  q [...,4] -> [...,3,3]."""
    norm = torch.sqrt(q[..., 0] * q[..., 0] + q[..., 1] * q[..., 1] + q[..., 2] * q[..., 2] + q[..., 3] * q[..., 3])
    q = q / norm[..., None]
    r, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    rows = [
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y),
    ]
    return torch.stack(rows, dim=-1).reshape(q.shape[:-1] + (3, 3))


def quat_mul(q1, q2):
    """renderer/latent_gs_renderer.py:135-147."""
    r1, x1, y1, z1 = q1.unbind(-1)
    r2, x2, y2, z2 = q2.unbind(-1)
    return torch.stack([
        r1 * r2 - x1 * x2 - y1 * y2 - z1 * z2,
        r1 * x2 + x1 * r2 + y1 * z2 - z1 * y2,
        r1 * y2 - x1 * z2 + y1 * r2 + z1 * x2,
        r1 * z2 + x1 * y2 - y1 * x2 + z1 * r2,
    ], dim=-1)


def lbs_weights(neighbor_dists, c_radius_act, neighbor_indices):
    """renderer/latent_gs_renderer.py:1192-1199.  c_radius_act = exp(_c_radius) [M,1]."""
    r_n = c_radius_act[neighbor_indices][:, :, 0]
    w = torch.exp(-1.0 * neighbor_dists ** 2 / (2.0 * (r_n ** 2)))
    w = w + LBS_EPS
    return F.normalize(w, p=1)


def lbs_deform(xyz, rot, c_xyz, c_radius_act, dxyz, dquat, neighbor_indices, neighbor_dists):
    """Stage-s2 skinning, renderer/latent_gs_renderer.py:1191-1209 + :1219 (local_frame=True).

    xyz [N,3], rot [N,4] (raw canonical quaternion), c_xyz [M,3], c_radius_act [M,1],
    dxyz [M,3], dquat [M,4], neighbor_indices [N,K] int64, neighbor_dists [N,K] (Euclidean).
    Returns means3D [N,3], rotations [N,4] (L2-normalised).
    """
    w = lbs_weights(neighbor_dists, c_radius_act, neighbor_indices)
    c_n = c_xyz[neighbor_indices]                 # N,K,3
    d_n = dxyz[neighbor_indices]                  # N,K,3
    q_n = dquat[neighbor_indices]                 # N,K,4
    Rn = build_rotation(q_n)                      # N,K,3,3
    local = (Rn @ (xyz[:, None] - c_n)[..., None]).squeeze(-1)
    pts = (w[..., None] * (local + c_n + d_n)).sum(dim=1)
    q_blend = (w[..., None] * q_n).sum(dim=1)
    rot_out = quat_mul(q_blend, rot)
    return pts, F.normalize(rot_out)


def s1_deform(xyz, rot, dxyz):
    """Stage-s1: renderer/latent_gs_renderer.py:1211-1212, :1219 (rotation only normalised)."""
    return xyz + dxyz, F.normalize(rot)


def activations(scaling, opacity):
    """renderer/latent_gs_renderer.py:257-265: exp / sigmoid."""
    return torch.exp(scaling), torch.sigmoid(opacity)
