// Per-Gaussian projection (forward + backward) for the tile rasteriser.
//
// Replaces the "preprocess" stage inside diff_gauss / diff_gaussian_rasterization's
// GaussianRasterizer (call sites renderer/latent_gs_renderer.py:1256-1266, 1268-1277).
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false.  The integer outputs of the rasteriser
// (radii, tile rectangles, tile counts, depth keys) are functions of the fp32 values computed
// here; with FMA contraction off, every operation below is a separately rounded IEEE fp32
// operation in exactly the order written, which oracle/raster.py::preprocess mirrors line by line.
// The kernel is HBM-bound (56 B read + 72 B written per Gaussian), so the lost FMA fusion is free.
#include "common.cuh"

namespace dimo {

__device__ __constant__ float SH_C0 = 0.28209479177387814f;
__device__ __constant__ float SH_C1 = 0.4886025119029199f;
__device__ __constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                          -1.0925484305920792f, 0.5462742152960396f};
__device__ __constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                                          0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                                          -0.5900435899266435f};

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  return (a0 * b0 + a1 * b1) + a2 * b2;
}

// tile rectangle of a splat; shared by preprocess and the key-emission kernel (raster_bin.cu)
__device__ __forceinline__ void tile_rect(float px, float py, float rad, int gx, int gy, int& x0, int& y0,
                                          int& x1, int& y1) {
  const float inv = 1.0f / TILE;  // exact (power of two)
  x0 = (int)fminf((float)gx, fmaxf(0.0f, truncf((px - rad) * inv)));
  y0 = (int)fminf((float)gy, fmaxf(0.0f, truncf((py - rad) * inv)));
  x1 = (int)fminf((float)gx, fmaxf(0.0f, truncf(((px + rad) + (float)(TILE - 1)) * inv)));
  y1 = (int)fminf((float)gy, fmaxf(0.0f, truncf(((py + rad) + (float)(TILE - 1)) * inv)));
}

struct Geo {      // everything the backward pass needs again
  float tvx, tvy, tvz, hx, hy, hw, p_w;
  float s[3];     // modified scales
  float R[3][3];
  float L[3][3];
  float S[3][3];
  float fx, fy, tx, ty;
  bool clampx, clampy;
  float clampvx, clampvy;
  float J00, J02, J11, J12;
  float M0[3], M1[3], v0[3], v1[3];
  float ca, cb, cc, det, det_inv;
  float conic_a, conic_b, conic_c;
  float radius_f, pix_x, pix_y;
  int kmin;
  float nsign;
};

__device__ __forceinline__ void project(const float* __restrict__ cam, float px, float py, float pz,
                                        const float sc[3], const float q[4], float scale_modifier, int W, int H,
                                        Geo& g) {
  const float* V = cam + CAM_VIEW;
  const float* P = cam + CAM_PROJ;
  g.tvx = ((px * V[0] + py * V[4]) + pz * V[8]) + V[12];
  g.tvy = ((px * V[1] + py * V[5]) + pz * V[9]) + V[13];
  g.tvz = ((px * V[2] + py * V[6]) + pz * V[10]) + V[14];
  g.hx = ((px * P[0] + py * P[4]) + pz * P[8]) + P[12];
  g.hy = ((px * P[1] + py * P[5]) + pz * P[9]) + P[13];
  g.hw = ((px * P[3] + py * P[7]) + pz * P[11]) + P[15];
  g.p_w = 1.0f / (g.hw + W_EPS);
  const float ndc_x = g.hx * g.p_w;
  const float ndc_y = g.hy * g.p_w;

  for (int j = 0; j < 3; ++j) g.s[j] = sc[j] * scale_modifier;
  const float r = q[0], x = q[1], y = q[2], z = q[3];
  g.R[0][0] = 1.0f - 2.0f * (y * y + z * z);
  g.R[0][1] = 2.0f * (x * y - r * z);
  g.R[0][2] = 2.0f * (x * z + r * y);
  g.R[1][0] = 2.0f * (x * y + r * z);
  g.R[1][1] = 1.0f - 2.0f * (x * x + z * z);
  g.R[1][2] = 2.0f * (y * z - r * x);
  g.R[2][0] = 2.0f * (x * z - r * y);
  g.R[2][1] = 2.0f * (y * z + r * x);
  g.R[2][2] = 1.0f - 2.0f * (x * x + y * y);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) g.L[i][j] = g.R[i][j] * g.s[j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = i; j < 3; ++j) {
      g.S[i][j] = dot3(g.L[i][0], g.L[i][1], g.L[i][2], g.L[j][0], g.L[j][1], g.L[j][2]);
      g.S[j][i] = g.S[i][j];
    }

  const float tanx = cam[CAM_TANX], tany = cam[CAM_TANY];
  const float limx = FOV_CLAMP * tanx, limy = FOV_CLAMP * tany;
  g.fx = (float)W / (2.0f * tanx);
  g.fy = (float)H / (2.0f * tany);
  const float txtz = g.tvx / g.tvz, tytz = g.tvy / g.tvz;
  g.clampvx = fminf(limx, fmaxf(-limx, txtz));
  g.clampvy = fminf(limy, fmaxf(-limy, tytz));
  g.clampx = (txtz < -limx) || (txtz > limx);
  g.clampy = (tytz < -limy) || (tytz > limy);
  g.tx = g.clampvx * g.tvz;
  g.ty = g.clampvy * g.tvz;
  g.J00 = g.fx / g.tvz;
  g.J02 = -(g.fx * g.tx) / (g.tvz * g.tvz);
  g.J11 = g.fy / g.tvz;
  g.J12 = -(g.fy * g.ty) / (g.tvz * g.tvz);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g.M0[k] = g.J00 * V[4 * k + 0] + g.J02 * V[4 * k + 2];
    g.M1[k] = g.J11 * V[4 * k + 1] + g.J12 * V[4 * k + 2];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    g.v0[i] = dot3(g.S[i][0], g.S[i][1], g.S[i][2], g.M0[0], g.M0[1], g.M0[2]);
    g.v1[i] = dot3(g.S[i][0], g.S[i][1], g.S[i][2], g.M1[0], g.M1[1], g.M1[2]);
  }
  g.ca = dot3(g.M0[0], g.M0[1], g.M0[2], g.v0[0], g.v0[1], g.v0[2]) + DILATION;
  g.cb = dot3(g.M0[0], g.M0[1], g.M0[2], g.v1[0], g.v1[1], g.v1[2]);
  g.cc = dot3(g.M1[0], g.M1[1], g.M1[2], g.v1[0], g.v1[1], g.v1[2]) + DILATION;
  g.det = g.ca * g.cc - g.cb * g.cb;
  g.det_inv = 1.0f / g.det;
  g.conic_a = g.cc * g.det_inv;
  g.conic_b = -g.cb * g.det_inv;
  g.conic_c = g.ca * g.det_inv;
  const float mid = 0.5f * (g.ca + g.cc);
  const float root = sqrtf(fmaxf(mid * mid - g.det, LAMBDA_FLOOR));
  const float lam = fmaxf(mid + root, mid - root);
  g.radius_f = ceilf(RADIUS_SIGMAS * sqrtf(lam));
  g.pix_x = ((ndc_x + 1.0f) * (float)W - 1.0f) * 0.5f;
  g.pix_y = ((ndc_y + 1.0f) * (float)H - 1.0f) * 0.5f;

  // shortest axis (first minimum on ties), oriented towards campos
  g.kmin = (g.s[0] <= g.s[1] && g.s[0] <= g.s[2]) ? 0 : (g.s[1] <= g.s[2] ? 1 : 2);
  const float* cp = cam + CAM_POS;
  const float dotp = dot3(g.R[0][g.kmin], g.R[1][g.kmin], g.R[2][g.kmin], cp[0] - px, cp[1] - py, cp[2] - pz);
  g.nsign = dotp < 0.0f ? -1.0f : 1.0f;
}

// SH basis b[0..K) for unit direction (x,y,z); deg <= 3.  Matches utils/sh_utils.py:57-112.
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
  b[0] = SH_C0;
  if (deg > 0) {
    b[1] = -SH_C1 * y; b[2] = SH_C1 * z; b[3] = -SH_C1 * x;
    if (deg > 1) {
      const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
      b[4] = SH_C2[0] * xy; b[5] = SH_C2[1] * yz; b[6] = SH_C2[2] * (2.0f * zz - xx - yy);
      b[7] = SH_C2[3] * xz; b[8] = SH_C2[4] * (xx - yy);
      if (deg > 2) {
        b[9] = SH_C3[0] * y * (3.0f * xx - yy);
        b[10] = SH_C3[1] * xy * z;
        b[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
        b[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        b[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
        b[14] = SH_C3[5] * z * (xx - yy);
        b[15] = SH_C3[6] * x * (xx - 3.0f * yy);
      }
    }
  }
}

// d b[k] / d(x,y,z)
__device__ __forceinline__ void sh_basis_grad(int deg, float x, float y, float z, float* bx, float* by, float* bz) {
  bx[0] = by[0] = bz[0] = 0.0f;
  if (deg > 0) {
    bx[1] = 0; by[1] = -SH_C1; bz[1] = 0;
    bx[2] = 0; by[2] = 0; bz[2] = SH_C1;
    bx[3] = -SH_C1; by[3] = 0; bz[3] = 0;
    if (deg > 1) {
      bx[4] = SH_C2[0] * y; by[4] = SH_C2[0] * x; bz[4] = 0;
      bx[5] = 0; by[5] = SH_C2[1] * z; bz[5] = SH_C2[1] * y;
      bx[6] = SH_C2[2] * (-2.0f * x); by[6] = SH_C2[2] * (-2.0f * y); bz[6] = SH_C2[2] * (4.0f * z);
      bx[7] = SH_C2[3] * z; by[7] = 0; bz[7] = SH_C2[3] * x;
      bx[8] = SH_C2[4] * (2.0f * x); by[8] = SH_C2[4] * (-2.0f * y); bz[8] = 0;
      if (deg > 2) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        bx[9] = SH_C3[0] * 6.0f * xy; by[9] = SH_C3[0] * (3.0f * xx - 3.0f * yy); bz[9] = 0;
        bx[10] = SH_C3[1] * yz; by[10] = SH_C3[1] * xz; bz[10] = SH_C3[1] * xy;
        bx[11] = SH_C3[2] * (-2.0f * xy); by[11] = SH_C3[2] * (4.0f * zz - xx - 3.0f * yy); bz[11] = SH_C3[2] * 8.0f * yz;
        bx[12] = SH_C3[3] * (-6.0f * xz); by[12] = SH_C3[3] * (-6.0f * yz); bz[12] = SH_C3[3] * (6.0f * zz - 3.0f * xx - 3.0f * yy);
        bx[13] = SH_C3[4] * (4.0f * zz - 3.0f * xx - yy); by[13] = SH_C3[4] * (-2.0f * xy); bz[13] = SH_C3[4] * 8.0f * xz;
        bx[14] = SH_C3[5] * 2.0f * xz; by[14] = SH_C3[5] * (-2.0f * yz); bz[14] = SH_C3[5] * (xx - yy);
        bx[15] = SH_C3[6] * (3.0f * xx - 3.0f * yy); by[15] = SH_C3[6] * (-6.0f * xy); bz[15] = 0;
      }
    }
  }
}

// act_flags: bit 0 = `scales` holds log-scales (the model's raw _scaling; exp applied here), bit 1 = `opacities` holds
// logits (raw _opacity; sigmoid applied here) -- GaussianModel.get_scaling / get_opacity
// (renderer/latent_gs_renderer.py:257-265, 340-355) folded into the projection pass and its backward.
constexpr int ACT_EXP_SCALE = 1, ACT_SIGMOID_OPACITY = 2;

__global__ void __launch_bounds__(256) preprocess_fwd_kernel(
    int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags,
    const float* __restrict__ cams, const int32_t* __restrict__ frame_src,
    const float* __restrict__ means3D, int64_t means3D_bs,
    const float* __restrict__ scales, int64_t scales_bs,
    const float* __restrict__ rotations, int64_t rot_bs,
    const float* __restrict__ opacities, int64_t op_bs,
    const float* __restrict__ shs, int64_t shs_bs,
    const float* __restrict__ colors, int64_t col_bs,
    float4* __restrict__ splats, int32_t* __restrict__ radii, uint32_t* __restrict__ tiles_touched,
    uint2* __restrict__ rects, uint32_t* __restrict__ depth_keys, unsigned long long* __restrict__ total_count) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // the CTA's tile-instance count is summed (warp shuffles + one shared atomic per warp) and added to the launch
  // set's total with one global atomic per CTA: the instance count R without a scan over the B*N counters
  __shared__ unsigned int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  unsigned int my_tiles = 0;
  if (idx < (int64_t)B * N) {
  const int b = (int)(idx / N);
  const int i = (int)(idx - (int64_t)b * N);
  const float* cam = cams + (int64_t)b * DIMO_CAM_FLOATS;

  const int bsrc = frame_src != nullptr ? frame_src[b] : b;   // deformation block of this frame ((motion, t) pair)
  const float* pm = means3D + bsrc * means3D_bs + 3 * (int64_t)i;
  const float px = pm[0], py = pm[1], pz = pm[2];
  const float* ps = scales + b * scales_bs + 3 * (int64_t)i;
  float sc[3] = {ps[0], ps[1], ps[2]};
  if (act_flags & ACT_EXP_SCALE) { sc[0] = expf(sc[0]); sc[1] = expf(sc[1]); sc[2] = expf(sc[2]); }
  const float4 q4 = *reinterpret_cast<const float4*>(rotations + bsrc * rot_bs + 4 * (int64_t)i);
  const float q[4] = {q4.x, q4.y, q4.z, q4.w};

  Geo g;
  project(cam, px, py, pz, sc, q, scale_modifier, W, H, g);

  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const bool finite = isfinite(g.pix_x) && isfinite(g.pix_y) && isfinite(g.radius_f);
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  if (finite) tile_rect(g.pix_x, g.pix_y, g.radius_f, gx, gy, x0, y0, x1, y1);
  const int ntiles = (x1 - x0) * (y1 - y0);
  const bool visible = (g.tvz > NEAR_CULL_Z) && (g.det != 0.0f) && finite && (ntiles > 0);

  float4* out = splats + 4 * idx;
  if (!visible) {
    radii[idx] = 0;
    tiles_touched[idx] = 0;
    rects[idx] = make_uint2(0u, 0u);
    depth_keys[idx] = 0xFFFFFFFFu;           // culled splats sort to the end of their frame and emit nothing
    const float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
    out[0] = zz; out[1] = zz; out[2] = zz; out[3] = zz;
  } else {

  float rgb[3];
  if (colors != nullptr) {
    const float* pc = colors + b * col_bs + 3 * (int64_t)i;
    rgb[0] = pc[0]; rgb[1] = pc[1]; rgb[2] = pc[2];
  } else {
    const float* cp = cam + CAM_POS;
    float dx = px - cp[0], dy = py - cp[1], dz = pz - cp[2];
    const float len = sqrtf((dx * dx + dy * dy) + dz * dz);
    dx = dx / len; dy = dy / len; dz = dz / len;
    float basis[16];
    sh_basis(sh_degree, dx, dy, dz, basis);
    const int K = (sh_degree + 1) * (sh_degree + 1);
    const float* sh = shs + b * shs_bs + (int64_t)i * sh_coeffs * 3;
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    for (int k = 0; k < K; ++k) {
      rgb[0] += basis[k] * sh[3 * k + 0];
      rgb[1] += basis[k] * sh[3 * k + 1];
      rgb[2] += basis[k] * sh[3 * k + 2];
    }
    rgb[0] = fmaxf(rgb[0] + 0.5f, 0.0f);
    rgb[1] = fmaxf(rgb[1] + 0.5f, 0.0f);
    rgb[2] = fmaxf(rgb[2] + 0.5f, 0.0f);
  }

  const float* V = cam + CAM_VIEW;
  const float nw0 = g.R[0][g.kmin] * g.nsign, nw1 = g.R[1][g.kmin] * g.nsign, nw2 = g.R[2][g.kmin] * g.nsign;
  const float nx = dot3(nw0, nw1, nw2, V[0], V[4], V[8]);
  const float ny = dot3(nw0, nw1, nw2, V[1], V[5], V[9]);
  const float nz = dot3(nw0, nw1, nw2, V[2], V[6], V[10]);
  float op = opacities[b * op_bs + i];
  if (act_flags & ACT_SIGMOID_OPACITY) op = 1.0f / (1.0f + expf(-op));

  radii[idx] = (int)g.radius_f;
  tiles_touched[idx] = (uint32_t)ntiles;
  my_tiles = (unsigned int)ntiles;
  rects[idx] = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));   // [x0,x1) x [y0,y1), tiles
  depth_keys[idx] = __float_as_uint(g.tvz);   // view depth > 0.2: the bits order as unsigned integers
  // blend record (include/dimo_b200.h; consumed as-is by raster_blend.cu, gathered by index):
  //   x, y, a2, b2 | c2, opacity, pthr2, r | g, b, depth, nx | ny, nz, own index, 0
  // the conic is pre-multiplied into log2 units (alpha = opacity * 2^(a2 dx^2 + b2 dx dy + c2 dy^2));
  // pairs with p2 < pthr2 cannot reach alpha >= 1/255 (0.01 in log2 units = 0.7 % safety margin on alpha)
  constexpr float LOG2E = 1.4426950408889634f;
  const float pthr2 = -log2f(255.0f * op) - 0.01f;
  out[0] = make_float4(g.pix_x, g.pix_y, (-0.5f * LOG2E) * g.conic_a, -LOG2E * g.conic_b);
  out[1] = make_float4((-0.5f * LOG2E) * g.conic_c, op, pthr2, rgb[0]);
  out[2] = make_float4(rgb[1], rgb[2], g.tvz, nx);
  out[3] = make_float4(ny, nz, __uint_as_float((uint32_t)idx), 0.f);
  }   // visible
  }   // idx < B*N
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_tiles += __shfl_xor_sync(0xffffffffu, my_tiles, o);
  if ((threadIdx.x & 31) == 0 && my_tiles) atomicAdd(&s_cnt, my_tiles);
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(total_count, (unsigned long long)s_cnt);
}

// everything one (frame, Gaussian) contributes to the backward of the projection (shared by the two kernels below)
struct PreBwdItem {
  float dmean[3], ndcx, ndcy, dscale[3], dop;
  float4 dq;
  float gm[3];          // colour gradient (after the SH clamp mask when SHs are evaluated)
  float basis[16];      // SH basis of the view direction: dL/dsh[k][c] = basis[k] * gm[c] for k < (deg+1)^2
};

__device__ __forceinline__ void preprocess_bwd_item(
    int b, int i, int64_t idx, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags,
    const float* __restrict__ opacities, int64_t op_bs, const float* __restrict__ cams,
    const int32_t* __restrict__ frame_src, const float* __restrict__ means3D, int64_t means3D_bs,
    const float* __restrict__ scales, int64_t scales_bs, const float* __restrict__ rotations, int64_t rot_bs,
    const float* __restrict__ shs, int64_t shs_bs, const float4* __restrict__ dL_dsplats, PreBwdItem& r) {
  const int K = (sh_degree + 1) * (sh_degree + 1);
  const float* cam = cams + (int64_t)b * DIMO_CAM_FLOATS;
  const float* V = cam + CAM_VIEW;
  const float* P = cam + CAM_PROJ;
  const int bsrc = frame_src != nullptr ? frame_src[b] : b;   // deformation block of this frame ((motion, t) pair)
  const float* pm = means3D + bsrc * means3D_bs + 3 * (int64_t)i;
  const float px = pm[0], py = pm[1], pz = pm[2];
  const float* ps = scales + b * scales_bs + 3 * (int64_t)i;
  float sc[3] = {ps[0], ps[1], ps[2]};
  if (act_flags & ACT_EXP_SCALE) { sc[0] = expf(sc[0]); sc[1] = expf(sc[1]); sc[2] = expf(sc[2]); }
  const float4 q4 = *reinterpret_cast<const float4*>(rotations + bsrc * rot_bs + 4 * (int64_t)i);
  const float q[4] = {q4.x, q4.y, q4.z, q4.w};
  Geo g;
  project(cam, px, py, pz, sc, q, scale_modifier, W, H, g);

  const float4 d0 = dL_dsplats[4 * idx + 0], d1 = dL_dsplats[4 * idx + 1], d2 = dL_dsplats[4 * idx + 2],
               d3 = dL_dsplats[4 * idx + 3];
  const float g_px = d0.x, g_py = d0.y, gA = d0.z, gB = d0.w, gC = d1.x, g_op = d1.y;
  const float g_rgb[3] = {d1.z, d1.w, d2.x};
  const float g_depth = d2.y;
  const float g_n[3] = {d2.z, d2.w, d3.x};

  float dmean[3] = {0.f, 0.f, 0.f};

  // ---- colour ----
  r.gm[0] = g_rgb[0]; r.gm[1] = g_rgb[1]; r.gm[2] = g_rgb[2];
  if (shs != nullptr) {
    const float* cp = cam + CAM_POS;
    float dx = px - cp[0], dy = py - cp[1], dz = pz - cp[2];
    const float len = sqrtf((dx * dx + dy * dy) + dz * dz);
    const float ux = dx / len, uy = dy / len, uz = dz / len;
    float* basis = r.basis;
    sh_basis(sh_degree, ux, uy, uz, basis);
    const float* sh = shs + b * shs_bs + (int64_t)i * sh_coeffs * 3;
    float rgb[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < K; ++k) {
      rgb[0] += basis[k] * sh[3 * k + 0]; rgb[1] += basis[k] * sh[3 * k + 1]; rgb[2] += basis[k] * sh[3 * k + 2];
    }
    float* gm = r.gm;
    for (int c = 0; c < 3; ++c) gm[c] = (rgb[c] + 0.5f < 0.0f) ? 0.0f : g_rgb[c];
    if (sh_degree > 0) {
      float bx[16], by[16], bz[16];
      sh_basis_grad(sh_degree, ux, uy, uz, bx, by, bz);
      float gdx = 0.f, gdy = 0.f, gdz = 0.f;
      for (int k = 1; k < K; ++k) {
        const float w = gm[0] * sh[3 * k + 0] + gm[1] * sh[3 * k + 1] + gm[2] * sh[3 * k + 2];
        gdx += bx[k] * w; gdy += by[k] * w; gdz += bz[k] * w;
      }
      const float dotg = ux * gdx + uy * gdy + uz * gdz;
      dmean[0] += (gdx - ux * dotg) / len;
      dmean[1] += (gdy - uy * dotg) / len;
      dmean[2] += (gdz - uz * dotg) / len;
    }
  }

  // ---- conic -> cov2D ----
  const float A = g.conic_a, Bc = g.conic_b, C = g.conic_c, hgB = 0.5f * gB;
  const float k00 = A * gA + Bc * hgB, k01 = A * hgB + Bc * gC, k10 = Bc * gA + C * hgB, k11 = Bc * hgB + C * gC;
  const float d_ca = -(k00 * A + k01 * Bc);
  const float d_cb = -2.0f * (k00 * Bc + k01 * C);
  const float d_cc = -(k10 * Bc + k11 * C);

  // ---- cov2D -> M0, M1, L ----
  float dM0[3], dM1[3], u0[3], u1[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dM0[k] = 2.0f * d_ca * g.v0[k] + d_cb * g.v1[k];
    dM1[k] = d_cb * g.v0[k] + 2.0f * d_cc * g.v1[k];
    u0[k] = g.M0[0] * g.L[0][k] + g.M0[1] * g.L[1][k] + g.M0[2] * g.L[2][k];
    u1[k] = g.M1[0] * g.L[0][k] + g.M1[1] * g.L[1][k] + g.M1[2] * g.L[2][k];
  }
  float dR[3][3];
  float dscale[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float dLL = 2.0f * d_ca * g.M0[r] * u0[c] + d_cb * (g.M0[r] * u1[c] + g.M1[r] * u0[c]) +
                        2.0f * d_cc * g.M1[r] * u1[c];
      dscale[c] += dLL * g.R[r][c];
      dR[r][c] = dLL * g.s[c];
    }

  // ---- normal -> R column kmin ----
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float gnw = g_n[0] * V[4 * r + 0] + g_n[1] * V[4 * r + 1] + g_n[2] * V[4 * r + 2];
    dR[r][g.kmin] += g.nsign * gnw;
  }

  // ---- J -> t ----
  const float dJ00 = dM0[0] * V[0] + dM0[1] * V[4] + dM0[2] * V[8];
  const float dJ02 = dM0[0] * V[2] + dM0[1] * V[6] + dM0[2] * V[10];
  const float dJ11 = dM1[0] * V[1] + dM1[1] * V[5] + dM1[2] * V[9];
  const float dJ12 = dM1[0] * V[2] + dM1[1] * V[6] + dM1[2] * V[10];
  const float tz = g.tvz, tz2 = 1.0f / (tz * tz), tz3 = tz2 / tz;
  const float dtx = -g.fx * tz2 * dJ02;
  const float dty = -g.fy * tz2 * dJ12;
  float dtz = -g.fx * tz2 * dJ00 - g.fy * tz2 * dJ11 + 2.0f * g.fx * g.tx * tz3 * dJ02 + 2.0f * g.fy * g.ty * tz3 * dJ12;
  float dtvx = g.clampx ? 0.0f : dtx;
  float dtvy = g.clampy ? 0.0f : dty;
  if (g.clampx) dtz += g.clampvx * dtx;
  if (g.clampy) dtz += g.clampvy * dty;
  dtz += g_depth;

  // ---- pixel centre -> homogeneous ----
  const float g_ndcx = g_px * 0.5f * (float)W, g_ndcy = g_py * 0.5f * (float)H;
  const float dhx = g_ndcx * g.p_w, dhy = g_ndcy * g.p_w;
  const float dhw = -(g_ndcx * g.hx + g_ndcy * g.hy) * g.p_w * g.p_w;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    dmean[r] += V[4 * r + 0] * dtvx + V[4 * r + 1] * dtvy + V[4 * r + 2] * dtz;
    dmean[r] += P[4 * r + 0] * dhx + P[4 * r + 1] * dhy + P[4 * r + 3] * dhw;
  }

  // ---- R -> quaternion ----
  const float qr = q[0], x = q[1], y = q[2], z = q[3];
  float4 dq;
  dq.x = 2.0f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
  dq.y = 2.0f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.0f * x * dR[1][1] - qr * dR[1][2] + z * dR[2][0] +
                 qr * dR[2][1] - 2.0f * x * dR[2][2]);
  dq.z = 2.0f * (-2.0f * y * dR[0][0] + x * dR[0][1] + qr * dR[0][2] + x * dR[1][0] + z * dR[1][2] - qr * dR[2][0] +
                 z * dR[2][1] - 2.0f * y * dR[2][2]);
  dq.w = 2.0f * (-2.0f * z * dR[0][0] - qr * dR[0][1] + x * dR[0][2] + qr * dR[1][0] - 2.0f * z * dR[1][1] +
                 y * dR[1][2] + x * dR[2][0] + y * dR[2][1]);

  r.dmean[0] = dmean[0]; r.dmean[1] = dmean[1]; r.dmean[2] = dmean[2];
  r.ndcx = g_ndcx; r.ndcy = g_ndcy;
  // d exp(x) = exp(x): the gradient lands on the log-scales when the activation is folded in
  const float e0 = (act_flags & ACT_EXP_SCALE) ? sc[0] : 1.0f, e1 = (act_flags & ACT_EXP_SCALE) ? sc[1] : 1.0f,
              e2 = (act_flags & ACT_EXP_SCALE) ? sc[2] : 1.0f;
  r.dscale[0] = dscale[0] * scale_modifier * e0;
  r.dscale[1] = dscale[1] * scale_modifier * e1;
  r.dscale[2] = dscale[2] * scale_modifier * e2;
  r.dq = dq;
  float g_opacity = g_op;
  if (act_flags & ACT_SIGMOID_OPACITY) {            // d sigmoid(x) = s (1 - s)
    const float sg = 1.0f / (1.0f + expf(-opacities[b * op_bs + i]));
    g_opacity = g_op * sg * (1.0f - sg);
  }
  r.dop = g_opacity;
}

// one thread per (frame, Gaussian); every gradient is written per frame (dL_dmeans2D may be NULL)
__global__ void __launch_bounds__(256) preprocess_bwd_kernel(
    int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags,
    const float* __restrict__ opacities, int64_t op_bs,
    const float* __restrict__ cams, const int32_t* __restrict__ frame_src,
    const float* __restrict__ means3D, int64_t means3D_bs,
    const float* __restrict__ scales, int64_t scales_bs,
    const float* __restrict__ rotations, int64_t rot_bs,
    const float* __restrict__ shs, int64_t shs_bs,
    const int32_t* __restrict__ radii, const float4* __restrict__ dL_dsplats,
    float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dscales,
    float4* __restrict__ dL_drot, float* __restrict__ dL_dop, float* __restrict__ dL_dshs,
    float* __restrict__ dL_dcolors) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * N) return;
  const int b = (int)(idx / N);
  const int i = (int)(idx - (int64_t)b * N);
  const int K = (sh_degree + 1) * (sh_degree + 1);

  if (radii[idx] <= 0) {
    dL_dmeans3D[3 * idx + 0] = 0.f; dL_dmeans3D[3 * idx + 1] = 0.f; dL_dmeans3D[3 * idx + 2] = 0.f;
    if (dL_dmeans2D) { dL_dmeans2D[3 * idx + 0] = 0.f; dL_dmeans2D[3 * idx + 1] = 0.f; dL_dmeans2D[3 * idx + 2] = 0.f; }
    dL_dscales[3 * idx + 0] = 0.f; dL_dscales[3 * idx + 1] = 0.f; dL_dscales[3 * idx + 2] = 0.f;
    dL_drot[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    dL_dop[idx] = 0.f;
    if (dL_dshs) for (int k = 0; k < sh_coeffs * 3; ++k) dL_dshs[idx * sh_coeffs * 3 + k] = 0.f;
    if (dL_dcolors) { dL_dcolors[3 * idx + 0] = 0.f; dL_dcolors[3 * idx + 1] = 0.f; dL_dcolors[3 * idx + 2] = 0.f; }
    return;
  }
  PreBwdItem r;
  preprocess_bwd_item(b, i, idx, N, W, H, sh_degree, sh_coeffs, scale_modifier, act_flags, opacities, op_bs, cams, frame_src,
                      means3D, means3D_bs, scales, scales_bs, rotations, rot_bs, dL_dcolors ? nullptr : shs, shs_bs,
                      dL_dsplats, r);
  if (dL_dcolors) {
    dL_dcolors[3 * idx + 0] = r.gm[0]; dL_dcolors[3 * idx + 1] = r.gm[1]; dL_dcolors[3 * idx + 2] = r.gm[2];
  } else {
    float* o = dL_dshs + idx * sh_coeffs * 3;
    for (int k = 0; k < sh_coeffs; ++k) {
      const float bk = k < K ? r.basis[k] : 0.0f;
      o[3 * k + 0] = bk * r.gm[0]; o[3 * k + 1] = bk * r.gm[1]; o[3 * k + 2] = bk * r.gm[2];
    }
  }
  dL_dmeans3D[3 * idx + 0] = r.dmean[0]; dL_dmeans3D[3 * idx + 1] = r.dmean[1]; dL_dmeans3D[3 * idx + 2] = r.dmean[2];
  if (dL_dmeans2D) { dL_dmeans2D[3 * idx + 0] = r.ndcx; dL_dmeans2D[3 * idx + 1] = r.ndcy; dL_dmeans2D[3 * idx + 2] = 0.f; }
  dL_dscales[3 * idx + 0] = r.dscale[0]; dL_dscales[3 * idx + 1] = r.dscale[1]; dL_dscales[3 * idx + 2] = r.dscale[2];
  dL_drot[idx] = r.dq;
  dL_dop[idx] = r.dop;
}

// The training step shares scales, opacities and SH coefficients between all B frames: their gradients are sums over
// the frames.  One CTA = 32 consecutive Gaussians x FW warps; warp w walks the frames w, w + FW, ... with the sums in
// registers, the FW partial sums meet in shared memory (fixed order: deterministic) and leave as [N, *] tensors.  Compared
// with per-frame outputs + dimo_segment_sum this drops 212 B written and read back per (frame, Gaussian) (192 B of it
// the SH gradient), i.e. ~60 % of the projection backward's HBM traffic at the bench shape.  Per-frame outputs
// (means3D, means2D, rotations) are written as before.
template <int DEG, int FW>
__global__ void __launch_bounds__(32 * FW) preprocess_bwd_shared_kernel(
    int B, int N, int W, int H, int sh_coeffs, float scale_modifier, int act_flags,
    const float* __restrict__ opacities, const float* __restrict__ cams, const int32_t* __restrict__ frame_src,
    const float* __restrict__ means3D, int64_t means3D_bs, const float* __restrict__ scales,
    const float* __restrict__ rotations, int64_t rot_bs, const float* __restrict__ shs,
    const int32_t* __restrict__ radii, const float4* __restrict__ dL_dsplats,
    float* __restrict__ dL_dmeans3D, float* __restrict__ dL_dmeans2D, float* __restrict__ dL_dscales,
    float4* __restrict__ dL_drot, float* __restrict__ dL_dop, float* __restrict__ dL_dshs, int accumulate) {
  constexpr int K = (DEG + 1) * (DEG + 1);
  constexpr int V = 4 + 3 * K;                       // dscale (3), dop (1), dsh (3 K)
  __shared__ float red[FW][V][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i0 = blockIdx.x * 32;
  const int i = i0 + lane;
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  if (i < N) {
    for (int b = w; b < B; b += FW) {
      const int64_t idx = (int64_t)b * N + i;
      if (radii[idx] <= 0) {
        dL_dmeans3D[3 * idx + 0] = 0.f; dL_dmeans3D[3 * idx + 1] = 0.f; dL_dmeans3D[3 * idx + 2] = 0.f;
        if (dL_dmeans2D) { dL_dmeans2D[3 * idx + 0] = 0.f; dL_dmeans2D[3 * idx + 1] = 0.f; dL_dmeans2D[3 * idx + 2] = 0.f; }
        dL_drot[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        continue;
      }
      PreBwdItem r;
      preprocess_bwd_item(b, i, idx, N, W, H, DEG, sh_coeffs, scale_modifier, act_flags, opacities, 0, cams, frame_src,
                          means3D, means3D_bs, scales, 0, rotations, rot_bs, shs, 0, dL_dsplats, r);
      dL_dmeans3D[3 * idx + 0] = r.dmean[0]; dL_dmeans3D[3 * idx + 1] = r.dmean[1]; dL_dmeans3D[3 * idx + 2] = r.dmean[2];
      if (dL_dmeans2D) { dL_dmeans2D[3 * idx + 0] = r.ndcx; dL_dmeans2D[3 * idx + 1] = r.ndcy; dL_dmeans2D[3 * idx + 2] = 0.f; }
      dL_drot[idx] = r.dq;
      acc[0] += r.dscale[0]; acc[1] += r.dscale[1]; acc[2] += r.dscale[2]; acc[3] += r.dop;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        acc[4 + 3 * k] += r.basis[k] * r.gm[0]; acc[5 + 3 * k] += r.basis[k] * r.gm[1]; acc[6 + 3 * k] += r.basis[k] * r.gm[2];
      }
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) red[w][v][lane] = acc[v];
  __syncthreads();
  const int n_here = min(32, N - i0);
  // scales [N,3] and opacities [N]: 4 values per Gaussian
  for (int e = threadIdx.x; e < n_here * 4; e += 32 * FW) {
    const int il = e >> 2, v = e & 3;
    float t = 0.f;
#pragma unroll
    for (int f = 0; f < FW; ++f) t += red[f][v][il];
    float* o = v < 3 ? dL_dscales + 3 * (int64_t)(i0 + il) + v : dL_dop + i0 + il;
    *o = accumulate ? *o + t : t;
  }
  // SH gradients [N, sh_coeffs, 3]: the CTA's 32 Gaussians are one contiguous block; inactive bands are zero
  const int per = sh_coeffs * 3;
  for (int e = threadIdx.x; e < n_here * per; e += 32 * FW) {
    const int il = e / per, v = e - il * per;
    float t = 0.f;
    if (v < 3 * K) {
#pragma unroll
      for (int f = 0; f < FW; ++f) t += red[f][4 + v][il];
    }
    float* o = dL_dshs + (int64_t)i0 * per + e;
    if (!accumulate) *o = t;
    else if (v < 3 * K) *o += t;
  }
}

// out[u, :] = sum over rows s with seg[s] == u of in[s, :]  (seg == NULL: every row belongs to segment 0), rows
// added in ascending s (deterministic).  Folds the per-frame gradients of the rasteriser backward onto the inputs
// they came from: a (motion, t) deformation shared by several views, or a parameter shared by all frames.
// HBM: 4 B read per input element + 4 B written per output element.
constexpr int SEG_MAX_ROWS = 1024;
template <int VEC>
__global__ void __launch_bounds__(256) segment_sum_kernel(int S, int64_t n, const int32_t* __restrict__ seg,
                                                          const float* __restrict__ in, float* __restrict__ out) {
  __shared__ int32_t s_seg[SEG_MAX_ROWS];
  for (int k = threadIdx.x; k < S; k += blockDim.x) s_seg[k] = seg != nullptr ? seg[k] : 0;
  __syncthreads();
  const int u = blockIdx.y;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= n) return;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  for (int r = 0; r < S; ++r) {
    if (s_seg[r] != u) continue;
    const float* row = in + (int64_t)r * n + e;
    if (VEC == 4) {
      const float4 v = *reinterpret_cast<const float4*>(row);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    } else {
      acc[0] += row[0];
    }
  }
  float* o = out + (int64_t)u * n + e;
  if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  else o[0] = acc[0];
}

}  // namespace dimo

namespace dimo {

int preprocess_launch(
    int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags, const float* cams,
    const int32_t* frame_src,
    const float* means3D, int64_t means3D_bstride, const float* scales, int64_t scales_bstride,
    const float* rotations, int64_t rotations_bstride, const float* opacities, int64_t opacities_bstride,
    const float* shs, int64_t shs_bstride, const float* colors_precomp, int64_t colors_bstride,
    float* splats, int32_t* radii, uint32_t* tiles_touched, uint32_t* rects, uint32_t* depth_keys,
    unsigned long long* total_count, cudaStream_t st) {
  const int64_t BN = (int64_t)B * N;
  if (BN == 0) return 0;
  preprocess_fwd_kernel<<<ceil_div(BN, 256), 256, 0, st>>>(
      B, N, W, H, sh_degree, sh_coeffs, scale_modifier, act_flags, cams, frame_src, means3D, means3D_bstride, scales,
      scales_bstride, rotations, rotations_bstride, opacities, opacities_bstride, shs, shs_bstride, colors_precomp,
      colors_bstride, reinterpret_cast<float4*>(splats), radii, tiles_touched, reinterpret_cast<uint2*>(rects),
      depth_keys, total_count);
  DIMO_CHECK_LAUNCH();
  return 0;
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_raster_preprocess_bwd(
    int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags, const float* cams,
    const int32_t* frame_src,
    const float* means3D, int64_t means3D_bstride, const float* scales, int64_t scales_bstride,
    const float* rotations, int64_t rotations_bstride, const float* opacities, int64_t opacities_bstride,
    const float* shs, int64_t shs_bstride,
    const int32_t* radii, const float* dL_dsplats, float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dscales,
    float* dL_drotations, float* dL_dopacities, float* dL_dshs, float* dL_dcolors, int reduce_shared, void* stream) {
  const int64_t BN = (int64_t)B * N;
  if (BN == 0) return 0;
  DIMO_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
  if (reduce_shared) {
    DIMO_REQUIRE(scales_bstride == 0 && opacities_bstride == 0 && shs != nullptr && shs_bstride == 0 && dL_dshs != nullptr &&
                     dL_dcolors == nullptr && (opacities != nullptr || !(act_flags & ACT_SIGMOID_OPACITY)),
                 "reduce_shared: scales, opacities and shs must be shared by all frames (batch stride 0)");
    constexpr int FW = 4;
    const dim3 grid(ceil_div(N, 32));
    cudaStream_t st = (cudaStream_t)stream;
#define DIMO_PRE_BWD_CASE(D)                                                                                         \
  case D:                                                                                                            \
    preprocess_bwd_shared_kernel<D, FW><<<grid, 32 * FW, 0, st>>>(                                                   \
        B, N, W, H, sh_coeffs, scale_modifier, act_flags, opacities, cams, frame_src, means3D, means3D_bstride, scales, \
        rotations, rotations_bstride, shs, radii, reinterpret_cast<const float4*>(dL_dsplats), dL_dmeans3D,          \
        dL_dmeans2D, dL_dscales, reinterpret_cast<float4*>(dL_drotations), dL_dopacities, dL_dshs,                   \
        reduce_shared == 2);                                                                                         \
    break;
    switch (sh_degree) { DIMO_PRE_BWD_CASE(0) DIMO_PRE_BWD_CASE(1) DIMO_PRE_BWD_CASE(2) DIMO_PRE_BWD_CASE(3) }
#undef DIMO_PRE_BWD_CASE
    DIMO_CHECK_LAUNCH();
    return 0;
  }
  DIMO_REQUIRE((dL_dshs != nullptr) != (dL_dcolors != nullptr), "exactly one of dL_dshs / dL_dcolors");
  DIMO_REQUIRE(dL_dcolors != nullptr || shs != nullptr, "shs required when colours come from SH");
  DIMO_REQUIRE(!(act_flags & ACT_SIGMOID_OPACITY) || opacities != nullptr, "opacities (logits) required when the sigmoid is folded in");
  preprocess_bwd_kernel<<<ceil_div(BN, 256), 256, 0, (cudaStream_t)stream>>>(
      B, N, W, H, sh_degree, sh_coeffs, scale_modifier, act_flags, opacities, opacities_bstride, cams, frame_src, means3D, means3D_bstride, scales,
      scales_bstride, rotations, rotations_bstride, shs, shs_bstride, radii, reinterpret_cast<const float4*>(dL_dsplats),
      dL_dmeans3D, dL_dmeans2D, dL_dscales, reinterpret_cast<float4*>(dL_drotations), dL_dopacities, dL_dshs,
      dL_dcolors);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_segment_sum(int S, int U, int64_t n, const int32_t* seg, const float* in, float* out, void* stream) {
  DIMO_REQUIRE(S >= 0 && S <= SEG_MAX_ROWS && U >= 0 && U <= 65535 && n >= 0, "segment_sum: S <= 1024, U <= 65535");
  if (U == 0 || n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (n & 3) == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0;
  if (vec) {
    dim3 grid(ceil_div(n / 4, 256), U);
    segment_sum_kernel<4><<<grid, 256, 0, st>>>(S, n, seg, in, out);
  } else {
    dim3 grid(ceil_div(n, 256), U);
    segment_sum_kernel<1><<<grid, 256, 0, st>>>(S, n, seg, in, out);
  }
  DIMO_CHECK_LAUNCH();
  return 0;
}
