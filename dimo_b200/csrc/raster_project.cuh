// Projection geometry shared by the forward (raster_preprocess.cu, compiled with -fmad=false so that radii / tile
// rectangles / sort keys are reproducible op-for-op by the oracle) and the backward (raster_preprocess_bwd.cu, FMA
// contraction on: gradients are compared at 1e-4, and the recomputed geometry feeds no integer decision there apart
// from the comparisons on stored inputs).
#pragma once
#include "common.cuh"

namespace dimo {


static __device__ __constant__ float SH_C0 = 0.28209479177387814f;
static __device__ __constant__ float SH_C1 = 0.4886025119029199f;
static __device__ __constant__ float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                          -1.0925484305920792f, 0.5462742152960396f};
static __device__ __constant__ float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
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
  // selects instead of a dynamic column index: g.R stays in registers (no local-memory copy of the struct)
  const float rk0 = g.kmin == 0 ? g.R[0][0] : (g.kmin == 1 ? g.R[0][1] : g.R[0][2]);
  const float rk1 = g.kmin == 0 ? g.R[1][0] : (g.kmin == 1 ? g.R[1][1] : g.R[1][2]);
  const float rk2 = g.kmin == 0 ? g.R[2][0] : (g.kmin == 1 ? g.R[2][1] : g.R[2][2]);
  const float dotp = dot3(rk0, rk1, rk2, cp[0] - px, cp[1] - py, cp[2] - pz);
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

}  // namespace dimo
