// Backward of the projection stage (dimo_raster_preprocess_bwd): per-frame gradients, or -- the training step -- the
// gradients of the parameters shared by all frames summed over the frames inside the kernel.  Replaces the
// preprocessCUDA backward + computeCov2DCUDA + computeColorFromSH backward of the reference's rasteriser submodule
// (diff_gauss, called from renderer/latent_gs_renderer.py:1262-1276 through autograd).
#include "raster_project.cuh"

namespace dimo {

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
#pragma unroll
    for (int c = 0; c < 3; ++c) dR[r][c] += (c == g.kmin) ? g.nsign * gnw : 0.0f;
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
__global__ void __launch_bounds__(32 * FW, DEG == 0 ? 5 : 1) preprocess_bwd_shared_kernel(
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
