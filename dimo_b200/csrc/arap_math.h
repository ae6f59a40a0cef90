// Per-vertex arithmetic of the ARAP term (Renderer.arap_loss_v2 -> utils/deform_utils.py:152-232), shared by the CUDA
// kernels (csrc/arap.cu) and a host build used by the CPU tests (tests/cpu_harness/arap_host.cpp): the SAME source is
// checked on the CPU against the reference-pinned torch formulation, so only the launch glue is GPU-only.
//
//   energy   = sum_t>=1 sum_i mult_i sum_k w_ik | (p^t_i - p^t_j) - R^t_i (p^0_i - p^0_j) |^2,   j = nbr[i][k]
//   R^t_i    = the proper rotation closest to S = sum_k w_ik e_src e_tgt^T (Kabsch; reference: torch.svd + reflection
//              fix, no gradient), R = I when the vertex's edge fan is bit-identical in at least one coordinate
//              (deform_utils.py:175-176)
//   gradient = d energy / d p (R treated as a constant, like the reference's torch.no_grad block)
//
// The rotation fit runs in double precision (a 3x3 Jacobi eigen-decomposition of S^T S; ~4k vertices x frames per
// call, so its cost is irrelevant) -- more accurate than the reference's fp32 LAPACK SVD, equal to it to ~1e-6.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DIMO_HD __host__ __device__ __forceinline__
#else
#define DIMO_HD inline
#endif

namespace dimo {
namespace arap {

constexpr int MAXK = 16;

// V (columns = eigenvectors) and eigenvalues lam of the symmetric 3x3 matrix B, sorted descending.
DIMO_HD void eig_sym3(double B[3][3], double V[3][3], double lam[3]) {
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) V[a][b] = (a == b) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 10; ++sweep) {
    const double off = fabs(B[0][1]) + fabs(B[0][2]) + fabs(B[1][2]);
    const double dia = fabs(B[0][0]) + fabs(B[1][1]) + fabs(B[2][2]);
    if (off <= 1e-30 * dia || off == 0.0) break;
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
      const double bpq = B[p][q];
      if (bpq == 0.0) continue;
      const double theta = (B[q][q] - B[p][p]) / (2.0 * bpq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) {           // B <- B J  (columns p, q)
        const double bkp = B[k][p], bkq = B[k][q];
        B[k][p] = c * bkp - s * bkq;
        B[k][q] = s * bkp + c * bkq;
      }
      for (int k = 0; k < 3; ++k) {           // B <- J^T B  (rows p, q)
        const double bpk = B[p][k], bqk = B[q][k];
        B[p][k] = c * bpk - s * bqk;
        B[q][k] = s * bpk + c * bqk;
      }
      for (int k = 0; k < 3; ++k) {           // V <- V J
        const double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - s * vkq;
        V[k][q] = s * vkp + c * vkq;
      }
    }
  }
  lam[0] = B[0][0]; lam[1] = B[1][1]; lam[2] = B[2][2];
  for (int a = 0; a < 2; ++a)                  // sort descending (3 elements: two passes of adjacent swaps)
    for (int b = 0; b < 2 - a; ++b)
      if (lam[b] < lam[b + 1]) {
        const double tl = lam[b]; lam[b] = lam[b + 1]; lam[b + 1] = tl;
        for (int k = 0; k < 3; ++k) { const double tv = V[k][b]; V[k][b] = V[k][b + 1]; V[k][b + 1] = tv; }
      }
}

// R = V U^T for S = U diag(sig) V^T, forced to det(R) = +1 by the sign of the weakest left singular vector
// (reference: R = W U^T, columns flipped where det(R) <= 0, deform_utils.py:179-191).  S == 0 -> identity.
DIMO_HD void rotation_from_covariance(const double S[3][3], double R[3][3]) {
  double B[3][3], V[3][3], lam[3];
  double smax = 0.0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) smax = fmax(smax, fabs(S[a][b]));
  if (smax == 0.0) {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) R[a][b] = (a == b) ? 1.0 : 0.0;
    return;
  }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double acc = 0.0;
      for (int k = 0; k < 3; ++k) acc += (S[k][a] / smax) * (S[k][b] / smax);      // S^T S, scaled against underflow
      B[a][b] = acc;
    }
  eig_sym3(B, V, lam);
  double U[3][3];                               // columns u1, u2, u3
  auto col_of_SV = [&](int c, double out[3]) {
    for (int a = 0; a < 3; ++a) out[a] = (S[a][0] * V[0][c] + S[a][1] * V[1][c] + S[a][2] * V[2][c]) / smax;
  };
  double u1[3], u2[3], u3[3];
  col_of_SV(0, u1);
  double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int a = 0; a < 3; ++a) u1[a] /= n1;
  col_of_SV(1, u2);
  double d = u2[0] * u1[0] + u2[1] * u1[1] + u2[2] * u1[2];
  for (int a = 0; a < 3; ++a) u2[a] -= d * u1[a];
  double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  if (n2 > 1e-7 * n1) {
    for (int a = 0; a < 3; ++a) u2[a] /= n2;
  } else {                                      // rank 1: any unit vector orthogonal to u1 (R e_src does not depend on it)
    const int m = (fabs(u1[0]) <= fabs(u1[1]) && fabs(u1[0]) <= fabs(u1[2])) ? 0 : (fabs(u1[1]) <= fabs(u1[2]) ? 1 : 2);
    double e[3] = {0.0, 0.0, 0.0};
    e[m] = 1.0;
    d = u1[m];
    for (int a = 0; a < 3; ++a) u2[a] = e[a] - d * u1[a];
    n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    for (int a = 0; a < 3; ++a) u2[a] /= n2;
  }
  u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
  u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
  u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
  const double detV = V[0][0] * (V[1][1] * V[2][2] - V[1][2] * V[2][1]) - V[0][1] * (V[1][0] * V[2][2] - V[1][2] * V[2][0]) +
                      V[0][2] * (V[1][0] * V[2][1] - V[1][1] * V[2][0]);
  const double sg = detV < 0.0 ? -1.0 : 1.0;    // det(U) = +1 by construction, so det(R) = det(V)
  for (int a = 0; a < 3; ++a) { U[a][0] = u1[a]; U[a][1] = u2[a]; U[a][2] = sg * u3[a]; }
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) R[a][b] = V[a][0] * U[b][0] + V[a][1] * U[b][1] + V[a][2] * U[b][2];
}

// One (frame t, vertex i) term.  p0 / pt: [M,3] positions of frame 0 / frame t; nbr_i: the K neighbour slots of i
// (-1 = empty).  Returns the energy of the vertex (times mult) and ADDS its gradient into g0 / gt ([M,3]) through Add
// (atomicAdd on the device, += on the host).
template <typename Add>
DIMO_HD float vertex_term(int K, const float* p0, const float* pt, const int64_t* nbr_i, int i, float mult, float* g0,
                          float* gt, Add add) {
  float es[MAXK][3], et[MAXK][3];
  bool same[3] = {true, true, true};
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < K; ++k) {
    const int64_t j = nbr_i[k];
    for (int c = 0; c < 3; ++c) {
      es[k][c] = j >= 0 ? p0[3 * i + c] - p0[3 * j + c] : 0.0f;
      et[k][c] = j >= 0 ? pt[3 * i + c] - pt[3 * j + c] : 0.0f;
      same[c] = same[c] && (es[k][c] == et[k][c]);
    }
    if (j >= 0)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) S[a][b] += (double)es[k][a] * (double)et[k][b];
  }
  if (same[0] || same[1] || same[2])
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) S[a][b] = 0.0;
  double Rd[3][3];
  rotation_from_covariance(S, Rd);
  float R[3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) R[a][b] = (float)Rd[a][b];
  float energy = 0.0f;
  float gi_t[3] = {0, 0, 0}, gi_0[3] = {0, 0, 0};
  for (int k = 0; k < K; ++k) {
    const int64_t j = nbr_i[k];
    if (j < 0) continue;
    float r[3];
    for (int a = 0; a < 3; ++a)
      r[a] = et[k][a] - (R[a][0] * es[k][0] + R[a][1] * es[k][1] + R[a][2] * es[k][2]);
    energy += r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    for (int a = 0; a < 3; ++a) {
      const float dt = 2.0f * mult * r[a];                                              // d / d e_tgt
      const float d0 = -2.0f * mult * (R[0][a] * r[0] + R[1][a] * r[1] + R[2][a] * r[2]);  // d / d e_src = -2 R^T r
      gi_t[a] += dt;
      gi_0[a] += d0;
      if (gt) add(gt + 3 * j + a, -dt);
      if (g0) add(g0 + 3 * j + a, -d0);
    }
  }
  for (int a = 0; a < 3; ++a) {
    if (gt) add(gt + 3 * i + a, gi_t[a]);
    if (g0) add(g0 + 3 * i + a, gi_0[a]);
  }
  return mult * energy;
}

// Connectivity of vertex i (deform_utils.py:115-150): per frame the first Kq points in INDEX order with squared
// distance < r2 (pytorch3d ball_query), the first hit dropped (assumed to be the vertex itself), intersected over all T
// frames; written ascending into nbr_i[K] (-1 padded).  Returns the count.  nodes [T,M,3].
DIMO_HD int common_neighbours(int T, int M, int Kq, int K, float r2, const float* nodes, int i, int64_t* nbr_i) {
  int cand[MAXK];
  int ncand = 0;
  for (int t = 0; t < T; ++t) {
    const float* p = nodes + (int64_t)t * M * 3;
    const float qx = p[3 * i], qy = p[3 * i + 1], qz = p[3 * i + 2];
    int list[MAXK + 1];
    int cnt = 0;
    for (int j = 0; j < M && cnt < Kq; ++j) {
      const float dx = p[3 * j] - qx, dy = p[3 * j + 1] - qy, dz = p[3 * j + 2] - qz;
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      if (d2 < r2) list[cnt++] = j;
    }
    if (t == 0) {
      for (int c = 1; c < cnt; ++c) cand[ncand++] = list[c];
    } else {
      int keep = 0;
      for (int a = 0; a < ncand; ++a) {
        bool found = false;
        for (int c = 1; c < cnt; ++c) found = found || (list[c] == cand[a]);
        if (found) cand[keep++] = cand[a];
      }
      ncand = keep;
    }
  }
  for (int k = 0; k < K; ++k) nbr_i[k] = k < ncand ? (int64_t)cand[k] : (int64_t)-1;
  return ncand < K ? ncand : K;
}

}  // namespace arap
}  // namespace dimo
