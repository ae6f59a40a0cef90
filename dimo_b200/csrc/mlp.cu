// TimeNet (renderer/latent_gs_renderer.py:184-235) building blocks: positional-encoding embedding
// (src/pos_enc.py:6-54) and fused Linear(+bias)(+ReLU) forward / data-grad / weight-grad.
//
// FP32 SIMT tiled GEMM (64x64x16 tiles, 4x4 register micro-tiles, register prefetch + double-buffered smem) with the
// epilogues fused (bias, ReLU, ReLU-mask on the incoming gradient, bias-gradient reduction, split-R
// weight-gradient accumulation).  It serves the 3- and 4-wide head layers and is the A/B reference (DIMO_TC=0) of
// the tcgen05 3xTF32 kernels in mlp_tc.cu, which run everything else.
//
// One generic kernel computes  C[i,j] (=|+=|atomic+=) sum_l A(i,l) * B(l,j)  with arbitrary element
// strides; the three layer operations pick strides so that global loads stay coalesced:
//   forward   Y[r,n]  = sum_k X[r,k]  W[n,k]      (A = X,  l contiguous; B = W, l contiguous)
//   data grad dX[r,k] = sum_n dYm[r,n] W[n,k]     (A = dY, l contiguous; B = W, j contiguous)
//   wgt grad  dW[n,k] = sum_r dYm[r,n] X[r,k]     (A = dY, i contiguous; B = X, j contiguous)
// where dYm = dY * [Y > 0] is applied on the fly while loading A.
#include "common.cuh"

namespace dimo {

constexpr int BM = 64, BN = 64, BK = 16;

struct GemmArgs {
  int I, J, L;                 // C is I x J, reduction length L
  const float* A; int64_t sAi, sAl;
  const float* mask; int64_t sMi, sMl;     // optional: A(i,l) *= [mask(i,l) > 0]
  const float* Bm; int64_t sBl, sBj;
  float* C; int64_t sCi;                   // C row stride (j contiguous)
  const float* bias;                       // per-j, forward only
  float* rowsum;                           // optional: rowsum[i] += sum_l A(i,l)  (bias gradient), blockIdx.y==0 only
  int relu;                                // epilogue ReLU
  int mode;                                // 0: C = v ; 1: C += v ; 2: atomicAdd(C, v)
  int l_per_split;                         // reduction range per blockIdx.z
  float det;                               // mode 2 / rowsum: deterministic fixed-point accumulation (common.cuh)
};

// A_LC: A's contiguous dimension is l (else i).  B_LC: B's contiguous dimension is l (else j).
template <bool A_LC, bool B_LC>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs p) {
  // double-buffered smem tiles; the next k-tile is prefetched from global into registers while the current one
  // is consumed (one __syncthreads per k-tile)
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
  const int l_begin = blockIdx.z * p.l_per_split;
  const int l_end = min(p.L, l_begin + p.l_per_split);
  const int ty = tid / 16, tx = tid % 16;     // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float rsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_rowsum = p.rowsum != nullptr && blockIdx.y == 0;

  // per-thread element coordinates inside a tile (4 A elements + 4 B elements)
  int a_ii[4], a_ll[4], b_jj[4], b_ll[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int lin = tid + e * 256;
    if (A_LC) { a_ii[e] = lin / BK; a_ll[e] = lin % BK; } else { a_ll[e] = lin / BM; a_ii[e] = lin % BM; }
    if (B_LC) { b_jj[e] = lin / BK; b_ll[e] = lin % BK; } else { b_ll[e] = lin / BN; b_jj[e] = lin % BN; }
  }
  float ra[4], rb[4];
  auto fetch = [&](int l0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int gi = i0 + a_ii[e], gl = l0 + a_ll[e];
      float v = 0.f;
      if (gi < p.I && gl < l_end) {
        v = p.A[gi * p.sAi + gl * p.sAl];
        if (p.mask != nullptr && !(p.mask[gi * p.sMi + gl * p.sMl] > 0.f)) v = 0.f;
      }
      ra[e] = v;
      const int gj = j0 + b_jj[e], gl2 = l0 + b_ll[e];
      rb[e] = (gj < p.J && gl2 < l_end) ? p.Bm[gl2 * p.sBl + gj * p.sBj] : 0.f;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      As[buf][a_ll[e]][a_ii[e]] = ra[e];
      Bs[buf][b_ll[e]][b_jj[e]] = rb[e];
    }
  };

  if (l_begin < l_end) {
    fetch(l_begin);
    stash(0);
  }
  __syncthreads();
  int buf = 0;
  for (int l0 = l_begin; l0 < l_end; l0 += BK) {
    const bool more = l0 + BK < l_end;
    if (more) fetch(l0 + BK);
#pragma unroll
    for (int l = 0; l < BK; ++l) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][l][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][l][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
      if (do_rowsum && tx == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) rsum[u] += a[u];
      }
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int gi = i0 + ty * 4 + u;
    if (gi >= p.I) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int gj = j0 + tx * 4 + v;
      if (gj >= p.J) continue;
      float val = acc[u][v];
      if (p.bias != nullptr) val += p.bias[gj];
      if (p.relu) val = fmaxf(val, 0.f);
      float* c = p.C + gi * p.sCi + gj;
      if (p.mode == 0) *c = val;
      else if (p.mode == 1) *c += val;
      else acc_add(p.C, (int64_t)gi * p.sCi + gj, val, p.det);
    }
    if (do_rowsum && tx == 0) acc_add(p.rowsum, gi, rsum[u], p.det);
  }
}

// ---------------------------------------------------------------------------------------------
// embedding  h0[r, 0:60) = posenc(x,10), [60:72) = posenc(t,6), [72:72+L) = latent
// order per src/pos_enc.py:27-36: for k: sin(2^k x_d) (all d), cos(2^k x_d) (all d)
// ---------------------------------------------------------------------------------------------
constexpr int PTS_FREQS = 10, TIME_FREQS = 6, EMB_XT = 3 * 2 * PTS_FREQS + 2 * TIME_FREQS;   // 72

__global__ void __launch_bounds__(128) embed_fwd_kernel(int G, int Mrows, int L, const float* __restrict__ pts,
                                                        const float* __restrict__ times,
                                                        const float* __restrict__ latents, float* __restrict__ h0,
                                                        int64_t ldh) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;   // one warp per row
  if (r >= (int64_t)G * Mrows) return;
  const int lane = threadIdx.x & 31;
  const int g = (int)(r / Mrows), m = (int)(r - (int64_t)g * Mrows);
  float* out = h0 + r * ldh;
  const float t = times[g];
  for (int c = lane; c < EMB_XT + L; c += 32) {
    float v;
    if (c < 60) {
      const int k = c / 6, rem = c - 6 * k, d = rem % 3;
      const float a = pts[3 * m + d] * (float)(1 << k);
      v = rem < 3 ? sinf(a) : cosf(a);
    } else if (c < EMB_XT) {
      const int cc = c - 60, k = cc >> 1;
      const float a = t * (float)(1 << k);
      v = (cc & 1) ? cosf(a) : sinf(a);
    } else {
      v = latents[(int64_t)g * L + (c - EMB_XT)];
    }
    out[c] = v;
  }
}

// dpts[m,d] += sum_g sum_k 2^k (cos(2^k x) dh[6k+d] - sin(2^k x) dh[6k+3+d]) ; dlatents[g,l] += sum_m dh[72+l].
// One warp per row, EMB_ROWS_PER_WARP rows per warp: lanes read the row's 104 gradient values and the sin/cos
// values the forward pass left in h0 (no transcendental is recomputed), fully coalesced.
constexpr int EMB_ROWS_PER_WARP = 4;

__global__ void __launch_bounds__(256) embed_bwd_kernel(int G, int Mrows, int L, const float* __restrict__ h0,
                                                        const float* __restrict__ dh0, int64_t ldh,
                                                        float* __restrict__ dpts, float* __restrict__ dlatents, float det) {
  extern __shared__ float slat[];   // [L] partial sums of this block
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int l = threadIdx.x; l < L; l += blockDim.x) slat[l] = 0.f;
  __syncthreads();
  const int row_base = (blockIdx.x * (blockDim.x >> 5) + warp) * EMB_ROWS_PER_WARP;
  float lat_acc[4] = {0.f, 0.f, 0.f, 0.f};     // latent columns lane, lane+32, ... (L <= 128)
  for (int rr = 0; rr < EMB_ROWS_PER_WARP; ++rr) {
    const int m = row_base + rr;
    if (m >= Mrows) break;
    const float* dh = dh0 + ((int64_t)g * Mrows + m) * ldh;
    const float* hv = h0 + ((int64_t)g * Mrows + m) * ldh;
    if (dpts != nullptr) {
      float gp[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = lane + 32 * half;
        if (c < 60) {
          const int k = c / 6, rem = c - 6 * k, d = rem % 3;
          const float f = (float)(1 << k);
          // column c holds sin(f x_d) if rem < 3 (its derivative uses the cos stored 3 columns later), else cos
          const float term = rem < 3 ? f * hv[c + 3] * dh[c] : -f * hv[c - 3] * dh[c];
          gp[0] += d == 0 ? term : 0.f;
          gp[1] += d == 1 ? term : 0.f;
          gp[2] += d == 2 ? term : 0.f;
        }
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gp[d] += __shfl_xor_sync(0xffffffffu, gp[d], o);
      }
      if (lane < 3) acc_add(dpts, 3 * (int64_t)m + lane, lane == 0 ? gp[0] : (lane == 1 ? gp[1] : gp[2]), det);
    }
    if (dlatents != nullptr) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int l = lane + 32 * q;
        if (l < L) lat_acc[q] += dh[EMB_XT + l];
      }
    }
  }
  if (dlatents != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int l = lane + 32 * q;
      if (l < L) {
        if (det != 0.f) acc_add(dlatents, (int64_t)g * L + l, lat_acc[q], det);     // order-independent: no staging
        else atomicAdd(&slat[l], lat_acc[q]);
      }
    }
  }
  __syncthreads();
  if (dlatents != nullptr && det == 0.f)
    for (int l = threadIdx.x; l < L; l += blockDim.x) atomicAdd(&dlatents[(int64_t)g * L + l], slat[l]);
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_linear_fwd(int R, int K, int No, const float* X, int64_t ldx, const float* Wt,
                               const float* bias, float* Y, int64_t ldy, int relu, void* stream) {
  if (R == 0) return 0;
  GemmArgs p{};
  p.I = R; p.J = No; p.L = K;
  p.A = X; p.sAi = ldx; p.sAl = 1;
  p.Bm = Wt; p.sBl = 1; p.sBj = K;
  p.C = Y; p.sCi = ldy; p.bias = bias; p.relu = relu; p.mode = 0; p.l_per_split = K;
  dim3 grid(ceil_div(R, BM), ceil_div(No, BN), 1);
  gemm_kernel<true, true><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_linear_bwd_data(int R, int K, int No, const float* dY, int64_t lddy, const float* Y,
                                    int64_t ldy, const float* Wt, float* dX, int64_t lddx, int accumulate,
                                    void* stream) {
  if (R == 0) return 0;
  GemmArgs p{};
  p.I = R; p.J = K; p.L = No;
  p.A = dY; p.sAi = lddy; p.sAl = 1;
  p.mask = Y; p.sMi = ldy; p.sMl = 1;
  p.Bm = Wt; p.sBl = K; p.sBj = 1;
  p.C = dX; p.sCi = lddx; p.mode = accumulate ? 1 : 0; p.l_per_split = No;
  dim3 grid(ceil_div(R, BM), ceil_div(K, BN), 1);
  gemm_kernel<true, false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_linear_bwd_weight(int R, int K, int No, const float* dY, int64_t lddy, const float* Y,
                                      int64_t ldy, const float* X, int64_t ldx, float* dW, float* db,
                                      void* stream) {
  if (R == 0) return 0;
  GemmArgs p{};
  p.I = No; p.J = K; p.L = R;
  p.A = dY; p.sAi = 1; p.sAl = lddy;
  p.mask = Y; p.sMi = 1; p.sMl = ldy;
  p.Bm = X; p.sBl = ldx; p.sBj = 1;
  p.C = dW; p.sCi = K; p.rowsum = db; p.mode = 2; p.det = dimo::det_scale();
  // split the long reduction over rows so the grid fills the machine (~2 waves of 148 SMs)
  const int tiles = ceil_div(No, BM) * ceil_div(K, BN);
  int splits = max(1, min(ceil_div(R, 4 * BK), ceil_div(296, tiles)));
  p.l_per_split = ceil_div(ceil_div(R, splits), BK) * BK;
  splits = ceil_div(R, p.l_per_split);
  dim3 grid(ceil_div(No, BM), ceil_div(K, BN), splits);
  gemm_kernel<false, false><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_timenet_embed_fwd(int G, int rows_per_group, int L, const float* pts, const float* times,
                                      const float* latents, float* h0, int64_t ldh, void* stream) {
  const int64_t rows = (int64_t)G * rows_per_group;
  if (rows == 0) return 0;
  embed_fwd_kernel<<<ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>(G, rows_per_group, L, pts, times, latents, h0,
                                                                       ldh);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_timenet_embed_bwd(int G, int rows_per_group, int L, const float* h0, const float* dh0,
                                      int64_t ldh, float* dpts, float* dlatents, void* stream) {
  if (G == 0 || rows_per_group == 0) return 0;
  DIMO_REQUIRE(L <= 128, "latent dimension must be <= 128");
  dim3 grid(ceil_div(rows_per_group, 8 * EMB_ROWS_PER_WARP), G);
  embed_bwd_kernel<<<grid, 256, sizeof(float) * (size_t)max(L, 1), (cudaStream_t)stream>>>(
      G, rows_per_group, L, h0, dh0, ldh, dpts, dlatents, dimo::det_scale());
  DIMO_CHECK_LAUNCH();
  return 0;
}
