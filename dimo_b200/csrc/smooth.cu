// Step regularisers on the rendered maps: edge-aware depth smoothness and bilateral normal
// smoothness, src/loss.py:64-107, added to the loss at main_train_dimo.py:363-372 once the step counter passes
// depth_reg_start_iter / normal_reg_start_iter.  Upstream: ~30 elementwise launches and as many full-size temporaries
// per motion, forward and again in autograd's backward.
//
// Here: for every horizontally / vertically adjacent pixel pair (p, q)
//   gi = mean_c |rgb_p - rgb_q|,  e = exp(-gi)
//   depth term  |d_p - d_q| e            normal term  sum_c sqrt(1 + (|n_p - n_q| e^3)^2)
// The forward kernel only reduces the four sums (x / y pairs of both terms; their normalisers differ) and adds the
// weighted total to the step's loss scalar.  The backward kernel recomputes the four pair terms that touch a pixel and
// writes that pixel's gradients w.r.t. depth, normal and rgb (the exp(-gi) factor depends on the RENDERED image) -- one
// thread per pixel, no atomics, gradients leave as final values.  Neighbour reads hit L1/L2.
// HBM: fwd 28 B/px read; bwd 28 B/px read + 28 B/px written (+ 12 B/px when the rgb gradient is accumulated in place).
#include "common.cuh"

namespace dimo {

struct SmoothPx {
  float c[3], d, n[3];
};

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ SmoothPx smooth_load(const float* __restrict__ rgb, const float* __restrict__ depth,
                                                const float* __restrict__ normal, int64_t b, int64_t hw, int64_t pix,
                                                int clamp01) {
  SmoothPx o;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float v = rgb[(b * 3 + k) * hw + pix];
    if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
    o.c[k] = v;
    o.n[k] = normal[(b * 3 + k) * hw + pix];
  }
  o.d = depth[b * hw + pix];
  return o;
}

// One pair (P, Q): the two loss terms and, optionally, the gradients w.r.t. the P endpoint scaled by (wd, wn); the
// gradients w.r.t. Q are their negatives.
template <bool GRAD>
__device__ __forceinline__ void smooth_pair(const SmoothPx& P, const SmoothPx& Q, float wd, float wn, float& term_d,
                                            float& term_n, float& g_d, float (&g_n)[3], float (&g_c)[3]) {
  const float a0 = P.c[0] - Q.c[0], a1 = P.c[1] - Q.c[1], a2 = P.c[2] - Q.c[2];
  const float gi = (fabsf(a0) + fabsf(a1) + fabsf(a2)) * (1.0f / 3.0f);
  const float e1 = expf(-gi), e3 = e1 * e1 * e1;
  const float dd = P.d - Q.d;
  term_d = fabsf(dd) * e1;
  term_n = 0.f;
  float dgi = -wd * term_d;
  if (GRAD) g_d = wd * sgnf(dd) * e1;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float diff = P.n[k] - Q.n[k];
    const float u = fabsf(diff) * e3;
    const float r = sqrtf(1.0f + u * u);
    term_n += r;
    if (GRAD) {
      const float kk = u / r;
      g_n[k] = wn * kk * e3 * sgnf(diff);
      dgi -= 3.0f * wn * kk * u;
    }
  }
  if (GRAD) {
    const float t = dgi * (1.0f / 3.0f);
    g_c[0] = t * sgnf(a0); g_c[1] = t * sgnf(a1); g_c[2] = t * sgnf(a2);
  }
}

__global__ void __launch_bounds__(256) smooth_fwd_kernel(int H, int W, int clamp01, const float* __restrict__ rgb,
                                                         const float* __restrict__ depth,
                                                         const float* __restrict__ normal, float* __restrict__ sums,
                                                         float* __restrict__ loss_acc, float wdx, float wdy, float wnx,
                                                         float wny) {
  __shared__ float red[4][8];
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  const int64_t b = blockIdx.z, hw = (int64_t)H * W;
  float s[4] = {0.f, 0.f, 0.f, 0.f};   // depth x, depth y, normal x, normal y
  if (x < W && y < H) {
    const int64_t pix = (int64_t)y * W + x;
    const SmoothPx P = smooth_load(rgb, depth, normal, b, hw, pix, clamp01);
    float gd, gn[3], gc[3];
    if (x + 1 < W) {
      const SmoothPx Q = smooth_load(rgb, depth, normal, b, hw, pix + 1, clamp01);
      smooth_pair<false>(P, Q, 0.f, 0.f, s[0], s[2], gd, gn, gc);
    }
    if (y + 1 < H) {
      const SmoothPx Q = smooth_load(rgb, depth, normal, b, hw, pix + W, clamp01);
      smooth_pair<false>(P, Q, 0.f, 0.f, s[1], s[3], gd, gn, gc);
    }
  }
  const int lane = threadIdx.x, warp = threadIdx.y;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v = s[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0 && lane < 4) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[lane][w];
    atomicAdd(&sums[lane], t);
    if (loss_acc != nullptr) {
      const float wk = lane == 0 ? wdx : lane == 1 ? wdy : lane == 2 ? wnx : wny;
      atomicAdd(loss_acc, wk * t);
    }
  }
}

__global__ void __launch_bounds__(256) smooth_bwd_kernel(int H, int W, int clamp01, const float* __restrict__ rgb,
                                                         const float* __restrict__ depth,
                                                         const float* __restrict__ normal, float wdx, float wdy,
                                                         float wnx, float wny, const float* __restrict__ g_dev,
                                                         float* __restrict__ d_rgb, int accumulate_rgb,
                                                         float* __restrict__ d_depth, float* __restrict__ d_normal) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= W || y >= H) return;
  const int64_t b = blockIdx.z, hw = (int64_t)H * W;
  const int64_t pix = (int64_t)y * W + x;
  const float gup = g_dev != nullptr ? g_dev[0] : 1.f;
  const SmoothPx P = smooth_load(rgb, depth, normal, b, hw, pix, clamp01);
  float acc_d = 0.f, acc_n[3] = {0.f, 0.f, 0.f}, acc_c[3] = {0.f, 0.f, 0.f};
  float td, tn, gd, gn[3], gc[3];
  if (x + 1 < W) {      // this pixel is the P endpoint of its right pair
    const SmoothPx Q = smooth_load(rgb, depth, normal, b, hw, pix + 1, clamp01);
    smooth_pair<true>(P, Q, wdx, wnx, td, tn, gd, gn, gc);
    acc_d += gd;
#pragma unroll
    for (int k = 0; k < 3; ++k) { acc_n[k] += gn[k]; acc_c[k] += gc[k]; }
  }
  if (x > 0) {          // ... and the Q endpoint of its left neighbour's pair
    const SmoothPx L = smooth_load(rgb, depth, normal, b, hw, pix - 1, clamp01);
    smooth_pair<true>(L, P, wdx, wnx, td, tn, gd, gn, gc);
    acc_d -= gd;
#pragma unroll
    for (int k = 0; k < 3; ++k) { acc_n[k] -= gn[k]; acc_c[k] -= gc[k]; }
  }
  if (y + 1 < H) {
    const SmoothPx Q = smooth_load(rgb, depth, normal, b, hw, pix + W, clamp01);
    smooth_pair<true>(P, Q, wdy, wny, td, tn, gd, gn, gc);
    acc_d += gd;
#pragma unroll
    for (int k = 0; k < 3; ++k) { acc_n[k] += gn[k]; acc_c[k] += gc[k]; }
  }
  if (y > 0) {
    const SmoothPx U = smooth_load(rgb, depth, normal, b, hw, pix - W, clamp01);
    smooth_pair<true>(U, P, wdy, wny, td, tn, gd, gn, gc);
    acc_d -= gd;
#pragma unroll
    for (int k = 0; k < 3; ++k) { acc_n[k] -= gn[k]; acc_c[k] -= gc[k]; }
  }
  d_depth[b * hw + pix] = gup * acc_d;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int64_t o = (b * 3 + k) * hw + pix;
    d_normal[o] = gup * acc_n[k];
    float g = gup * acc_c[k];
    if (clamp01) {      // torch.clamp backward: pass-through on [0,1] inclusive
      const float raw = rgb[o];
      if (raw < 0.f || raw > 1.f) g = 0.f;
    }
    d_rgb[o] = accumulate_rgb ? d_rgb[o] + g : g;
  }
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_smooth_fwd(int B, int H, int W, int clamp01, const float* rgb, const float* depth,
                               const float* normal, float* sums4, float* loss_acc, float wdx, float wdy, float wnx,
                               float wny, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DIMO_CHECK_CUDA(cudaMemsetAsync(sums4, 0, 4 * sizeof(float), st));
  if (B == 0 || H == 0 || W == 0) return 0;
  DIMO_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "smooth: bad sizes (B <= 65535)");
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), B), block(32, 8);
  smooth_fwd_kernel<<<grid, block, 0, st>>>(H, W, clamp01, rgb, depth, normal, sums4, loss_acc, wdx, wdy, wnx, wny);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_smooth_bwd(int B, int H, int W, int clamp01, const float* rgb, const float* depth,
                               const float* normal, float wdx, float wdy, float wnx, float wny, const float* g_dev,
                               float* d_rgb, int accumulate_rgb, float* d_depth, float* d_normal, void* stream) {
  if (B == 0 || H == 0 || W == 0) return 0;
  DIMO_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "smooth: bad sizes (B <= 65535)");
  dim3 grid(ceil_div(W, 32), ceil_div(H, 8), B), block(32, 8);
  smooth_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(H, W, clamp01, rgb, depth, normal, wdx, wdy, wnx, wny,
                                                             g_dev, d_rgb, accumulate_rgb, d_depth, d_normal);
  DIMO_CHECK_LAUNCH();
  return 0;
}
