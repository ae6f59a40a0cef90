// Fused image loss: SSIM (11x11 Gaussian window, sigma 1.5, zero padding, C1=0.01^2, C2=0.03^2) + L1 + MSE
// partial sums, forward and backward.  Replaces src/loss.py:144-178 (ssim/_ssim: five cuDNN depthwise
// convs + ~15 elementwise launches, l1_loss), F.mse_loss (main_train_dimo.py:333) and
// fused_ssim.fused_ssim (main_test_dimo.py:979).
//
// One CTA = one 16x16 output tile of one (batch, channel) plane.  The 26x26 halo of both images is
// staged in shared memory, the five moments are filtered separably (horizontal into smem, vertical
// into registers), the SSIM map and its three partial derivatives are formed per pixel, and the
// plane sums are reduced with warp shuffles -> one atomicAdd per CTA.  HBM-bound:
// fwd 8 B/px read + 12 B/px written (dm maps), bwd 20 B/px read + 4 B/px written.
#include "common.cuh"

namespace dimo {

constexpr int SS_T = 16;               // tile edge
constexpr int SS_R = 5;                // window radius (11 taps)
constexpr int SS_H = SS_T + 2 * SS_R;  // 26
constexpr float SSIM_C1 = 0.0001f, SSIM_C2 = 0.0009f;

// g[x] = exp(-(x-5)^2 / (2*1.5^2)) / sum, computed as src/loss.py:132-134 does (fp32 tensor, fp32 sum)
__device__ __constant__ float SS_W[11] = {0.0010283801f, 0.0075987582f, 0.0360007733f, 0.1093606874f,
                                          0.2130055279f, 0.2660117149f, 0.2130055279f, 0.1093606874f,
                                          0.0360007733f, 0.0075987582f, 0.0010283801f};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int tid = threadIdx.y * SS_T + threadIdx.x;
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (tid < 8) s = red[tid];
  if (tid < 32) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  __syncthreads();
  return s;   // valid in thread 0
}

__global__ void __launch_bounds__(256) ssim_fwd_kernel(int H, int W, const float* __restrict__ img1,
                                                       const float* __restrict__ img2, float* __restrict__ sums,
                                                       float* __restrict__ dm, int64_t plane_count, int clamp01) {
  __shared__ float s1[SS_H][SS_H + 1];
  __shared__ float s2[SS_H][SS_H + 1];
  __shared__ float hz[5][SS_H][SS_T + 1];
  __shared__ float red[8];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
  const int tid = threadIdx.y * SS_T + threadIdx.x;
  const int64_t hw = (int64_t)H * W;
  const float* p1 = img1 + plane * hw;
  const float* p2 = img2 + plane * hw;

  for (int e = tid; e < SS_H * SS_H; e += 256) {
    const int ly = e / SS_H, lx = e - ly * SS_H;
    const int gy = y0 + ly - SS_R, gx = x0 + lx - SS_R;
    float a = 0.f, b = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) { a = p1[(int64_t)gy * W + gx]; b = p2[(int64_t)gy * W + gx]; }
    if (clamp01) a = fminf(fmaxf(a, 0.f), 1.f);
    s1[ly][lx] = a; s2[ly][lx] = b;
  }
  __syncthreads();
  // horizontal pass: 26 rows x 16 cols
  for (int e = tid; e < SS_H * SS_T; e += 256) {
    const int ly = e / SS_T, lx = e - ly * SS_T;
    float m1 = 0.f, m2 = 0.f, q1 = 0.f, q2 = 0.f, q12 = 0.f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = SS_W[k], a = s1[ly][lx + k], b = s2[ly][lx + k];
      m1 += w * a; m2 += w * b; q1 += w * a * a; q2 += w * b * b; q12 += w * a * b;
    }
    hz[0][ly][lx] = m1; hz[1][ly][lx] = m2; hz[2][ly][lx] = q1; hz[3][ly][lx] = q2; hz[4][ly][lx] = q12;
  }
  __syncthreads();
  const int lx = threadIdx.x, ly = threadIdx.y;
  const int gx = x0 + lx, gy = y0 + ly;
  const bool inside = gx < W && gy < H;
  float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
  for (int k = 0; k < 11; ++k) {
    const float w = SS_W[k];
    mu1 += w * hz[0][ly + k][lx]; mu2 += w * hz[1][ly + k][lx];
    e11 += w * hz[2][ly + k][lx]; e22 += w * hz[3][ly + k][lx]; e12 += w * hz[4][ly + k][lx];
  }
  float v_ssim = 0.f, v_l1 = 0.f, v_mse = 0.f;
  if (inside) {
    const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
    const float sg1 = e11 - mu1s, sg2 = e22 - mu2s, sg12 = e12 - mu12;
    const float num1 = 2.f * mu12 + SSIM_C1, num2 = 2.f * sg12 + SSIM_C2;
    const float den1 = mu1s + mu2s + SSIM_C1, den2 = sg1 + sg2 + SSIM_C2;
    const float inv = 1.f / (den1 * den2);
    v_ssim = num1 * num2 * inv;
    const float a = s1[ly + SS_R][lx + SS_R], b = s2[ly + SS_R][lx + SS_R];
    v_l1 = fabsf(a - b);
    v_mse = (a - b) * (a - b);
    if (dm != nullptr) {
      // partial derivatives of the map w.r.t. sigma1^2, sigma12 and (total) mu1
      const float d_sg1 = -num1 * num2 * inv / den2;
      const float d_sg12 = 2.f * num1 * inv;
      const float d_mu1 = 2.f * mu2 * num2 * inv - 2.f * mu1 * num1 * num2 * inv / den1 - 2.f * mu1 * d_sg1 -
                          mu2 * d_sg12;
      const int64_t o = plane * hw + (int64_t)gy * W + gx;
      dm[o] = d_mu1;
      dm[plane_count * hw + o] = d_sg1;
      dm[2 * plane_count * hw + o] = d_sg12;
    }
  }
  const float t0 = block_sum_256(v_ssim, red);
  const float t1 = block_sum_256(v_l1, red);
  const float t2 = block_sum_256(v_mse, red);
  if (tid == 0) { atomicAdd(&sums[0], t0); atomicAdd(&sums[1], t1); atomicAdd(&sums[2], t2); }
}

__global__ void __launch_bounds__(256) ssim_bwd_kernel(int H, int W, const float* __restrict__ img1,
                                                       const float* __restrict__ img2, const float* __restrict__ dm,
                                                       float w_ssim, float w_l1, float w_mse,
                                                       float* __restrict__ dL_dimg1, int64_t plane_count,
                                                       int clamp01) {
  __shared__ float sm[3][SS_H][SS_H + 1];
  __shared__ float hz[3][SS_H][SS_T + 1];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
  const int tid = threadIdx.y * SS_T + threadIdx.x;
  const int64_t hw = (int64_t)H * W;
  const int lx = threadIdx.x, ly = threadIdx.y;
  const int gx = x0 + lx, gy = y0 + ly;
  const bool inside = gx < W && gy < H;
  float conv[3] = {0.f, 0.f, 0.f};
  if (w_ssim != 0.f && dm != nullptr) {
    for (int e = tid; e < SS_H * SS_H; e += 256) {
      const int yy = e / SS_H, xx = e - yy * SS_H;
      const int sy = y0 + yy - SS_R, sx = x0 + xx - SS_R;
      const bool ok = sy >= 0 && sy < H && sx >= 0 && sx < W;
      const int64_t o = plane * hw + (int64_t)sy * W + sx;
#pragma unroll
      for (int c = 0; c < 3; ++c) sm[c][yy][xx] = ok ? dm[c * plane_count * hw + o] : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < SS_H * SS_T; e += 256) {
      const int yy = e / SS_T, xx = e - yy * SS_T;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const float w = SS_W[k];
        a0 += w * sm[0][yy][xx + k]; a1 += w * sm[1][yy][xx + k]; a2 += w * sm[2][yy][xx + k];
      }
      hz[0][yy][xx] = a0; hz[1][yy][xx] = a1; hz[2][yy][xx] = a2;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const float w = SS_W[k];
      conv[0] += w * hz[0][ly + k][lx]; conv[1] += w * hz[1][ly + k][lx]; conv[2] += w * hz[2][ly + k][lx];
    }
  }
  if (inside) {
    const int64_t o = plane * hw + (int64_t)gy * W + gx;
    const float araw = img1[o], b = img2[o];
    const float a = clamp01 ? fminf(fmaxf(araw, 0.f), 1.f) : araw;
    float g = w_ssim * (conv[0] + 2.f * a * conv[1] + b * conv[2]);
    const float d = a - b;
    g += w_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    g += w_mse * 2.f * d;
    if (clamp01 && (araw < 0.f || araw > 1.f)) g = 0.f;   // torch.clamp backward: pass-through on [0,1] inclusive
    dL_dimg1[o] = g;
  }
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_ssim_fwd(int B, int C, int H, int W, int clamp01, const float* img1, const float* img2,
                             float* sums, float* dm, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DIMO_CHECK_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(float), st));
  const int planes = B * C;
  if (planes == 0) return 0;
  DIMO_REQUIRE(planes <= 65535, "B*C must be <= 65535");
  dim3 grid(ceil_div(W, SS_T), ceil_div(H, SS_T), planes), block(SS_T, SS_T);
  ssim_fwd_kernel<<<grid, block, 0, st>>>(H, W, img1, img2, sums, dm, (int64_t)planes, clamp01);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_ssim_bwd(int B, int C, int H, int W, int clamp01, const float* img1, const float* img2,
                             const float* dm, float w_ssim, float w_l1, float w_mse, float* dL_dimg1,
                             void* stream) {
  const int planes = B * C;
  if (planes == 0) return 0;
  DIMO_REQUIRE(planes <= 65535, "B*C must be <= 65535");
  DIMO_REQUIRE(w_ssim == 0.f || dm != nullptr, "dm maps required when w_ssim != 0");
  dim3 grid(ceil_div(W, SS_T), ceil_div(H, SS_T), planes), block(SS_T, SS_T);
  ssim_bwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(H, W, img1, img2, dm, w_ssim, w_l1, w_mse, dL_dimg1,
                                                           (int64_t)planes, clamp01);
  DIMO_CHECK_LAUNCH();
  return 0;
}
