// Fused image loss: SSIM (11x11 Gaussian window, sigma 1.5, zero padding, C1=0.01^2, C2=0.03^2) + L1 + MSE
// partial sums, forward and backward.  Replaces src/loss.py:144-178 (ssim/_ssim: five cuDNN depthwise
// convs + ~15 elementwise launches, l1_loss), F.mse_loss (main_train_dimo.py:333) and
// fused_ssim.fused_ssim (main_test_dimo.py:979).
//
// One CTA (256 threads) = one 64x32 output tile of one (batch, channel) plane, two phases:
//   1. horizontal: item = (halo row, group of 4 columns).  The 14-wide input window of both images comes straight from
//      global memory (three aligned 128-bit loads + two scalars per image, neighbours overlap in L1 -- no staging pass,
//      no per-element index arithmetic), the five moments are filtered with packed FFMA2 taps on the pairs
//      (a, b), (a^2, b^2) + scalar ab, and the 42 x 64 filtered rows go to shared memory;
//   2. vertical: thread = (column, group of 4 rows) filters a 14-row register window and forms the SSIM map, its
//      three partial derivatives and the L1 / MSE terms per pixel; plane sums by warp shuffles -> one atomicAdd
//      triple per CTA.
// ~150 issue slots per pixel (the first version staged the halo element-wise and spent 40 % of its instructions on
// integer address arithmetic: ncu profiles/r2r).  HBM: fwd 8 B/px read + 12 B/px written (dm maps), bwd 20 B/px read +
// 4 B/px written.
#include "common.cuh"

namespace dimo {

constexpr int SS_TW = 64, SS_TH = 32;      // output tile
constexpr int SS_R = 5;                    // window radius (11 taps)
constexpr int SS_HH = SS_TH + 2 * SS_R;    // 42 halo rows
constexpr int SS_THREADS = 256;
constexpr int SS_G = SS_TW / 4;            // 16 groups of 4 outputs per row
constexpr float SSIM_C1 = 0.0001f, SSIM_C2 = 0.0009f;
constexpr int SS_FWD_SMEM = SS_HH * SS_TW * 20;     // (mu1, mu2), (E[a^2], E[b^2]) pairs + E[ab]: 53.75 KB

// g[x] = exp(-(x-5)^2 / (2*1.5^2)) / sum, computed as src/loss.py:132-134 does (fp32 tensor, fp32 sum)
__device__ __constant__ float SS_W[11] = {0.0010283801f, 0.0075987582f, 0.0360007733f, 0.1093606874f,
                                          0.2130055279f, 0.2660117149f, 0.2130055279f, 0.1093606874f,
                                          0.0360007733f, 0.0075987582f, 0.0010283801f};

template <int NW>
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int tid = threadIdx.x;
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < NW; ++w) s += red[w];
  __syncthreads();
  return s;
}

// columns c .. c + 13 of one image row (zero outside [0, W) and when the row itself is outside the image).  c + 1 is a
// multiple of 4; VEC: rows are 16-byte aligned and W % 4 == 0, so columns c + 1 .. c + 12 are three aligned float4.
template <bool VEC>
__device__ __forceinline__ void ss_load14(const float* __restrict__ row, bool rowok, int c, int W, float (&v)[14]) {
  if (VEC) {
    v[0] = (rowok && c >= 0 && c < W) ? row[c] : 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int cc = c + 1 + 4 * i;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rowok && cc >= 0 && cc < W) t = *reinterpret_cast<const float4*>(row + cc);
      v[1 + 4 * i] = t.x; v[2 + 4 * i] = t.y; v[3 + 4 * i] = t.z; v[4 + 4 * i] = t.w;
    }
    v[13] = (rowok && c + 13 >= 0 && c + 13 < W) ? row[c + 13] : 0.f;
  } else {
#pragma unroll
    for (int k = 0; k < 14; ++k) v[k] = (rowok && c + k >= 0 && c + k < W) ? row[c + k] : 0.f;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(SS_THREADS, 4) ssim_fwd_kernel(int H, int W, const float* __restrict__ img1,
                                                                 const float* __restrict__ img2,
                                                                 float* __restrict__ sums, float* __restrict__ dm,
                                                                 int64_t plane_count, int clamp01, int C,
                                                                 const float* __restrict__ mse_frame_w,
                                                                 float* __restrict__ loss_acc, float lw_ssim, float lw_l1,
                                                                 float lw_mse) {
  extern __shared__ __align__(16) uint8_t ss_smem[];
  float2(*hzm)[SS_TW] = reinterpret_cast<float2(*)[SS_TW]>(ss_smem);
  float2(*hzq)[SS_TW] = reinterpret_cast<float2(*)[SS_TW]>(ss_smem + SS_HH * SS_TW * 8);
  float(*hzx)[SS_TW] = reinterpret_cast<float(*)[SS_TW]>(ss_smem + SS_HH * SS_TW * 16);
  __shared__ float red[SS_THREADS / 32];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * SS_TW, y0 = blockIdx.y * SS_TH;
  const int tid = threadIdx.x;
  const int64_t hw = (int64_t)H * W;
  const float* p1 = img1 + plane * hw;
  const float* p2 = img2 + plane * hw;

  // ---- phase 1: horizontal pass, item = (halo row, group of 4 columns); a warp = 2 rows x 16 groups ----
  for (int e = tid; e < SS_HH * SS_G; e += SS_THREADS) {
    const int row = e >> 4, g = e & (SS_G - 1);
    const int gy = y0 + row - SS_R;
    const bool rowok = gy >= 0 && gy < H;
    const int64_t ro = (int64_t)(rowok ? gy : 0) * W;
    float a[14], b[14];
    ss_load14<VEC>(p1 + ro, rowok, x0 + 4 * g - SS_R, W, a);
    ss_load14<VEC>(p2 + ro, rowok, x0 + 4 * g - SS_R, W, b);
    // (a, b) come from different loads: pairing them would cost two MOVs per tap input, so the means stay scalar; the
    // squares are computed values and land in adjacent registers for free -> packed FFMA2
    f2 sq[14];
    float x[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      if (clamp01) a[k] = fminf(fmaxf(a[k], 0.f), 1.f);
      sq[k] = f2{a[k] * a[k], b[k] * b[k]};
      x[k] = a[k] * b[k];
    }
    float2 om[4], oq[4];
    float ox[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f2 q = f2_bcast(0.f);
      float m1 = 0.f, m2 = 0.f, q12 = 0.f;
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        m1 = fmaf(SS_W[k], a[j + k], m1); m2 = fmaf(SS_W[k], b[j + k], m2);
        q = fma2(f2_bcast(SS_W[k]), sq[j + k], q); q12 = fmaf(SS_W[k], x[j + k], q12);
      }
      om[j] = make_float2(m1, m2); oq[j] = make_float2(q.x, q.y); ox[j] = q12;
    }
    float4* dmv = reinterpret_cast<float4*>(&hzm[row][4 * g]);
    float4* dqv = reinterpret_cast<float4*>(&hzq[row][4 * g]);
    dmv[0] = make_float4(om[0].x, om[0].y, om[1].x, om[1].y); dmv[1] = make_float4(om[2].x, om[2].y, om[3].x, om[3].y);
    dqv[0] = make_float4(oq[0].x, oq[0].y, oq[1].x, oq[1].y); dqv[1] = make_float4(oq[2].x, oq[2].y, oq[3].x, oq[3].y);
    *reinterpret_cast<float4*>(&hzx[row][4 * g]) = make_float4(ox[0], ox[1], ox[2], ox[3]);
  }
  __syncthreads();
  // ---- phase 2: vertical pass, thread = (column, group of 4 rows) ----
  const int col = tid & (SS_TW - 1);
  const int gx = x0 + col;
  float v_ssim = 0.f, v_l1 = 0.f, v_mse = 0.f;
  for (int rg = tid >> 6; rg < SS_TH / 4; rg += SS_THREADS / SS_TW) {
    float mom[5][4];
    {
      f2 wm[14], wq[14];
      float wx[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) {
        const float2 m = hzm[4 * rg + k][col], q = hzq[4 * rg + k][col];
        wm[k] = f2{m.x, m.y}; wq[k] = f2{q.x, q.y}; wx[k] = hzx[4 * rg + k][col];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f2 m = f2_bcast(0.f), q = f2_bcast(0.f);
        float q12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const f2 w = f2_bcast(SS_W[k]);
          m = fma2(w, wm[j + k], m); q = fma2(w, wq[j + k], q); q12 = fmaf(SS_W[k], wx[j + k], q12);
        }
        mom[0][j] = m.x; mom[1][j] = m.y; mom[2][j] = q.x; mom[3][j] = q.y; mom[4][j] = q12;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gy = y0 + 4 * rg + j;
      if (gx < W && gy < H) {
        const float mu1 = mom[0][j], mu2 = mom[1][j], e11 = mom[2][j], e22 = mom[3][j], e12 = mom[4][j];
        const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
        const float sg1 = e11 - mu1s, sg2 = e22 - mu2s, sg12 = e12 - mu12;
        const float num1 = 2.f * mu12 + SSIM_C1, num2 = 2.f * sg12 + SSIM_C2;
        const float den1 = mu1s + mu2s + SSIM_C1, den2 = sg1 + sg2 + SSIM_C2;
        const float r1 = 1.f / den1, r2 = 1.f / den2, inv = r1 * r2;       // two reciprocals instead of three divisions
        v_ssim += num1 * num2 * inv;
        const int64_t o = plane * hw + (int64_t)gy * W + gx;
        float a = img1[o];
        const float b = img2[o];
        if (clamp01) a = fminf(fmaxf(a, 0.f), 1.f);
        v_l1 += fabsf(a - b);
        v_mse += (a - b) * (a - b);
        if (dm != nullptr) {
          // partial derivatives of the map w.r.t. sigma1^2, sigma12 and (total) mu1
          const float d_sg1 = -num1 * num2 * inv * r2;
          const float d_sg12 = 2.f * num1 * inv;
          const float d_mu1 = 2.f * mu2 * num2 * inv - 2.f * mu1 * num1 * num2 * inv * r1 - 2.f * mu1 * d_sg1 -
                              mu2 * d_sg12;
          dm[o] = d_mu1;
          dm[plane_count * hw + o] = d_sg1;
          dm[2 * plane_count * hw + o] = d_sg12;
        }
      }
    }
  }
  const float t0 = block_sum<SS_THREADS / 32>(v_ssim, red);
  const float t1 = block_sum<SS_THREADS / 32>(v_l1, red);
  float t2 = block_sum<SS_THREADS / 32>(v_mse, red);
  if (tid == 0) {
    if (mse_frame_w != nullptr) t2 *= mse_frame_w[plane / C];
    atomicAdd(&sums[0], t0); atomicAdd(&sums[1], t1); atomicAdd(&sums[2], t2);
    if (loss_acc != nullptr) atomicAdd(loss_acc, lw_ssim * t0 + lw_l1 * t1 + lw_mse * t2);
  }
}

// sum (a-b)^2 over n floats: the mask term F.mse_loss(alpha, gt_mask) (main_train_dimo.py:350) needs no SSIM
// moments.  128-bit loads over the first 4*n4 elements (n4 = 0 when a pointer is not 16-byte aligned), scalar loads for
// the rest, one atomic per CTA.  HBM: 8 B per element.
__global__ void __launch_bounds__(256) sqdiff_sum_kernel(int64_t n, int64_t n4, const float* __restrict__ a,
                                                         const float* __restrict__ b, float* __restrict__ sum,
                                                         float* __restrict__ loss_acc, float lw) {
  __shared__ float red[8];
  float v = 0.f;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 x = a4[i], y = b4[i];
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    v += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float d = a[i] - b[i];
    v = fmaf(d, d, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(sum, t);
    if (loss_acc != nullptr) atomicAdd(loss_acc, lw * t);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(SS_THREADS, 4) ssim_bwd_kernel(int H, int W, const float* __restrict__ img1,
                                                                 const float* __restrict__ img2,
                                                                 const float* __restrict__ dm, float w_ssim, float w_l1,
                                                                 float w_mse, float* __restrict__ dL_dimg1,
                                                                 int64_t plane_count, int clamp01, int C,
                                                                 const float* __restrict__ mse_frame_w,
                                                                 const float* __restrict__ g_dev) {
  __shared__ __align__(16) float2 hz01[SS_HH][SS_TW];   // maps 0, 1 as pairs (packed FFMA2 taps), map 2 scalar
  __shared__ __align__(16) float hz2[SS_HH][SS_TW];
  const int plane = blockIdx.z;
  const int x0 = blockIdx.x * SS_TW, y0 = blockIdx.y * SS_TH;
  const int tid = threadIdx.x;
  const int64_t hw = (int64_t)H * W;
  const bool conv_on = w_ssim != 0.f && dm != nullptr;
  if (conv_on) {
    const float* d0 = dm + plane * hw;
    const float* d1 = d0 + plane_count * hw;
    const float* d2 = d1 + plane_count * hw;
    for (int e = tid; e < SS_HH * SS_G; e += SS_THREADS) {
      const int row = e >> 4, g = e & (SS_G - 1);
      const int gy = y0 + row - SS_R;
      const bool rowok = gy >= 0 && gy < H;
      const int64_t ro = (int64_t)(rowok ? gy : 0) * W;
      float a[14], b[14], c[14];
      ss_load14<VEC>(d0 + ro, rowok, x0 + 4 * g - SS_R, W, a);
      ss_load14<VEC>(d1 + ro, rowok, x0 + 4 * g - SS_R, W, b);
      ss_load14<VEC>(d2 + ro, rowok, x0 + 4 * g - SS_R, W, c);
      float2 o01[4];
      float o2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          acc0 = fmaf(SS_W[k], a[j + k], acc0); acc1 = fmaf(SS_W[k], b[j + k], acc1); acc2 = fmaf(SS_W[k], c[j + k], acc2);
        }
        o01[j] = make_float2(acc0, acc1); o2[j] = acc2;
      }
      float4* dv = reinterpret_cast<float4*>(&hz01[row][4 * g]);
      dv[0] = make_float4(o01[0].x, o01[0].y, o01[1].x, o01[1].y); dv[1] = make_float4(o01[2].x, o01[2].y, o01[3].x, o01[3].y);
      *reinterpret_cast<float4*>(&hz2[row][4 * g]) = make_float4(o2[0], o2[1], o2[2], o2[3]);
    }
    __syncthreads();
  }
  const int col = tid & (SS_TW - 1);
  const int gx = x0 + col;
  // upstream gradient of the scalar loss (device scalar, so no host read and no extra elementwise pass)
  const float gup = g_dev != nullptr ? g_dev[0] : 1.f;
  w_ssim *= gup; w_l1 *= gup;
  w_mse *= gup * (mse_frame_w != nullptr ? mse_frame_w[plane / C] : 1.f);
  for (int rg = tid >> 6; rg < SS_TH / 4; rg += SS_THREADS / SS_TW) {
    float conv[3][4];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) conv[c][j] = 0.f;
    if (conv_on) {
      f2 w01[14];
      float w2[14];
#pragma unroll
      for (int k = 0; k < 14; ++k) {
        const float2 v = hz01[4 * rg + k][col];
        w01[k] = f2{v.x, v.y}; w2[k] = hz2[4 * rg + k][col];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        f2 acc = f2_bcast(0.f);
        float acc2 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) { acc = fma2(f2_bcast(SS_W[k]), w01[j + k], acc); acc2 = fmaf(SS_W[k], w2[j + k], acc2); }
        conv[0][j] = acc.x; conv[1][j] = acc.y; conv[2][j] = acc2;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gy = y0 + 4 * rg + j;
      if (gx < W && gy < H) {
        const int64_t o = plane * hw + (int64_t)gy * W + gx;
        const float araw = img1[o], b = img2[o];
        const float a = clamp01 ? fminf(fmaxf(araw, 0.f), 1.f) : araw;
        float g = w_ssim * (conv[0][j] + 2.f * a * conv[1][j] + b * conv[2][j]);
        const float d = a - b;
        g += w_l1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
        g += w_mse * 2.f * d;
        if (clamp01 && (araw < 0.f || araw > 1.f)) g = 0.f;   // torch.clamp backward: pass-through on [0,1] inclusive
        dL_dimg1[o] = g;
      }
    }
  }
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_ssim_fwd(int B, int C, int H, int W, int clamp01, const float* img1, const float* img2,
                             float* sums, float* dm, const float* mse_frame_w, float* loss_acc, float lw_ssim,
                             float lw_l1, float lw_mse, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DIMO_CHECK_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(float), st));
  const int planes = B * C;
  if (planes == 0) return 0;
  DIMO_REQUIRE(planes <= 65535, "B*C must be <= 65535");
  dim3 grid(ceil_div(W, SS_TW), ceil_div(H, SS_TH), planes);
  static bool attr_done = false;
  if (!attr_done) {
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(ssim_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS_FWD_SMEM));
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(ssim_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SS_FWD_SMEM));
    attr_done = true;
  }
  // 128-bit row loads need 16-byte aligned rows: W % 4 == 0 and aligned plane bases (H * W * 4 is then a multiple of 16)
  const bool vec = W % 4 == 0 && ((uintptr_t)img1 & 15) == 0 && ((uintptr_t)img2 & 15) == 0;
  if (vec)
    ssim_fwd_kernel<true><<<grid, SS_THREADS, SS_FWD_SMEM, st>>>(H, W, img1, img2, sums, dm, (int64_t)planes, clamp01, C,
                                                                 mse_frame_w, loss_acc, lw_ssim, lw_l1, lw_mse);
  else
    ssim_fwd_kernel<false><<<grid, SS_THREADS, SS_FWD_SMEM, st>>>(H, W, img1, img2, sums, dm, (int64_t)planes, clamp01, C,
                                                                  mse_frame_w, loss_acc, lw_ssim, lw_l1, lw_mse);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_ssim_bwd(int B, int C, int H, int W, int clamp01, const float* img1, const float* img2,
                             const float* dm, float w_ssim, float w_l1, float w_mse, const float* mse_frame_w,
                             const float* g_dev, float* dL_dimg1, void* stream) {
  const int planes = B * C;
  if (planes == 0) return 0;
  DIMO_REQUIRE(planes <= 65535, "B*C must be <= 65535");
  DIMO_REQUIRE(w_ssim == 0.f || dm != nullptr, "dm maps required when w_ssim != 0");
  dim3 grid(ceil_div(W, SS_TW), ceil_div(H, SS_TH), planes);
  const bool vec = W % 4 == 0 && dm != nullptr && ((uintptr_t)dm & 15) == 0;
  if (vec)
    ssim_bwd_kernel<true><<<grid, SS_THREADS, 0, (cudaStream_t)stream>>>(H, W, img1, img2, dm, w_ssim, w_l1, w_mse, dL_dimg1,
                                                                         (int64_t)planes, clamp01, C, mse_frame_w, g_dev);
  else
    ssim_bwd_kernel<false><<<grid, SS_THREADS, 0, (cudaStream_t)stream>>>(H, W, img1, img2, dm, w_ssim, w_l1, w_mse, dL_dimg1,
                                                                          (int64_t)planes, clamp01, C, mse_frame_w, g_dev);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_sqdiff_sum(int64_t n, const float* a, const float* b, float* sum, float* loss_acc, float lw,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DIMO_CHECK_CUDA(cudaMemsetAsync(sum, 0, sizeof(float), st));
  if (n == 0) return 0;
  DIMO_REQUIRE(n > 0, "n must not be negative");
  const bool aligned = ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0;
  const int64_t n4 = aligned ? n / 4 : 0;
  const int64_t want = ((aligned ? n4 + 3 : n) + 255) / 256;
  const int grid = (int)(want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8);
  sqdiff_sum_kernel<<<grid, 256, 0, st>>>(n, n4, a, b, sum, loss_acc, lw);
  DIMO_CHECK_LAUNCH();
  return 0;
}
