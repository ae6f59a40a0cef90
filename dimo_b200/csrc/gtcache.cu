// Ground truth resident in HBM (data half).
//
// The reference keeps every ground-truth frame in HOST memory as fp32 ([views, frames, 1, 3, ref, ref] per motion,
// main_train_dimo.py:102-126), uploads each frame of a step over PCIe and bilinearly resamples it to the step's render
// resolution (128 -> 256 -> 512, :263, 283-284, 305-313) -- S uploads + 2 S interpolate launches per step.
// B200: 180 GB of HBM hold the whole training set as the 8-bit samples it was decoded from (c3: 128 motions x 32 frames
// x 9 views x 512^2 x RGBA = 38.7 GB; the reference's own float value is byte / 255, utils/load_utils.py:25, 70), so a
// step's ground truth is ONE gather + convert + resample launch over a list of frame slots:
//
//   dimo_gt_fetch   store [F,4,Hs,Ws] (u8 or f32: R,G,B,mask) + slot list [S] -> rgb [S,3,Ho,Wo], mask [S,1,Ho,Wo] fp32
//
// Resampling = torch's F.interpolate(mode="bilinear", align_corners=False) (aten upsample_bilinear2d): scale = in / out,
// src = max(scale * (dst + 0.5) - 0.5, 0), taps i0 = floor(src), i1 = min(i0 + 1, in - 1), weights (1 - l, l); equal
// sizes reduce to an exact copy.  HBM-bound: reads <= 4 taps x 4 B (or 1 B) per output sample (neighbouring threads
// share taps through L1/L2), writes 4 B per output sample.
#include "common.cuh"

namespace dimo {

template <typename T>
__device__ __forceinline__ float gt_load(const T* p);
template <>
__device__ __forceinline__ float gt_load<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float gt_load<uint8_t>(const uint8_t* p) { return (float)__ldg(p) / 255.0f; }

template <typename T>
__global__ void __launch_bounds__(256) gt_fetch_kernel(int S, int Hs, int Ws, int Ho, int Wo, float sy, float sx,
                                                       const T* __restrict__ store, const int32_t* __restrict__ slots,
                                                       float* __restrict__ rgb, float* __restrict__ mask) {
  const int64_t per_frame = (int64_t)4 * Ho * Wo;
  const int64_t total = (int64_t)S * per_frame;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(e % Wo);
    const int y = (int)((e / Wo) % Ho);
    const int c = (int)((e / ((int64_t)Wo * Ho)) % 4);
    const int s = (int)(e / per_frame);
    float fy = sy * ((float)y + 0.5f) - 0.5f;
    float fx = sx * ((float)x + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = min((int)fy, Hs - 1), x0 = min((int)fx, Ws - 1);
    const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const T* plane = store + ((int64_t)slots[s] * 4 + c) * Hs * Ws;
    const float p00 = gt_load(plane + (int64_t)y0 * Ws + x0), p01 = gt_load(plane + (int64_t)y0 * Ws + x1);
    const float p10 = gt_load(plane + (int64_t)y1 * Ws + x0), p11 = gt_load(plane + (int64_t)y1 * Ws + x1);
    const float v = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
    if (c < 3) rgb[((int64_t)s * 3 + c) * Ho * Wo + (int64_t)y * Wo + x] = v;
    else mask[(int64_t)s * Ho * Wo + (int64_t)y * Wo + x] = v;
  }
}

// Equal sizes (no resampling), 8-bit store: pure gather + convert.  One thread = 16 consecutive samples of one plane:
// one 128-bit load, four 128-bit stores (the generic kernel moves one sample per thread with 1-byte loads: 7.5 % of
// the HBM peak, profiles/r1q_widen_bench.jsonl).
__global__ void __launch_bounds__(256) gt_convert_u8_kernel(int S, int64_t plane16, const uint8_t* __restrict__ store,
                                                            const int32_t* __restrict__ slots, float* __restrict__ rgb,
                                                            float* __restrict__ mask) {
  const int64_t total = (int64_t)S * 4 * plane16;       // plane16 = H * W / 16
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = e % plane16;
    const int c = (int)((e / plane16) % 4);
    const int s = (int)(e / (4 * plane16));
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(store + ((int64_t)slots[s] * 4 + c) * plane16 * 16) + v);
    float* dst = c < 3 ? rgb + ((int64_t)s * 3 + c) * plane16 * 16 : mask + (int64_t)s * plane16 * 16;
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 o;
      o.x = (float)(w[k] & 0xFF) / 255.0f; o.y = (float)((w[k] >> 8) & 0xFF) / 255.0f;
      o.z = (float)((w[k] >> 16) & 0xFF) / 255.0f; o.w = (float)(w[k] >> 24) / 255.0f;
      reinterpret_cast<float4*>(dst)[v * 4 + k] = o;
    }
  }
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_gt_fetch(int S, int Hs, int Ws, int Ho, int Wo, int store_is_u8, const void* store,
                             const int32_t* slots, float* rgb, float* mask, void* stream) {
  DIMO_REQUIRE(Hs > 0 && Ws > 0 && Ho > 0 && Wo > 0, "empty image");
  if (S == 0) return 0;
  const float sy = (float)Hs / (float)Ho, sx = (float)Ws / (float)Wo;
  const int64_t total = (int64_t)S * 4 * Ho * Wo;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t plane = (int64_t)Hs * Ws;
  if (store_is_u8 && Hs == Ho && Ws == Wo && (plane & 15) == 0 && ((uintptr_t)store & 15) == 0 &&
      ((uintptr_t)rgb & 15) == 0 && ((uintptr_t)mask & 15) == 0) {
    const int64_t items = (int64_t)S * 4 * (plane / 16);
    const int g2 = (int)((items + 255) / 256 < 148 * 16 ? (items + 255) / 256 : 148 * 16);
    gt_convert_u8_kernel<<<g2, 256, 0, st>>>(S, plane / 16, (const uint8_t*)store, slots, rgb, mask);
    DIMO_CHECK_LAUNCH();
    return 0;
  }
  if (store_is_u8)
    gt_fetch_kernel<uint8_t><<<grid, 256, 0, st>>>(S, Hs, Ws, Ho, Wo, sy, sx, (const uint8_t*)store, slots, rgb, mask);
  else
    gt_fetch_kernel<float><<<grid, 256, 0, st>>>(S, Hs, Ws, Ho, Wo, sy, sx, (const float*)store, slots, rgb, mask);
  DIMO_CHECK_LAUNCH();
  return 0;
}
