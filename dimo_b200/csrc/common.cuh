// Shared device/host helpers for libdimo_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/dimo_b200.h"

namespace dimo {

constexpr int TILE = 16;
constexpr int TILE_PIX = TILE * TILE;

// rasteriser constants -- the oracle (oracle/raster.py) carries the same names
constexpr float NEAR_CULL_Z = 0.2f;
constexpr float FOV_CLAMP = 1.3f;
constexpr float DILATION = 0.3f;
constexpr float LAMBDA_FLOOR = 0.1f;
constexpr float RADIUS_SIGMAS = 3.0f;
constexpr float W_EPS = 1e-7f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float T_MIN = 1e-4f;

// camera block offsets (DIMO_CAM_FLOATS)
constexpr int CAM_VIEW = 0, CAM_PROJ = 16, CAM_POS = 32, CAM_TANX = 35, CAM_TANY = 36, CAM_BG = 37;

void set_error(const char* fmt, ...);

#define DIMO_CHECK_CUDA(expr)                                                          \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      dimo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)

#define DIMO_CHECK_LAUNCH() DIMO_CHECK_CUDA(cudaGetLastError())

#define DIMO_REQUIRE(cond, msg)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      dimo::set_error("%s:%d: %s", __FILE__, __LINE__, msg);      \
      return -2;                                                  \
    }                                                             \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Deterministic accumulation (dimo_set_deterministic): every accumulation target that several CTAs add into is then an
// INT64 buffer of the same element count; addends are rounded to multiples of 1 / DET_SCALE and summed with native 64-bit
// integer reductions, which are associative -- the result does not depend on the order the CTAs arrive in.  Range
// +-3.4e10, resolution 3.7e-9 per addend.  dimo_fixed_to_float converts back.
constexpr float DET_SCALE = 268435456.0f;          // 2^28
float det_scale();                                  // 0 = off (fp32 atomics), DET_SCALE = on   (raster_bin.cu)
#ifdef __CUDACC__
// Packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for two lanes' worth of fp32 math).  The blend
// loops are issue-bound (smsp__issue_active ~80 %, profiles/r2c_*), and every thread runs the same arithmetic on four
// pixels, so the FMA-pipe instructions are issued on pixel PAIRS; compares, selects and MUFU stay scalar.
struct f2 { float x, y; };
__device__ __forceinline__ unsigned long long f2_pack(f2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ f2 f2_unpack(unsigned long long v) {
  f2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ f2 f2_bcast(float a) { return f2{a, a}; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)), "l"(f2_pack(c)));
  return f2_unpack(d);
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(d);
}

__device__ __forceinline__ void acc_add(float* base, int64_t idx, float v, float det) {
  if (det != 0.f)
    atomicAdd(reinterpret_cast<unsigned long long*>(base) + idx, (unsigned long long)__float2ll_rn(v * det));
  else
    atomicAdd(base + idx, v);
}
#endif

}  // namespace dimo
