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
__device__ __forceinline__ void acc_add(float* base, int64_t idx, float v, float det) {
  if (det != 0.f)
    atomicAdd(reinterpret_cast<unsigned long long*>(base) + idx, (unsigned long long)__float2ll_rn(v * det));
  else
    atomicAdd(base + idx, v);
}
#endif

}  // namespace dimo
