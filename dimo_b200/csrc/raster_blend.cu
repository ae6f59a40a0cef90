// Per-tile alpha blending, forward and backward (the "renderCUDA" stage of diff_gauss /
// diff_gaussian_rasterization; call site renderer/latent_gs_renderer.py:1256-1277).
//
// B200 design: one CTA per 16x16 tile per frame.  The tile's depth-sorted splat list is a
// contiguous run of 64-byte records (raster_bin.cu packs it), streamed into shared memory by
// 1-D bulk TMA (cp.async.bulk -> mbarrier complete_tx), double-buffered, 256 records (16 KB) per
// stage; every thread then reads records as broadcast LDS.128.  Bound: FP32 FMA + MUFU.EX2 issue
// and shared-memory broadcast bandwidth, HBM secondary.
//
// Backward: back-to-front replay from the tile's deepest contributor; per-splat gradients are
// reduced over the 32 pixels of a warp with shuffles, over the 8 warps in shared memory, and
// flushed with one global atomicAdd per (tile, splat, field) -- ~256x fewer global atomics than a
// per-pixel scheme.
#include "common.cuh"

namespace dimo {

constexpr int CHUNK = 256;                      // splat records per smem stage
constexpr int REC_F4 = DIMO_SPLAT_FLOATS / 4;   // float4 per record
constexpr int NGRAD = 13;                       // gradient fields per splat (x,y,ca,cb,cc,op,r,g,b,depth,nx,ny,nz)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__global__ void __launch_bounds__(TILE_PIX) blend_fwd_kernel(
    int W, int H, int gx, int tiles_per_frame, const float* __restrict__ cams, const float4* __restrict__ packed,
    const uint2* __restrict__ ranges, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_normal, float* __restrict__ out_alpha, float* __restrict__ final_T,
    int32_t* __restrict__ n_contrib) {
  __shared__ __align__(128) float4 sm[2][CHUNK * REC_F4];
  __shared__ __align__(8) uint64_t bar[2];

  const int tile = blockIdx.x;
  const int b = tile / tiles_per_frame;
  const int t = tile - b * tiles_per_frame;
  const int ty = t / gx, tx = t - ty * gx;
  const int tid = threadIdx.y * TILE + threadIdx.x;
  const int pxi = tx * TILE + threadIdx.x, pyi = ty * TILE + threadIdx.y;
  const bool inside = pxi < W && pyi < H;
  const float pxf = (float)pxi, pyf = (float)pyi;

  const uint2 rng = ranges[tile];
  const int n = (int)(rng.y - rng.x);
  const int nchunks = (n + CHUNK - 1) / CHUNK;
  const float4* src = packed + (int64_t)rng.x * REC_F4;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int c = 0; c < 2 && c < nchunks; ++c) {
      const int cnt = min(CHUNK, n - c * CHUNK);
      mbar_expect_tx(&bar[c], cnt * 64);
      bulk_g2s(&sm[c][0], src + (int64_t)c * CHUNK * REC_F4, cnt * 64, &bar[c]);
    }
  }

  float T = 1.0f;
  float Cr = 0.f, Cg = 0.f, Cb = 0.f, D = 0.f, Nx = 0.f, Ny = 0.f, Nz = 0.f;
  int contributor = 0, last = 0;
  bool done = !inside;

  int c = 0;
  for (; c < nchunks; ++c) {
    const int stage = c & 1;
    mbar_wait(&bar[stage], (c >> 1) & 1);
    const int cnt = min(CHUNK, n - c * CHUNK);
    if (!done) {
      const float4* s = &sm[stage][0];
      for (int j = 0; j < cnt; ++j) {
        ++contributor;
        const float4 a = s[j * REC_F4 + 0];   // x, y, conic_a, conic_b
        const float4 bq = s[j * REC_F4 + 1];  // conic_c, opacity, r, g
        const float dx = a.x - pxf, dy = a.y - pyf;
        const float power = -0.5f * (a.z * dx * dx + bq.x * dy * dy) - a.w * dx * dy;
        if (power > 0.0f) continue;
        const float alpha = fminf(ALPHA_MAX, bq.y * __expf(power));
        if (alpha < ALPHA_MIN) continue;
        const float test_T = T * (1.0f - alpha);
        if (test_T < T_MIN) {
          done = true;
          break;
        }
        const float4 cq = s[j * REC_F4 + 2];  // b, depth, nx, ny
        const float nzv = s[j * REC_F4 + 3].x;
        const float w = alpha * T;
        Cr += bq.z * w; Cg += bq.w * w; Cb += cq.x * w;
        D += cq.y * w;
        Nx += cq.z * w; Ny += cq.w * w; Nz += nzv * w;
        T = test_T;
        last = contributor;
      }
    }
    const int num_done = __syncthreads_count(done);
    if (num_done == TILE_PIX) break;
    if (tid == 0 && c + 2 < nchunks) {
      const int cnt2 = min(CHUNK, n - (c + 2) * CHUNK);
      mbar_expect_tx(&bar[stage], cnt2 * 64);
      bulk_g2s(&sm[stage][0], src + (int64_t)(c + 2) * CHUNK * REC_F4, cnt2 * 64, &bar[stage]);
    }
  }
  // an early break can leave chunk c+1 in flight: it must land before this CTA's smem is released
  if (c < nchunks && c + 1 < nchunks) mbar_wait(&bar[(c + 1) & 1], ((c + 1) >> 1) & 1);

  if (inside) {
    const float* bg = cams + (int64_t)b * DIMO_CAM_FLOATS + CAM_BG;
    const int64_t hw = (int64_t)H * W;
    const int64_t pix = (int64_t)pyi * W + pxi;
    out_color[((int64_t)b * 3 + 0) * hw + pix] = Cr + T * bg[0];
    out_color[((int64_t)b * 3 + 1) * hw + pix] = Cg + T * bg[1];
    out_color[((int64_t)b * 3 + 2) * hw + pix] = Cb + T * bg[2];
    out_depth[(int64_t)b * hw + pix] = D;
    out_normal[((int64_t)b * 3 + 0) * hw + pix] = Nx;
    out_normal[((int64_t)b * 3 + 1) * hw + pix] = Ny;
    out_normal[((int64_t)b * 3 + 2) * hw + pix] = Nz;
    out_alpha[(int64_t)b * hw + pix] = 1.0f - T;
    final_T[(int64_t)b * hw + pix] = T;
    n_contrib[(int64_t)b * hw + pix] = last;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(TILE_PIX) blend_bwd_kernel(
    int W, int H, int gx, int tiles_per_frame, const float* __restrict__ cams, const float4* __restrict__ packed,
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ vals_sorted, const float* __restrict__ final_T,
    const int32_t* __restrict__ n_contrib, const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth,
    const float* __restrict__ dL_dnormal, const float* __restrict__ dL_dalpha, float* __restrict__ dL_dsplats) {
  __shared__ __align__(128) float4 sm[2][CHUNK * REC_F4];
  __shared__ float acc[CHUNK * NGRAD];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int s_max;

  const int tile = blockIdx.x;
  const int b = tile / tiles_per_frame;
  const int t = tile - b * tiles_per_frame;
  const int ty = t / gx, tx = t - ty * gx;
  const int tid = threadIdx.y * TILE + threadIdx.x;
  const int lane = tid & 31;
  const int pxi = tx * TILE + threadIdx.x, pyi = ty * TILE + threadIdx.y;
  const bool inside = pxi < W && pyi < H;
  const float pxf = (float)pxi, pyf = (float)pyi;
  const uint2 rng = ranges[tile];
  if (rng.y <= rng.x) return;

  const int64_t hw = (int64_t)H * W;
  const int64_t pix = (int64_t)pyi * W + pxi;
  float gc0 = 0.f, gc1 = 0.f, gc2 = 0.f, gd = 0.f, gn0 = 0.f, gn1 = 0.f, gn2 = 0.f, ga = 0.f, Tf = 1.f;
  int last = 0;
  if (inside) {
    gc0 = dL_dcolor[((int64_t)b * 3 + 0) * hw + pix];
    gc1 = dL_dcolor[((int64_t)b * 3 + 1) * hw + pix];
    gc2 = dL_dcolor[((int64_t)b * 3 + 2) * hw + pix];
    gd = dL_ddepth[(int64_t)b * hw + pix];
    gn0 = dL_dnormal[((int64_t)b * 3 + 0) * hw + pix];
    gn1 = dL_dnormal[((int64_t)b * 3 + 1) * hw + pix];
    gn2 = dL_dnormal[((int64_t)b * 3 + 2) * hw + pix];
    ga = dL_dalpha[(int64_t)b * hw + pix];
    Tf = final_T[(int64_t)b * hw + pix];
    last = n_contrib[(int64_t)b * hw + pix];
  }
  const float* bg = cams + (int64_t)b * DIMO_CAM_FLOATS + CAM_BG;
  // Qp = sum_{j>i} (g.f_j) alpha_j T_j  -  (ga - g_rgb.bg) * T_final
  float Qp = -(ga - (gc0 * bg[0] + gc1 * bg[1] + gc2 * bg[2])) * Tf;
  float T = Tf;

  if (tid == 0) {
    s_max = 0;
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  for (int k = tid; k < CHUNK * NGRAD; k += TILE_PIX) acc[k] = 0.f;
  __syncthreads();
  {
    int m = last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0) atomicMax(&s_max, m);
  }
  __syncthreads();
  const int nproc = s_max;   // list positions [0, nproc) hold every contributor of this tile
  if (nproc == 0) return;
  const int nchunks = (nproc + CHUNK - 1) / CHUNK;
  const float4* src = packed + (int64_t)rng.x * REC_F4;

  // chunk k of the descending walk is list chunk (nchunks-1-k); stage = k&1
  if (tid == 0) {
    for (int k = 0; k < 2 && k < nchunks; ++k) {
      const int cc = nchunks - 1 - k;
      const int cnt = min(CHUNK, nproc - cc * CHUNK);
      mbar_expect_tx(&bar[k], cnt * 64);
      bulk_g2s(&sm[k][0], src + (int64_t)cc * CHUNK * REC_F4, cnt * 64, &bar[k]);
    }
  }

  for (int k = 0; k < nchunks; ++k) {
    const int stage = k & 1;
    const int cc = nchunks - 1 - k;
    const int cnt = min(CHUNK, nproc - cc * CHUNK);
    mbar_wait(&bar[stage], (k >> 1) & 1);
    const float4* s = &sm[stage][0];
    for (int j = cnt - 1; j >= 0; --j) {
      const int pos = cc * CHUNK + j;
      bool active = pos < last;
      float dx = 0.f, dy = 0.f, G = 0.f, alpha = 0.f;
      float4 a, bq;
      if (__any_sync(0xffffffffu, active)) {
        a = s[j * REC_F4 + 0];
        bq = s[j * REC_F4 + 1];
        if (active) {
          dx = a.x - pxf; dy = a.y - pyf;
          const float power = -0.5f * (a.z * dx * dx + bq.x * dy * dy) - a.w * dx * dy;
          if (power > 0.0f) {
            active = false;
          } else {
            G = __expf(power);
            alpha = fminf(ALPHA_MAX, bq.y * G);
            if (alpha < ALPHA_MIN) active = false;
          }
        }
      }
      if (!__any_sync(0xffffffffu, active)) continue;
      float v[NGRAD];
#pragma unroll
      for (int q = 0; q < NGRAD; ++q) v[q] = 0.f;
      if (active) {
        const float4 cq = s[j * REC_F4 + 2];
        const float nzv = s[j * REC_F4 + 3].x;
        const float inv = 1.0f / (1.0f - alpha);
        T = T * inv;
        const float w = alpha * T;
        const float dotf = gc0 * bq.z + gc1 * bq.w + gc2 * cq.x + gd * cq.y + gn0 * cq.z + gn1 * cq.w + gn2 * nzv;
        const float dLa = T * dotf - Qp * inv;
        Qp += dotf * w;
        const float gp = dLa * bq.y * G;
        v[0] = -(a.z * dx + a.w * dy) * gp;
        v[1] = -(bq.x * dy + a.w * dx) * gp;
        v[2] = -0.5f * dx * dx * gp;
        v[3] = -dx * dy * gp;
        v[4] = -0.5f * dy * dy * gp;
        v[5] = G * dLa;
        v[6] = gc0 * w; v[7] = gc1 * w; v[8] = gc2 * w;
        v[9] = gd * w;
        v[10] = gn0 * w; v[11] = gn1 * w; v[12] = gn2 * w;
      }
#pragma unroll
      for (int q = 0; q < NGRAD; ++q) v[q] = warp_sum(v[q]);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NGRAD; ++q) atomicAdd(&acc[j * NGRAD + q], v[q]);
      }
    }
    __syncthreads();   // all reads of sm[stage] and all smem atomics of this chunk are done
    if (tid == 0 && k + 2 < nchunks) {
      const int c2 = nchunks - 1 - (k + 2);
      const int cnt2 = min(CHUNK, nproc - c2 * CHUNK);
      mbar_expect_tx(&bar[stage], cnt2 * 64);
      bulk_g2s(&sm[stage][0], src + (int64_t)c2 * CHUNK * REC_F4, cnt2 * 64, &bar[stage]);
    }
    // flush this chunk's accumulators: one global atomic per (splat, field)
    for (int e = tid; e < cnt * NGRAD; e += TILE_PIX) {
      const int j = e / NGRAD, q = e - j * NGRAD;
      const float val = acc[e];
      acc[e] = 0.f;
      if (val != 0.f) {
        const uint32_t gid = vals_sorted[(int64_t)rng.x + cc * CHUNK + j];
        atomicAdd(&dL_dsplats[(int64_t)gid * DIMO_SPLAT_FLOATS + q], val);
      }
    }
    __syncthreads();
  }
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_raster_blend_fwd(int B, int W, int H, const float* cams, const float* packed,
                                     const uint32_t* ranges, float* out_color, float* out_depth, float* out_normal,
                                     float* out_alpha, float* final_T, int32_t* n_contrib, void* stream) {
  DIMO_REQUIRE(B >= 0 && W > 0 && H > 0, "bad sizes");
  if (B == 0) return 0;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  dim3 block(TILE, TILE);
  blend_fwd_kernel<<<B * gx * gy, block, 0, (cudaStream_t)stream>>>(
      W, H, gx, gx * gy, cams, reinterpret_cast<const float4*>(packed), reinterpret_cast<const uint2*>(ranges),
      out_color, out_depth, out_normal, out_alpha, final_T, n_contrib);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_raster_blend_bwd(int B, int N, int W, int H, const float* cams, const float* packed,
                                     const uint32_t* ranges, const uint32_t* vals_sorted, const float* final_T,
                                     const int32_t* n_contrib, const float* dL_dcolor, const float* dL_ddepth,
                                     const float* dL_dnormal, const float* dL_dalpha, float* dL_dsplats,
                                     void* stream) {
  DIMO_REQUIRE(B >= 0 && W > 0 && H > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0 || N == 0) return 0;
  DIMO_CHECK_CUDA(cudaMemsetAsync(dL_dsplats, 0, sizeof(float) * DIMO_SPLAT_FLOATS * (size_t)B * N, st));
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  dim3 block(TILE, TILE);
  blend_bwd_kernel<<<B * gx * gy, block, 0, st>>>(
      W, H, gx, gx * gy, cams, reinterpret_cast<const float4*>(packed), reinterpret_cast<const uint2*>(ranges),
      vals_sorted, final_T, n_contrib, dL_dcolor, dL_ddepth, dL_dnormal, dL_dalpha, dL_dsplats);
  DIMO_CHECK_LAUNCH();
  return 0;
}
