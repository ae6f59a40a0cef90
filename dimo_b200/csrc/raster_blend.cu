// Per-tile alpha blending, forward and backward (the "renderCUDA" stage of diff_gauss /
// diff_gaussian_rasterization; call site renderer/latent_gs_renderer.py:1256-1277).
//
// B200 design
//  * one CTA (64 threads = 2 warps) per 16x16 tile per frame; every thread owns 4 horizontally
//    adjacent pixels, so the per-splat shared-memory reads (broadcast LDS.128) and the row-dependent
//    part of the quadratic form are amortised over 4 pixels and outputs leave as 128-bit stores;
//  * the tile's depth-sorted list is a run of record indices (`vals_sorted`); the 64-byte "blend records"
//    themselves live once per (frame, Gaussian) in the table the preprocess kernel writes.  Every thread
//    gathers two records per stage into shared memory with 64-byte bulk async copies
//    (cp.async.bulk -> mbarrier complete_tx; or 4 x 16-byte cp.async, DIMO knob 3), double-buffered,
//    128 records (8 KB) per stage, indices prefetched one stage further ahead in registers;
//  * records carry the conic pre-multiplied by -0.5*log2(e) (resp. -log2(e)), so
//    alpha = opacity * ex2(p2) is one MUFU.EX2 without the extra multiply, and a conservative
//    threshold p2 >= -log2(255*opacity) - margin that skips the MUFU for pairs that cannot reach
//    alpha >= 1/255 (the exact test still follows, so results are unchanged);
//  * backward: back-to-front replay from the tile's deepest contributor; per-splat gradients are
//    summed over a thread's 4 pixels, reduced over the warp with shuffles, over the 2 warps in shared
//    memory, and flushed with one global atomicAdd per (tile, splat, field).
// Bound: FP32 FMA + MUFU.EX2 issue; HBM is secondary.
#include "common.cuh"

namespace dimo {

constexpr int CHUNK = 128;                      // blend records per smem stage
constexpr int REC_F4 = DIMO_SPLAT_FLOATS / 4;   // float4 per record
constexpr int PPT = 4;                          // pixels per thread
constexpr int BLEND_THREADS = TILE_PIX / PPT;   // 64
constexpr float LN2 = 0.6931471805599453f;

// blend record layout (written by preprocess_fwd_kernel, raster_preprocess.cu):
//   f4#0: x, y, a2 = -0.5*log2e*conic_a, b2 = -log2e*conic_b
//   f4#1: c2 = -0.5*log2e*conic_c, opacity, pthr2, r
//   f4#2: g, b, depth, nx
//   f4#3: ny, nz, gid (own table index, int bits), 0

// Gather mode (dimo_tc_debug_set key 3): 0 = one 64-byte bulk copy per record, 1 = four 16-byte cp.async per record.
int g_blend_gather_mode = 0;
// Records per shared-memory stage of the backward kernel (dimo_tc_debug_set key 4): 128, or 64 = half the shared memory
// per CTA (12.7 KB instead of 25.6 KB) -> occupancy is no longer shared-memory bound, at twice the barrier rounds.
int g_blend_bwd_chunk = 64;
int g_blend_fwd_chunk = 64;      // same for the forward kernel (dimo_tc_debug_set key 5)

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
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Stage loader shared by both kernels.  `ids` = this tile's record indices; list positions [pos0, pos0 + cnt) go to
// `dst`.  Thread t fetches records t and t + 64 of the stage; their indices were prefetched into id0/id1.
//   MODE 0: bar expects 1 arrival (thread 0, with the stage's byte count) + the copies' complete_tx;
//   MODE 1: bar expects BLEND_THREADS arrivals, each fired when that thread's cp.async group has landed.
template <int MODE>
__device__ __forceinline__ void stage_gather(float4* dst, const float4* __restrict__ table, uint32_t id0, uint32_t id1,
                                             int cnt, int tid, uint64_t* bar) {
  if (MODE == 0) {
    if (tid == 0) mbar_expect_tx(bar, cnt * 64);
    if (tid < cnt) bulk_g2s(dst + tid * REC_F4, table + (int64_t)id0 * REC_F4, 64, bar);
    if (tid + 64 < cnt) bulk_g2s(dst + (tid + 64) * REC_F4, table + (int64_t)id1 * REC_F4, 64, bar);
  } else {
    if (tid < cnt) {
      const float4* src = table + (int64_t)id0 * REC_F4;
#pragma unroll
      for (int q = 0; q < REC_F4; ++q) cp_async16(dst + tid * REC_F4 + q, src + q);
    }
    if (tid + 64 < cnt) {
      const float4* src = table + (int64_t)id1 * REC_F4;
#pragma unroll
      for (int q = 0; q < REC_F4; ++q) cp_async16(dst + (tid + 64) * REC_F4 + q, src + q);
    }
    cp_async_arrive_noinc(bar);
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

static_assert(CHUNK == 2 * BLEND_THREADS, "stage_gather assigns two records per thread");

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));   // one MUFU.RCP (__fdividef expands to a denormal-guarded sequence)
  return y;
}

// Forward.  DN = false: depth / normal are not produced (out_depth = out_normal = NULL; the MSE + SSIM + mask step has
// no consumer for them): 3 instead of 7 accumulators per pixel, 24 instead of 40 output bytes per pixel.
//
// Inner loop, per record and warp:
//   test    p_i for the thread's 4 pixels (2 FFMA each), ONE compare of max(p_i) against the record's conservative
//           threshold, one vote -> records that cannot reach alpha >= 1/255 anywhere in the warp's 16 x 8 pixels cost
//           ~25 instructions;
//   blend   BRANCH-FREE over the 4 pixels: the accept / saturate decisions become predicates, the weight of a
//           rejected pair is selected to 0 and the accumulators are updated unconditionally (x + c * 0 = x exactly), so
//           the four dependent chains (EX2 -> alpha -> T) interleave instead of running one divergent region per pixel.
// A pixel that saturates (or lies outside the image) gets its own alpha threshold raised from 1/255 to 2: no pair can be
// accepted for it any more, so the loop needs no per-pixel "alive" test (the ALU pipe -- compares, selects, min -- is
// the binding one here: 8 ALU + 7 FMA-pipe + 1 MUFU instruction per pixel and record).
template <int MODE, bool DN, int CH>
__global__ void __launch_bounds__(BLEND_THREADS, DN ? 12 : 16) blend_fwd_kernel(
    int W, int H, int gx, int tiles_per_frame, uint32_t vmask, uint32_t frame_stride,
    const float* __restrict__ cams, const float4* __restrict__ table,
    const uint32_t* __restrict__ vals_sorted, const uint2* __restrict__ ranges, float* __restrict__ out_color,
    float* __restrict__ out_depth, float* __restrict__ out_normal, float* __restrict__ out_alpha,
    float* __restrict__ final_T, int32_t* __restrict__ n_contrib) {
  __shared__ __align__(128) float4 sm[2][CH * REC_F4];
  __shared__ __align__(8) uint64_t bar[2];

  const int tile = blockIdx.x;
  const int b = tile / tiles_per_frame;
  const int t = tile - b * tiles_per_frame;
  const int ty = t / gx, tx = t - ty * gx;
  const int tid = threadIdx.x;
  const int px0 = tx * TILE + (tid & 3) * PPT, pyi = ty * TILE + (tid >> 2);
  const float pxf0 = (float)px0, pyf = (float)pyi;
  const f2 pxp[2] = {f2{pxf0, pxf0 + 1.0f}, f2{pxf0 + 2.0f, pxf0 + 3.0f}};     // x of the thread's four pixels

  const uint2 rng = ranges[tile];
  const int n = (int)(rng.y - rng.x);
  const int nchunks = (n + CH - 1) / CH;
  const uint32_t* ids = vals_sorted + rng.x;
  // record indices of list positions c*CH + tid and + 64 (0 when past the end: never dereferenced).  An instance
  // word is the record index itself (vmask = all ones, frame_stride = 0) or, packed, (tile key | index within the
  // frame): record = (word & vmask) + frame * N.
  const uint32_t fbase = (uint32_t)b * frame_stride;
  auto load_ids = [&](int c, uint32_t& i0, uint32_t& i1) {
    const int p0 = c * CH + tid;
    i0 = p0 < n ? (ids[p0] & vmask) + fbase : 0u;
    i1 = (CH > BLEND_THREADS && p0 + 64 < n) ? (ids[p0 + 64] & vmask) + fbase : 0u;
  };

  if (tid == 0) {
    mbar_init(&bar[0], MODE == 0 ? 1 : BLEND_THREADS);
    mbar_init(&bar[1], MODE == 0 ? 1 : BLEND_THREADS);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t nid0 = 0, nid1 = 0;   // indices of the next stage to be issued
  for (int c = 0; c < 2 && c < nchunks; ++c) {
    load_ids(c, nid0, nid1);
    stage_gather<MODE>(&sm[c][0], table, nid0, nid1, min(CH, n - c * CH), tid, &bar[c]);
  }
  if (nchunks > 2) load_ids(2, nid0, nid1);

  constexpr int NCH = DN ? 7 : 3;            // r, g, b [, depth, nx, ny, nz]
  float T[PPT], A[NCH][PPT], amin[PPT];
  int last[PPT];
  unsigned alive = 0;                        // bit i: pixel i is inside the image and not saturated
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    last[i] = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) A[k][i] = 0.f;
    const bool inside = px0 + i < W && pyi < H;
    T[i] = 1.f;
    amin[i] = inside ? ALPHA_MIN : 2.0f;     // outside the image: "saturated" from the start
    if (inside) alive |= 1u << i;
  }

  int c = 0;
  for (; c < nchunks; ++c) {
    const int stage = c & 1;
    mbar_wait(&bar[stage], (c >> 1) & 1);
    const int cnt = min(CH, n - c * CH);
    // The loop is kept WARP-UNIFORM (votes decide every branch): a per-thread `continue`/`break` loop de-converges the
    // warp for good -- the first version of this kernel ran at 1.65 active threads per instruction
    // (profiles/r1_blend_v1_*.csv).
    if (__any_sync(0xffffffffu, alive != 0)) {
      const float4* s = &sm[stage][0];
      const int base = c * CH + 1;
      for (int j = 0; j < cnt; ++j) {
        const float4 a = s[j * REC_F4 + 0];   // x, y, a2, b2
        const float4 bq = s[j * REC_F4 + 1];  // c2, opacity, pthr2, r
        const float dy = a.y - pyf;
        const float by = a.w * dy, cy = bq.x * dy * dy;
        // pixels (0,1) and (2,3) as packed pairs; their x coordinates live in two register pairs for the whole kernel
        // (offsets built from immediates cost four uniform-register moves per visited record)
        f2 pp[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const f2 dx = sub2(f2_bcast(a.x), pxp[h]);
          pp[h] = fma2(fma2(f2_bcast(a.z), dx, f2_bcast(by)), dx, f2_bcast(cy));
        }
        const float p[PPT] = {pp[0].x, pp[0].y, pp[1].x, pp[1].y};
        // pthr2 is conservative (0.7 % margin on alpha): below it no pixel can reach alpha >= 1/255
        const float pmax = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[2], p[3]));
        if (!__any_sync(0xffffffffu, pmax >= bq.z && alive != 0)) continue;
        const float4 cq = s[j * REC_F4 + 2];  // g, b, depth, nx
        float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (DN) dq = s[j * REC_F4 + 3];       // ny, nz, gid, -
        const int idx = base + j;
        bool died[PPT];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i0 = 2 * h, i1 = 2 * h + 1;
          const f2 og = mul2(f2_bcast(bq.y), f2{ex2_approx(p[i0]), ex2_approx(p[i1])});
          const f2 alpha = f2{fminf(ALPHA_MAX, og.x), fminf(ALPHA_MAX, og.y)};
          const f2 Tp = f2{T[i0], T[i1]};
          const f2 test_T = mul2(Tp, sub2(f2_bcast(1.0f), alpha));
          const f2 wa = mul2(alpha, Tp);
          const bool cand0 = (p[i0] <= 0.f) & (alpha.x >= amin[i0]), cand1 = (p[i1] <= 0.f) & (alpha.y >= amin[i1]);
          const bool ok0 = cand0 & (test_T.x >= T_MIN), ok1 = cand1 & (test_T.y >= T_MIN);
          died[i0] = cand0 & !ok0; died[i1] = cand1 & !ok1;
          const f2 w = f2{ok0 ? wa.x : 0.f, ok1 ? wa.y : 0.f};
          f2 acc;
          acc = fma2(f2_bcast(bq.w), w, f2{A[0][i0], A[0][i1]}); A[0][i0] = acc.x; A[0][i1] = acc.y;
          acc = fma2(f2_bcast(cq.x), w, f2{A[1][i0], A[1][i1]}); A[1][i0] = acc.x; A[1][i1] = acc.y;
          acc = fma2(f2_bcast(cq.y), w, f2{A[2][i0], A[2][i1]}); A[2][i0] = acc.x; A[2][i1] = acc.y;
          if (DN) {
            acc = fma2(f2_bcast(cq.z), w, f2{A[3][i0], A[3][i1]}); A[3][i0] = acc.x; A[3][i1] = acc.y;
            acc = fma2(f2_bcast(cq.w), w, f2{A[4][i0], A[4][i1]}); A[4][i0] = acc.x; A[4][i1] = acc.y;
            acc = fma2(f2_bcast(dq.x), w, f2{A[5][i0], A[5][i1]}); A[5][i0] = acc.x; A[5][i1] = acc.y;
            acc = fma2(f2_bcast(dq.y), w, f2{A[6][i0], A[6][i1]}); A[6][i0] = acc.x; A[6][i1] = acc.y;
          }
          T[i0] = ok0 ? test_T.x : T[i0]; T[i1] = ok1 ? test_T.y : T[i1];
          last[i0] = ok0 ? idx : last[i0]; last[i1] = ok1 ? idx : last[i1];
        }
        if (__any_sync(0xffffffffu, died[0] | died[1] | died[2] | died[3])) {      // rare: a pixel saturated here
#pragma unroll
          for (int i = 0; i < PPT; ++i)
            if (died[i]) { amin[i] = 2.0f; alive &= ~(1u << i); }
          if (!__any_sync(0xffffffffu, alive != 0)) break;   // whole warp saturated
        }
      }
    }
    const int num_alive = __syncthreads_count(alive != 0);
    if (num_alive == 0) break;
    if (c + 2 < nchunks) {
      stage_gather<MODE>(&sm[stage][0], table, nid0, nid1, min(CH, n - (c + 2) * CH), tid, &bar[stage]);
      if (c + 3 < nchunks) load_ids(c + 3, nid0, nid1);
    }
  }
  // an early break can leave chunk c+1 in flight: it must land before this CTA's smem is released
  if (c < nchunks && c + 1 < nchunks) mbar_wait(&bar[(c + 1) & 1], ((c + 1) >> 1) & 1);

  if (pyi >= H || px0 >= W) return;
  const float* bg = cams + (int64_t)b * DIMO_CAM_FLOATS + CAM_BG;
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
  const int64_t hw = (int64_t)H * W;
  const int64_t pix = (int64_t)pyi * W + px0;
  float* oc = out_color + (int64_t)b * 3 * hw + pix;
  float* oa = out_alpha + (int64_t)b * hw + pix;
  float* oT = final_T + (int64_t)b * hw + pix;
  int32_t* oN = n_contrib + (int64_t)b * hw + pix;
  float* od = DN ? out_depth + (int64_t)b * hw + pix : nullptr;
  float* on = DN ? out_normal + (int64_t)b * 3 * hw + pix : nullptr;
  if ((W & 3) == 0 && px0 + PPT <= W) {
    // W % 4 == 0 and px0 % 4 == 0: 16-byte aligned rows -> 128-bit stores
    *reinterpret_cast<float4*>(oc) = make_float4(A[0][0] + T[0] * bg0, A[0][1] + T[1] * bg0, A[0][2] + T[2] * bg0, A[0][3] + T[3] * bg0);
    *reinterpret_cast<float4*>(oc + hw) = make_float4(A[1][0] + T[0] * bg1, A[1][1] + T[1] * bg1, A[1][2] + T[2] * bg1, A[1][3] + T[3] * bg1);
    *reinterpret_cast<float4*>(oc + 2 * hw) = make_float4(A[2][0] + T[0] * bg2, A[2][1] + T[1] * bg2, A[2][2] + T[2] * bg2, A[2][3] + T[3] * bg2);
    if (DN) {
      *reinterpret_cast<float4*>(od) = make_float4(A[NCH - 4][0], A[NCH - 4][1], A[NCH - 4][2], A[NCH - 4][3]);
      *reinterpret_cast<float4*>(on) = make_float4(A[NCH - 3][0], A[NCH - 3][1], A[NCH - 3][2], A[NCH - 3][3]);
      *reinterpret_cast<float4*>(on + hw) = make_float4(A[NCH - 2][0], A[NCH - 2][1], A[NCH - 2][2], A[NCH - 2][3]);
      *reinterpret_cast<float4*>(on + 2 * hw) = make_float4(A[NCH - 1][0], A[NCH - 1][1], A[NCH - 1][2], A[NCH - 1][3]);
    }
    *reinterpret_cast<float4*>(oa) = make_float4(1.f - T[0], 1.f - T[1], 1.f - T[2], 1.f - T[3]);
    *reinterpret_cast<float4*>(oT) = make_float4(T[0], T[1], T[2], T[3]);
    *reinterpret_cast<int4*>(oN) = make_int4(last[0], last[1], last[2], last[3]);
  } else {
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      if (px0 + i < W) {
        oc[i] = A[0][i] + T[i] * bg0; oc[hw + i] = A[1][i] + T[i] * bg1; oc[2 * hw + i] = A[2][i] + T[i] * bg2;
        if (DN) {
          od[i] = A[NCH - 4][i];
          on[i] = A[NCH - 3][i]; on[hw + i] = A[NCH - 2][i]; on[2 * hw + i] = A[NCH - 1][i];
        }
        oa[i] = 1.f - T[i]; oT[i] = T[i]; oN[i] = last[i];
      }
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Backward.  Per (pixel, splat) pair the chain rule needs t = G * dL/dalpha and w = alpha * T; everything that is
// quadratic in the pixel offset is accumulated as MOMENTS of t over the thread's 4 pixels
//   St = sum t, Sx = sum dx t, Sxx = sum dx^2 t        (dy is constant per thread: Sy = dy St, Sxy = dy Sx, Syy = dy^2 St)
// and turned into the record's gradients only once per (tile, splat) at flush time (LOG2E * LN2 = 1):
//   dL/dx = op ln2 (2 a2 Sx + b2 Sy)      dL/dy = op ln2 (2 c2 Sy + b2 Sx)      dL/dop = St
//   dL/dconic_a = -op/2 Sxx               dL/dconic_b = -op Sxy                 dL/dconic_c = -op/2 Syy
// which cuts the per-pixel work from ~45 to ~24 instructions.  DN = false (no gradient arrives for depth / normal, e.g.
// the MSE + SSIM + mask step): those four channels are dropped from the pixel loop and the reduction (9 fields: an
// 8-slot transposed butterfly + one plain warp sum = 14 shuffles; DN = true: 13 fields in a 16-slot butterfly).
// Each warp owns its own accumulator rows in shared memory (plain stores, no shared-memory atomics); one thread
// per record folds the two warps' rows and issues 128-bit global reductions.
template <bool DN>
struct BwdFields {
  static constexpr int NG = DN ? 13 : 9;   // St, Sx, Sy, Sxx, Sxy, Syy, r, g, b [, depth, nx, ny, nz]
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  // 16-byte aligned vector reduction (sm_90+): one RED.128 instead of four
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int MODE, bool DN, int CH>
__global__ void __launch_bounds__(BLEND_THREADS) blend_bwd_kernel(
    int W, int H, int gx, int tiles_per_frame, uint32_t vmask, uint32_t frame_stride,
    const float* __restrict__ cams, const float4* __restrict__ table,
    const uint32_t* __restrict__ vals_sorted, const uint2* __restrict__ ranges, const float* __restrict__ final_T,
    const int32_t* __restrict__ n_contrib,
    const float* __restrict__ dL_dcolor, const float* __restrict__ dL_ddepth, const float* __restrict__ dL_dnormal,
    const float* __restrict__ dL_dalpha, float* __restrict__ dL_dsplats, float det) {
  constexpr int NG = BwdFields<DN>::NG;
  __shared__ __align__(128) float4 sm[2][CH * REC_F4];
  __shared__ float acc[2][CH * NG];          // [warp][record][field]
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ int s_max;

  const int tile = blockIdx.x;
  const int b = tile / tiles_per_frame;
  const int t = tile - b * tiles_per_frame;
  const int ty = t / gx, tx = t - ty * gx;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int px0 = tx * TILE + (tid & 3) * PPT, pyi = ty * TILE + (tid >> 2);
  const float pxf0 = (float)px0, pyf = (float)pyi;
  const f2 pxp[2] = {f2{pxf0, pxf0 + 1.0f}, f2{pxf0 + 2.0f, pxf0 + 3.0f}};     // x of the thread's four pixels
  const uint2 rng = ranges[tile];
  if (rng.y <= rng.x) return;

  const int64_t hw = (int64_t)H * W;
  const float* bg = cams + (int64_t)b * DIMO_CAM_FLOATS + CAM_BG;
  const float bg0 = bg[0], bg1 = bg[1], bg2 = bg[2];
  float gc0[PPT], gc1[PPT], gc2[PPT], T[PPT], Qp[PPT];
  float gd[DN ? PPT : 1], gn0[DN ? PPT : 1], gn1[DN ? PPT : 1], gn2[DN ? PPT : 1];
  int last[PPT];
  int lmax = 0;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    gc0[i] = gc1[i] = gc2[i] = 0.f; T[i] = 1.f; Qp[i] = 0.f; last[i] = 0;
    if (DN) { gd[i] = gn0[i] = gn1[i] = gn2[i] = 0.f; }
    if (px0 + i < W && pyi < H) {
      const int64_t pix = (int64_t)pyi * W + px0 + i;
      gc0[i] = dL_dcolor[((int64_t)b * 3 + 0) * hw + pix];
      gc1[i] = dL_dcolor[((int64_t)b * 3 + 1) * hw + pix];
      gc2[i] = dL_dcolor[((int64_t)b * 3 + 2) * hw + pix];
      if (DN) {
        gd[i] = dL_ddepth[(int64_t)b * hw + pix];
        gn0[i] = dL_dnormal[((int64_t)b * 3 + 0) * hw + pix];
        gn1[i] = dL_dnormal[((int64_t)b * 3 + 1) * hw + pix];
        gn2[i] = dL_dnormal[((int64_t)b * 3 + 2) * hw + pix];
      }
      const float ga = dL_dalpha[(int64_t)b * hw + pix];
      T[i] = final_T[(int64_t)b * hw + pix];
      last[i] = n_contrib[(int64_t)b * hw + pix];
      // Qp = sum_{j>i} (g.f_j) alpha_j T_j  -  (ga - g_rgb.bg) * T_final
      Qp[i] = -(ga - (gc0[i] * bg0 + gc1[i] * bg1 + gc2[i] * bg2)) * T[i];
      lmax = max(lmax, last[i]);
    }
  }

  if (tid == 0) {
    s_max = 0;
    mbar_init(&bar[0], MODE == 0 ? 1 : BLEND_THREADS);
    mbar_init(&bar[1], MODE == 0 ? 1 : BLEND_THREADS);
    fence_mbar_init();
  }
  for (int k = tid; k < 2 * CH * NG; k += BLEND_THREADS) (&acc[0][0])[k] = 0.f;
  __syncthreads();
  {
    int m = lmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0) atomicMax(&s_max, m);
  }
  __syncthreads();
  const int nproc = s_max;   // list positions [0, nproc) hold every contributor of this tile
  if (nproc == 0) return;
  const int nchunks = (nproc + CH - 1) / CH;
  const uint32_t* ids = vals_sorted + rng.x;
  const uint32_t fbase = (uint32_t)b * frame_stride;    // instance word -> record index, see blend_fwd_kernel
  auto load_ids = [&](int cc, uint32_t& i0, uint32_t& i1) {
    const int p0 = cc * CH + tid;
    i0 = p0 < nproc ? (ids[p0] & vmask) + fbase : 0u;
    i1 = (CH > BLEND_THREADS && p0 + 64 < nproc) ? (ids[p0 + 64] & vmask) + fbase : 0u;
  };

  // chunk k of the descending walk is list chunk (nchunks-1-k); stage = k&1
  uint32_t nid0 = 0, nid1 = 0;   // indices of the next stage to be issued
  for (int k = 0; k < 2 && k < nchunks; ++k) {
    const int cc = nchunks - 1 - k;
    load_ids(cc, nid0, nid1);
    stage_gather<MODE>(&sm[k][0], table, nid0, nid1, min(CH, nproc - cc * CH), tid, &bar[k]);
  }
  if (nchunks > 2) load_ids(nchunks - 3, nid0, nid1);

  float* const my_acc = &acc[warp][0];
  for (int k = 0; k < nchunks; ++k) {
    const int stage = k & 1;
    const int cc = nchunks - 1 - k;
    const int cnt = min(CH, nproc - cc * CH);
    mbar_wait(&bar[stage], (k >> 1) & 1);
    const float4* s = &sm[stage][0];
    for (int j = cnt - 1; j >= 0; --j) {
      const int pos = cc * CH + j;
      // test: one compare of max(p_i) against the record's conservative threshold + "this thread still has
      // contributors at or behind this list position", one vote (see blend_fwd_kernel)
      const float4 a = s[j * REC_F4 + 0];
      const float4 bq = s[j * REC_F4 + 1];
      const float dy = a.y - pyf;
      const float by = a.w * dy, cy = bq.x * dy * dy;
      f2 dxp[2], pp[2];                 // pixels (0,1) and (2,3) as packed pairs (FFMA2 / FMUL2 / FADD2)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        dxp[h] = sub2(f2_bcast(a.x), pxp[h]);
        pp[h] = fma2(fma2(f2_bcast(a.z), dxp[h], f2_bcast(by)), dxp[h], f2_bcast(cy));
      }
      const float p[PPT] = {pp[0].x, pp[0].y, pp[1].x, pp[1].y};
      const float pmax = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[2], p[3]));
      if (!__any_sync(0xffffffffu, pmax >= bq.z && pos < lmax)) continue;
      // blend: branch-free over the 4 pixels; a pair the forward pass skipped contributes exact zeros
      const float4 cq = s[j * REC_F4 + 2];  // g, b, depth, nx
      float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
      if (DN) dq = s[j * REC_F4 + 3];       // ny, nz, gid, -
      f2 vt2 = f2_bcast(0.f), vsx2 = f2_bcast(0.f), vsxx2 = f2_bcast(0.f), vr2 = f2_bcast(0.f), vg2 = f2_bcast(0.f),
         vb2 = f2_bcast(0.f), vd2 = f2_bcast(0.f), vn02 = f2_bcast(0.f), vn12 = f2_bcast(0.f), vn22 = f2_bcast(0.f);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int i0 = 2 * h, i1 = 2 * h + 1;
        const f2 G = f2{ex2_approx(p[i0]), ex2_approx(p[i1])};
        const f2 og = mul2(f2_bcast(bq.y), G);
        const f2 alpha = f2{fminf(ALPHA_MAX, og.x), fminf(ALPHA_MAX, og.y)};
        const bool valid0 = (pos < last[i0]) & (p[i0] <= 0.f) & (alpha.x >= ALPHA_MIN);
        const bool valid1 = (pos < last[i1]) & (p[i1] <= 0.f) & (alpha.y >= ALPHA_MIN);
        const f2 om = sub2(f2_bcast(1.0f), alpha);
        const f2 inv = f2{rcp_approx(om.x), rcp_approx(om.y)};      // alpha <= 0.99: MUFU.RCP (1 ulp) is ample for the 1e-4 bound
        const f2 Tp = f2{T[i0], T[i1]}, Qq = f2{Qp[i0], Qp[i1]};
        const f2 Tn = mul2(Tp, inv);                                // transmittance in front of this record
        const f2 w = mul2(alpha, Tn);
        const f2 g0 = f2{gc0[i0], gc0[i1]}, g1 = f2{gc1[i0], gc1[i1]}, g2 = f2{gc2[i0], gc2[i1]};
        f2 dotf = fma2(g2, f2_bcast(cq.y), fma2(g1, f2_bcast(cq.x), mul2(g0, f2_bcast(bq.w))));
        f2 gdp = f2_bcast(0.f), gn0p = gdp, gn1p = gdp, gn2p = gdp;
        if (DN) {
          gdp = f2{gd[i0], gd[i1]}; gn0p = f2{gn0[i0], gn0[i1]}; gn1p = f2{gn1[i0], gn1[i1]}; gn2p = f2{gn2[i0], gn2[i1]};
          dotf = fma2(gn2p, f2_bcast(dq.y), fma2(gn1p, f2_bcast(dq.x), fma2(gn0p, f2_bcast(cq.w), fma2(gdp, f2_bcast(cq.z), dotf))));
        }
        const f2 dLa = sub2(mul2(Tn, dotf), mul2(Qq, inv));
        const f2 Qn = fma2(dotf, w, Qq);
        const f2 tg = mul2(G, dLa);                                 // dL/dopacity contribution; dL/dp2 = tt * op * ln2
        const f2 tt = f2{valid0 ? tg.x : 0.f, valid1 ? tg.y : 0.f};
        const f2 wv = f2{valid0 ? w.x : 0.f, valid1 ? w.y : 0.f};
        T[i0] = valid0 ? Tn.x : T[i0]; T[i1] = valid1 ? Tn.y : T[i1];
        Qp[i0] = valid0 ? Qn.x : Qp[i0]; Qp[i1] = valid1 ? Qn.y : Qp[i1];
        const f2 dxt = mul2(dxp[h], tt);
        vt2 = add2(vt2, tt); vsx2 = add2(vsx2, dxt); vsxx2 = fma2(dxp[h], dxt, vsxx2);
        vr2 = fma2(g0, wv, vr2); vg2 = fma2(g1, wv, vg2); vb2 = fma2(g2, wv, vb2);
        if (DN) {
          vd2 = fma2(gdp, wv, vd2);
          vn02 = fma2(gn0p, wv, vn02); vn12 = fma2(gn1p, wv, vn12); vn22 = fma2(gn2p, wv, vn22);
        }
      }
      const float vt = vt2.x + vt2.y, vsx = vsx2.x + vsx2.y, vsxx = vsxx2.x + vsxx2.y;
      const float vr = vr2.x + vr2.y, vg = vg2.x + vg2.y, vb = vb2.x + vb2.y;
      const float vd = vd2.x + vd2.y, vn0 = vn02.x + vn02.y, vn1 = vn12.x + vn12.y, vn2 = vn22.x + vn22.y;
      const float vsy = dy * vt, vsxy = dy * vsx, vsyy = dy * vsy;
      float* const row = my_acc + j * NG;
      if (!DN) {
        // 8-slot transposed butterfly (4+2+1 exchanges, then two folds): lane l ends with the warp total of slot l>>2
        float w8[8] = {vt, vsx, vsy, vsxx, vsxy, vsyy, vr, vg};
#pragma unroll
        for (int half = 4, m = 16; half >= 1; half >>= 1, m >>= 1) {
          const bool upper = (lane & m) != 0;
#pragma unroll
          for (int q = 0; q < half; ++q) {
            const float keep = upper ? w8[q + half] : w8[q];
            const float send = upper ? w8[q] : w8[q + half];
            w8[q] = keep + __shfl_xor_sync(0xffffffffu, send, m);
          }
        }
        float tot = w8[0] + __shfl_xor_sync(0xffffffffu, w8[0], 2);
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        const float totb = warp_sum(vb);
        if ((lane & 3) == 0) row[lane >> 2] = tot;
        if (lane == 1) row[8] = totb;
      } else {
        // 16-slot transposed butterfly (8+4+2+1 exchanges + one fold): lane l ends with the warp total of slot l>>1
        float w16[16] = {vt, vsx, vsy, vsxx, vsxy, vsyy, vr, vg, vb, vd, vn0, vn1, vn2, 0.f, 0.f, 0.f};
#pragma unroll
        for (int half = 8, m = 16; half >= 1; half >>= 1, m >>= 1) {
          const bool upper = (lane & m) != 0;
#pragma unroll
          for (int q = 0; q < half; ++q) {
            const float keep = upper ? w16[q + half] : w16[q];
            const float send = upper ? w16[q] : w16[q + half];
            w16[q] = keep + __shfl_xor_sync(0xffffffffu, send, m);
          }
        }
        const float tot = w16[0] + __shfl_xor_sync(0xffffffffu, w16[0], 1);
        const int slot = lane >> 1;
        if ((lane & 1) == 0 && slot < NG) row[slot] = tot;
      }
    }
    __syncthreads();   // both warps' accumulator rows of this chunk are complete; sm[stage] is still intact
    // flush: one thread per record folds the two warps' rows, converts moments to gradients, 128-bit global reductions
    for (int j = tid; j < cnt; j += BLEND_THREADS) {
      float f[NG];
      bool any = false;
#pragma unroll
      for (int q = 0; q < NG; ++q) {
        const float u0 = acc[0][j * NG + q], u1 = acc[1][j * NG + q];
        f[q] = u0 + u1;
        any |= (u0 != 0.f) | (u1 != 0.f);
      }
      if (any) {
#pragma unroll
        for (int q = 0; q < NG; ++q) { acc[0][j * NG + q] = 0.f; acc[1][j * NG + q] = 0.f; }
        const float4 a = s[j * REC_F4 + 0];    // x, y, a2, b2
        const float4 bq = s[j * REC_F4 + 1];   // c2, opacity, pthr2, r
        const uint32_t gid = __float_as_uint(s[j * REC_F4 + 3].z);
        const float op = bq.y, k = op * LN2;
        const float St = f[0], Sx = f[1], Sy = f[2], Sxx = f[3], Sxy = f[4], Syy = f[5];
        float* out = dL_dsplats + (int64_t)gid * DIMO_SPLAT_FLOATS;
        const float o0 = k * (2.f * a.z * Sx + a.w * Sy), o1 = k * (2.f * bq.x * Sy + a.w * Sx), o2 = -0.5f * op * Sxx,
                    o3 = -op * Sxy, o4 = -0.5f * op * Syy;
        if (det != 0.f) {          // deterministic mode: 64-bit fixed-point reductions into an int64 record table
          const int64_t e0 = (int64_t)gid * DIMO_SPLAT_FLOATS;
          const float o[9] = {o0, o1, o2, o3, o4, St, f[6], f[7], f[8]};
#pragma unroll
          for (int q = 0; q < 9; ++q) acc_add(dL_dsplats, e0 + q, o[q], det);
          if (DN) {
#pragma unroll
            for (int q = 9; q < NG; ++q) acc_add(dL_dsplats, e0 + q, f[q], det);
          }
        } else {
          red_add_v4(out, o0, o1, o2, o3);
          red_add_v4(out + 4, o4, St, f[6], f[7]);
          if (DN) {
            red_add_v4(out + 8, f[8], f[9], f[10], f[11]);
            atomicAdd(out + 12, f[12]);
          } else {
            atomicAdd(out + 8, f[8]);
          }
        }
      }
    }
    __syncthreads();
    if (k + 2 < nchunks) {
      const int c2 = nchunks - 1 - (k + 2);
      stage_gather<MODE>(&sm[stage][0], table, nid0, nid1, min(CH, nproc - c2 * CH), tid, &bar[stage]);
      if (k + 3 < nchunks) load_ids(c2 - 1, nid0, nid1);
    }
  }
}

}  // namespace dimo

using namespace dimo;

namespace dimo {
// instance-word decoding parameters for a launch set (see dimo_raster_packed_value_bits, raster_bin.cu)
static inline void instance_decode(int value_bits, int N, uint32_t& vmask, uint32_t& frame_stride) {
  vmask = value_bits > 0 ? ((1u << value_bits) - 1u) : 0xFFFFFFFFu;
  frame_stride = value_bits > 0 ? (uint32_t)N : 0u;
}
}  // namespace dimo

extern "C" int dimo_raster_blend_fwd(int B, int N, int W, int H, int value_bits, const float* cams, const float* splats,
                                     const uint32_t* vals_sorted, const uint32_t* ranges, float* out_color,
                                     float* out_depth, float* out_normal, float* out_alpha, float* final_T,
                                     int32_t* n_contrib, void* stream) {
  DIMO_REQUIRE(B >= 0 && W > 0 && H > 0, "bad sizes");
  if (B == 0) return 0;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  uint32_t vmask, fstride;
  instance_decode(value_bits, N, vmask, fstride);
  DIMO_REQUIRE((out_depth == nullptr) == (out_normal == nullptr), "out_depth and out_normal: pass both or neither");
  const bool dn = out_depth != nullptr;
  const bool small = g_blend_fwd_chunk == 64;
  auto kern = g_blend_gather_mode == 0
                  ? (dn ? (small ? blend_fwd_kernel<0, true, 64> : blend_fwd_kernel<0, true, 128>)
                        : (small ? blend_fwd_kernel<0, false, 64> : blend_fwd_kernel<0, false, 128>))
                  : (dn ? (small ? blend_fwd_kernel<1, true, 64> : blend_fwd_kernel<1, true, 128>)
                        : (small ? blend_fwd_kernel<1, false, 64> : blend_fwd_kernel<1, false, 128>));
  kern<<<B * gx * gy, BLEND_THREADS, 0, (cudaStream_t)stream>>>(
      W, H, gx, gx * gy, vmask, fstride, cams, reinterpret_cast<const float4*>(splats), vals_sorted,
      reinterpret_cast<const uint2*>(ranges), out_color, out_depth, out_normal, out_alpha, final_T, n_contrib);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_raster_blend_bwd(int B, int N, int W, int H, int value_bits, const float* cams, const float* splats,
                                     const uint32_t* vals_sorted, const uint32_t* ranges, const float* final_T,
                                     const int32_t* n_contrib, const float* dL_dcolor, const float* dL_ddepth,
                                     const float* dL_dnormal, const float* dL_dalpha, float* dL_dsplats,
                                     void* stream) {
  DIMO_REQUIRE(B >= 0 && W > 0 && H > 0, "bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0 || N == 0) return 0;
  const float det = dimo::det_scale();        // deterministic mode: dL_dsplats is an int64 table of the same shape
  DIMO_CHECK_CUDA(cudaMemsetAsync(dL_dsplats, 0, (det != 0.f ? 8 : 4) * DIMO_SPLAT_FLOATS * (size_t)B * N, st));
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  // no gradient for depth AND normal (NULL): the four channels are compiled out of the pixel loop and the reduction
  DIMO_REQUIRE((dL_ddepth == nullptr) == (dL_dnormal == nullptr), "dL_ddepth and dL_dnormal: pass both or neither");
  DIMO_REQUIRE(dL_dcolor != nullptr && dL_dalpha != nullptr, "dL_dcolor / dL_dalpha must not be NULL");
  const bool dn = dL_ddepth != nullptr;
  const bool small = g_blend_bwd_chunk == 64;
  auto kern = g_blend_gather_mode == 0
                  ? (dn ? (small ? blend_bwd_kernel<0, true, 64> : blend_bwd_kernel<0, true, 128>)
                        : (small ? blend_bwd_kernel<0, false, 64> : blend_bwd_kernel<0, false, 128>))
                  : (dn ? (small ? blend_bwd_kernel<1, true, 64> : blend_bwd_kernel<1, true, 128>)
                        : (small ? blend_bwd_kernel<1, false, 64> : blend_bwd_kernel<1, false, 128>));
  uint32_t vmask, fstride;
  instance_decode(value_bits, N, vmask, fstride);
  kern<<<B * gx * gy, BLEND_THREADS, 0, st>>>(
      W, H, gx, gx * gy, vmask, fstride, cams, reinterpret_cast<const float4*>(splats), vals_sorted,
      reinterpret_cast<const uint2*>(ranges), final_T, n_contrib, dL_dcolor, dL_ddepth, dL_dnormal, dL_dalpha,
      dL_dsplats, det);
  DIMO_CHECK_LAUNCH();
  return 0;
}
