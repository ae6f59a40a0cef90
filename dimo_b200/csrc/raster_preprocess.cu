// Per-Gaussian projection (forward; the backward lives in raster_preprocess_bwd.cu) for the tile rasteriser.
//
// Replaces the "preprocess" stage inside diff_gauss / diff_gaussian_rasterization's
// GaussianRasterizer (call sites renderer/latent_gs_renderer.py:1256-1266, 1268-1277).
//
// THIS TRANSLATION UNIT IS COMPILED WITH -fmad=false.  The integer outputs of the rasteriser
// (radii, tile rectangles, tile counts, depth keys) are functions of the fp32 values computed
// here; with FMA contraction off, every operation below is a separately rounded IEEE fp32
// operation in exactly the order written, which oracle/raster.py::preprocess mirrors line by line.
// The kernel is HBM-bound (56 B read + 72 B written per Gaussian), so the lost FMA fusion is free.
#include "raster_project.cuh"

namespace dimo {


__global__ void __launch_bounds__(256, 4) preprocess_fwd_kernel(
    int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags,
    const float* __restrict__ cams, const int32_t* __restrict__ frame_src,
    const float* __restrict__ means3D, int64_t means3D_bs,
    const float* __restrict__ scales, int64_t scales_bs,
    const float* __restrict__ rotations, int64_t rot_bs,
    const float* __restrict__ opacities, int64_t op_bs,
    const float* __restrict__ shs, int64_t shs_bs,
    const float* __restrict__ colors, int64_t col_bs,
    float4* __restrict__ splats, int32_t* __restrict__ radii, uint32_t* __restrict__ tiles_touched,
    uint2* __restrict__ rects, uint32_t* __restrict__ depth_keys, unsigned long long* __restrict__ total_count) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // the CTA's tile-instance count is summed (warp shuffles + one shared atomic per warp) and added to the launch
  // set's total with one global atomic per CTA: the instance count R without a scan over the B*N counters
  __shared__ unsigned int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  unsigned int my_tiles = 0;
  if (idx < (int64_t)B * N) {
  const int b = (int)(idx / N);
  const int i = (int)(idx - (int64_t)b * N);
  const float* cam = cams + (int64_t)b * DIMO_CAM_FLOATS;

  const int bsrc = frame_src != nullptr ? frame_src[b] : b;   // deformation block of this frame ((motion, t) pair)
  const float* pm = means3D + bsrc * means3D_bs + 3 * (int64_t)i;
  const float px = pm[0], py = pm[1], pz = pm[2];
  const float* ps = scales + b * scales_bs + 3 * (int64_t)i;
  float sc[3] = {ps[0], ps[1], ps[2]};
  if (act_flags & ACT_EXP_SCALE) { sc[0] = expf(sc[0]); sc[1] = expf(sc[1]); sc[2] = expf(sc[2]); }
  const float4 q4 = *reinterpret_cast<const float4*>(rotations + bsrc * rot_bs + 4 * (int64_t)i);
  const float q[4] = {q4.x, q4.y, q4.z, q4.w};

  Geo g;
  project(cam, px, py, pz, sc, q, scale_modifier, W, H, g);

  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const bool finite = isfinite(g.pix_x) && isfinite(g.pix_y) && isfinite(g.radius_f);
  int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
  if (finite) tile_rect(g.pix_x, g.pix_y, g.radius_f, gx, gy, x0, y0, x1, y1);
  const int ntiles = (x1 - x0) * (y1 - y0);
  const bool visible = (g.tvz > NEAR_CULL_Z) && (g.det != 0.0f) && finite && (ntiles > 0);

  float4* out = splats + 4 * idx;
  if (!visible) {
    radii[idx] = 0;
    tiles_touched[idx] = 0;
    rects[idx] = make_uint2(0u, 0u);
    depth_keys[idx] = 0xFFFFFFFFu;           // culled splats sort to the end of their frame and emit nothing
    const float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
    out[0] = zz; out[1] = zz; out[2] = zz; out[3] = zz;
  } else {

  float rgb[3];
  if (colors != nullptr) {
    const float* pc = colors + b * col_bs + 3 * (int64_t)i;
    rgb[0] = pc[0]; rgb[1] = pc[1]; rgb[2] = pc[2];
  } else {
    const float* cp = cam + CAM_POS;
    float dx = px - cp[0], dy = py - cp[1], dz = pz - cp[2];
    const float len = sqrtf((dx * dx + dy * dy) + dz * dz);
    dx = dx / len; dy = dy / len; dz = dz / len;
    float basis[16];
    sh_basis(sh_degree, dx, dy, dz, basis);
    const int K = (sh_degree + 1) * (sh_degree + 1);
    const float* sh = shs + b * shs_bs + (int64_t)i * sh_coeffs * 3;
    rgb[0] = rgb[1] = rgb[2] = 0.0f;
    for (int k = 0; k < K; ++k) {
      rgb[0] += basis[k] * sh[3 * k + 0];
      rgb[1] += basis[k] * sh[3 * k + 1];
      rgb[2] += basis[k] * sh[3 * k + 2];
    }
    rgb[0] = fmaxf(rgb[0] + 0.5f, 0.0f);
    rgb[1] = fmaxf(rgb[1] + 0.5f, 0.0f);
    rgb[2] = fmaxf(rgb[2] + 0.5f, 0.0f);
  }

  const float* V = cam + CAM_VIEW;
  const float nw0 = (g.kmin == 0 ? g.R[0][0] : (g.kmin == 1 ? g.R[0][1] : g.R[0][2])) * g.nsign;
  const float nw1 = (g.kmin == 0 ? g.R[1][0] : (g.kmin == 1 ? g.R[1][1] : g.R[1][2])) * g.nsign;
  const float nw2 = (g.kmin == 0 ? g.R[2][0] : (g.kmin == 1 ? g.R[2][1] : g.R[2][2])) * g.nsign;
  const float nx = dot3(nw0, nw1, nw2, V[0], V[4], V[8]);
  const float ny = dot3(nw0, nw1, nw2, V[1], V[5], V[9]);
  const float nz = dot3(nw0, nw1, nw2, V[2], V[6], V[10]);
  float op = opacities[b * op_bs + i];
  if (act_flags & ACT_SIGMOID_OPACITY) op = 1.0f / (1.0f + expf(-op));

  radii[idx] = (int)g.radius_f;
  tiles_touched[idx] = (uint32_t)ntiles;
  my_tiles = (unsigned int)ntiles;
  rects[idx] = make_uint2((uint32_t)x0 | ((uint32_t)y0 << 16), (uint32_t)x1 | ((uint32_t)y1 << 16));   // [x0,x1) x [y0,y1), tiles
  depth_keys[idx] = __float_as_uint(g.tvz);   // view depth > 0.2: the bits order as unsigned integers
  // blend record (include/dimo_b200.h; consumed as-is by raster_blend.cu, gathered by index):
  //   x, y, a2, b2 | c2, opacity, pthr2, r | g, b, depth, nx | ny, nz, own index, 0
  // the conic is pre-multiplied into log2 units (alpha = opacity * 2^(a2 dx^2 + b2 dx dy + c2 dy^2));
  // pairs with p2 < pthr2 cannot reach alpha >= 1/255 (0.01 in log2 units = 0.7 % safety margin on alpha)
  constexpr float LOG2E = 1.4426950408889634f;
  const float pthr2 = -log2f(255.0f * op) - 0.01f;
  out[0] = make_float4(g.pix_x, g.pix_y, (-0.5f * LOG2E) * g.conic_a, -LOG2E * g.conic_b);
  out[1] = make_float4((-0.5f * LOG2E) * g.conic_c, op, pthr2, rgb[0]);
  out[2] = make_float4(rgb[1], rgb[2], g.tvz, nx);
  out[3] = make_float4(ny, nz, __uint_as_float((uint32_t)idx), 0.f);
  }   // visible
  }   // idx < B*N
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_tiles += __shfl_xor_sync(0xffffffffu, my_tiles, o);
  if ((threadIdx.x & 31) == 0 && my_tiles) atomicAdd(&s_cnt, my_tiles);
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(total_count, (unsigned long long)s_cnt);
}

// out[u, :] = sum over rows s with seg[s] == u of in[s, :]  (seg == NULL: every row belongs to segment 0), rows
// added in ascending s (deterministic).  Folds the per-frame gradients of the rasteriser backward onto the inputs
// they came from: a (motion, t) deformation shared by several views, or a parameter shared by all frames.
// HBM: 4 B read per input element + 4 B written per output element.
constexpr int SEG_MAX_ROWS = 1024;
template <int VEC>
__global__ void __launch_bounds__(256) segment_sum_kernel(int S, int64_t n, const int32_t* __restrict__ seg,
                                                          const float* __restrict__ in, float* __restrict__ out) {
  __shared__ int32_t s_seg[SEG_MAX_ROWS];
  for (int k = threadIdx.x; k < S; k += blockDim.x) s_seg[k] = seg != nullptr ? seg[k] : 0;
  __syncthreads();
  const int u = blockIdx.y;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (e >= n) return;
  float acc[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
  for (int r = 0; r < S; ++r) {
    if (s_seg[r] != u) continue;
    const float* row = in + (int64_t)r * n + e;
    if (VEC == 4) {
      const float4 v = *reinterpret_cast<const float4*>(row);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    } else {
      acc[0] += row[0];
    }
  }
  float* o = out + (int64_t)u * n + e;
  if (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  else o[0] = acc[0];
}

}  // namespace dimo

namespace dimo {

int preprocess_launch(
    int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags, const float* cams,
    const int32_t* frame_src,
    const float* means3D, int64_t means3D_bstride, const float* scales, int64_t scales_bstride,
    const float* rotations, int64_t rotations_bstride, const float* opacities, int64_t opacities_bstride,
    const float* shs, int64_t shs_bstride, const float* colors_precomp, int64_t colors_bstride,
    float* splats, int32_t* radii, uint32_t* tiles_touched, uint32_t* rects, uint32_t* depth_keys,
    unsigned long long* total_count, cudaStream_t st) {
  const int64_t BN = (int64_t)B * N;
  if (BN == 0) return 0;
  preprocess_fwd_kernel<<<ceil_div(BN, 256), 256, 0, st>>>(
      B, N, W, H, sh_degree, sh_coeffs, scale_modifier, act_flags, cams, frame_src, means3D, means3D_bstride, scales,
      scales_bstride, rotations, rotations_bstride, opacities, opacities_bstride, shs, shs_bstride, colors_precomp,
      colors_bstride, reinterpret_cast<float4*>(splats), radii, tiles_touched, reinterpret_cast<uint2*>(rects),
      depth_keys, total_count);
  DIMO_CHECK_LAUNCH();
  return 0;
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_segment_sum(int S, int U, int64_t n, const int32_t* seg, const float* in, float* out, void* stream) {
  DIMO_REQUIRE(S >= 0 && S <= SEG_MAX_ROWS && U >= 0 && U <= 65535 && n >= 0, "segment_sum: S <= 1024, U <= 65535");
  if (U == 0 || n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (n & 3) == 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0;
  if (vec) {
    dim3 grid(ceil_div(n / 4, 256), U);
    segment_sum_kernel<4><<<grid, 256, 0, st>>>(S, n, seg, in, out);
  } else {
    dim3 grid(ceil_div(n, 256), U);
    segment_sum_kernel<1><<<grid, 256, 0, st>>>(S, n, seg, in, out);
  }
  DIMO_CHECK_LAUNCH();
  return 0;
}
