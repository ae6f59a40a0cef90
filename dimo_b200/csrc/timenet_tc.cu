// TimeNet (renderer/latent_gs_renderer.py:184-235) as a chain of TMA-fed tcgen05 GEMMs on PRE-SPLIT operands.
//
// 3xTF32:  x = hi + lo (hi = tf32(x), lo = tf32(x - hi));  X W^T ~= Xhi Whi^T + Xhi Wlo^T + Xlo Whi^T, FP32 accumulate in
// TMEM (the 1e-4 parity bound rules out plain TF32 / BF16).  Round 1 split the operands INSIDE every GEMM
// (LDG -> cvt -> STS per K tile: 2.45 us per tile, load / convert / MMA phases serialised).  Here every operand lives in
// global memory already split and already in the shared-memory image the tensor core reads (canonical K-major core
// matrices), so a K tile is moved by ONE bulk asynchronous copy per operand (cp.async.bulk -> mbarrier complete_tx) and
// the kernel is a producer warp / MMA-issuer warp / epilogue warps pipeline with no conversion work on the critical
// path:
//   * weights are packed once per optimizer step (tn_pack_kernel: forward tiles W, data-gradient tiles W^T);
//   * every GEMM's epilogue (TMEM -> registers: bias, ReLU or ReLU-mask) writes its result three ways: split tiles
//     that are the NEXT layer's A operand, TRANSPOSED split tiles that are the weight-gradient GEMM's operand (reduction
//     index = row index contiguous), and plain fp32 where a SIMT consumer needs it (3- / 4-wide heads, embedding);
//   * the weight gradients of all ten 256-wide layers run as ONE grouped launch over (layer, tile, row split).
//
// Tile formats ("plane" = tf32 values in 32-bit words; a split tile is the hi plane followed by the lo plane):
//   operand tile   [rows x 32 k]  offset(row, q) = ((row >> 3) * 8 + q) * 128 + (row & 7) * 16 bytes, q = 16-byte chunk
//                  (UMMA canonical K-major, no swizzle: LBO = 128 B between k chunks, SBO = 1024 B between 8-row groups)
//   activations    A[rb][kt]      rb = 128-row block, kt = 32-column block of the activation matrix; 2 x 16 KB
//   weights        B[kt]          rows = output features of the GEMM (padded to a multiple of 16), 2 x rows*128 B
//   transposed     T[rt][plane][fb]  rt = 32-row block, fb = 128-feature block; tile rows = features, k = row index
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dimo {

constexpr int TN_BM = 128;                 // rows per CTA
constexpr int TN_KT = 32;                  // k per stage
constexpr int TN_PLANE_A = TN_BM * TN_KT * 4;          // 16 KB
constexpr int TN_STAGE_A = 2 * TN_PLANE_A;             // 32 KB  (hi | lo)
constexpr int TN_MAXN = 256;
constexpr int TN_STAGE_B = 2 * TN_MAXN * TN_KT * 4;    // 64 KB
constexpr int TN_STAGE = TN_STAGE_A + TN_STAGE_B;      // 96 KB
constexpr int TN_SMEM = 2 * TN_STAGE + 1024;
constexpr int TN_THREADS = 192;            // warps 0-3 epilogue, 4 producer, 5 MMA issuer
// Transposed tiles: same canonical K-major layout but with the k chunks 144 B apart instead of 128 (16 B of padding per
// core matrix).  Their writers are warps whose 32 lanes hold 32 consecutive k (= row) indices of ONE tile row: with a
// 128 B chunk stride the eight chunks of a warp store fall on the same four shared-memory banks (8-way conflict);
// 144 B = 36 words shifts every chunk by four banks -> conflict-free.  The tensor core does not care (LBO is free).
constexpr int TT_LBO = 144, TT_SBO = 8 * TT_LBO;       // 1152
constexpr int TN_FB_BYTES = 16 * TT_SBO;               // 18 KB: one 128-feature block of a transposed tile plane

__device__ __forceinline__ uint32_t tn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tn_split4(const float4 v, uint4& hi, uint4& lo) {
  hi = make_uint4(tn_tf32(v.x), tn_tf32(v.y), tn_tf32(v.z), tn_tf32(v.w));
  lo = make_uint4(tn_tf32(v.x - __uint_as_float(hi.x)), tn_tf32(v.y - __uint_as_float(hi.y)),
                  tn_tf32(v.z - __uint_as_float(hi.z)), tn_tf32(v.w - __uint_as_float(hi.w)));
}
__device__ __forceinline__ void tn_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tn_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tn_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tn_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tn_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tn_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tn_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(tn_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tn_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tn_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tn_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint64_t tn_desc_ls(uint32_t saddr, uint32_t lbo, uint32_t sbo) {   // no swizzle, version 1
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint64_t tn_desc(uint32_t saddr) { return tn_desc_ls(saddr, 128u, 1024u); }
__device__ __forceinline__ uint64_t tn_desc_t(uint32_t saddr) { return tn_desc_ls(saddr, TT_LBO, TT_SBO); }
__device__ __forceinline__ uint32_t tn_idesc(int M, int N) {      // kind::tf32, FP32 accumulate, both operands K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tn_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tn_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tn_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tn_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __host__ __forceinline__ uint32_t tn_off_t(int row, int q) {   // the same inside a transposed tile plane
  return (uint32_t)((row >> 3) * TT_SBO + q * TT_LBO + ((row & 7) << 4));
}
// byte offset of (row, 16-byte chunk q) inside an operand tile plane
__device__ __host__ __forceinline__ uint32_t tn_off(int row, int q) { return (uint32_t)((((row >> 3) * 8 + q) << 7) + ((row & 7) << 4)); }

// ---------------------------------------------------------------------------------------------------------------
// weight packing: W [No, ld] (row-major) -> operand tiles
// ---------------------------------------------------------------------------------------------------------------
struct PackJob {
  const float* W; int ld;
  int rows;          // valid tile rows
  int rows_pad;      // tile rows (multiple of 16)
  int kk;            // valid reduction length
  int nkt;           // 32-wide k tiles
  int transposed;    // 0: tile(row n, k) = W[n][col0 + k];  1: tile(row c, k) = W[k][col0 + c]
  int col0;
  uint8_t* dst;      // nkt split tiles of 2 * rows_pad * 128 bytes
  int chunk0;        // first global chunk id of this job (prefix over jobs)
};
constexpr int TN_MAX_PACK = 32;
struct PackArgs { int n; int total_chunks; PackJob job[TN_MAX_PACK]; };

__global__ void __launch_bounds__(256) tn_pack_kernel(const __grid_constant__ PackArgs a) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.total_chunks) return;
  int j = 0;
  while (j + 1 < a.n && g >= a.job[j + 1].chunk0) ++j;
  const PackJob& p = a.job[j];
  const int c = g - p.chunk0;                       // chunk index: (kt, row, q)
  const int q = c & 7, row = (c >> 3) % p.rows_pad, kt = (c >> 3) / p.rows_pad;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = kt * TN_KT + q * 4 + e;
    float x = 0.f;
    if (row < p.rows && k < p.kk) x = p.transposed ? p.W[(int64_t)k * p.ld + p.col0 + row] : p.W[(int64_t)row * p.ld + p.col0 + k];
    v[e] = x;
  }
  uint4 hi, lo;
  tn_split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
  const size_t plane = (size_t)p.rows_pad * 128;
  uint8_t* t = p.dst + (size_t)kt * 2 * plane + tn_off(row, q);
  *reinterpret_cast<uint4*>(t) = hi;
  *reinterpret_cast<uint4*>(t + plane) = lo;
}

// ---------------------------------------------------------------------------------------------------------------
// plain fp32 [R, ld] (cols valid columns, optional ReLU mask from a second plain matrix) -> split tiles + transposed tiles
// ---------------------------------------------------------------------------------------------------------------
struct TilesArgs {
  int R, Rp, cols, nkt;
  const float* X; int64_t ldx;
  const float* mask; int64_t ldm;      // optional: X(r,c) *= [mask(r,c) > 0]
  uint8_t* out; int out_nkt, out_kt0;  // A[rb][kt] (NULL: skip)
  uint8_t* outT; int t_nfb, t_f0;      // T[rt][plane][fb] (NULL: skip); feature offset of column 0
};

__global__ void __launch_bounds__(256) tn_tiles_kernel(const TilesArgs p) {
  // one thread per (row, 16-byte chunk); consecutive threads take consecutive rows (coalesced tile stores)
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int nq = p.nkt * 8;
  if (g >= (int64_t)p.Rp * nq) return;
  const int r = (int)(g % p.Rp), qq = (int)(g / p.Rp);
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = qq * 4 + e;
    float x = 0.f;
    if (r < p.R && c < p.cols) {
      x = p.X[(int64_t)r * p.ldx + c];
      if (p.mask != nullptr && !(p.mask[(int64_t)r * p.ldm + c] > 0.f)) x = 0.f;
    }
    v[e] = x;
  }
  uint4 hi, lo;
  tn_split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
  const int rb = r >> 7, rl = r & 127, kt = qq >> 3, q = qq & 7;
  if (p.out != nullptr) {
    uint8_t* t = p.out + ((size_t)rb * p.out_nkt + p.out_kt0 + kt) * TN_STAGE_A + tn_off(rl, q);
    *reinterpret_cast<uint4*>(t) = hi;
    *reinterpret_cast<uint4*>(t + TN_PLANE_A) = lo;
  }
  if (p.outT != nullptr) {
    const int rt = r >> 5, rq = (r & 31) >> 2, re = r & 3;
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = p.t_f0 + qq * 4 + e, fb = f >> 7, fl = f & 127;
      uint8_t* t = p.outT + (((size_t)rt * 2) * p.t_nfb + fb) * TN_FB_BYTES + tn_off_t(fl, rq) + re * 4;
      *reinterpret_cast<uint32_t*>(t) = h[e];
      *reinterpret_cast<uint32_t*>(t + (size_t)p.t_nfb * TN_FB_BYTES) = l[e];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the 3- and 4-wide output layers (pts_layers.2, rot_layers.2) in FP32 SIMT: one launch per direction
// ---------------------------------------------------------------------------------------------------------------
// value (hi + lo) of 4 consecutive columns 4 qq .. 4 qq + 3 of row r of an activation stored as split tiles A[rb][kt]
__device__ __forceinline__ float4 tn_tile_load4(const uint8_t* __restrict__ tiles, int nkt, int r, int qq) {
  const uint8_t* t = tiles + ((size_t)(r >> 7) * nkt + (qq >> 3)) * TN_STAGE_A + tn_off(r & 127, qq & 7);
  const float4 h = *reinterpret_cast<const float4*>(t), l = *reinterpret_cast<const float4*>(t + TN_PLANE_A);
  return make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
}

// forward: one warp per row; dxyz[r, :] = hp[r, :] W9^T + b9, dquat[r, :] = hr[r, :] W11^T + b11 (hp / hr: split tiles)
__global__ void __launch_bounds__(256) tn_heads_fwd_kernel(int R, const uint8_t* __restrict__ hp, const uint8_t* __restrict__ hr,
                                                           const float* __restrict__ W9, const float* __restrict__ b9,
                                                           const float* __restrict__ W11, const float* __restrict__ b11,
                                                           float* __restrict__ dxyz, float* __restrict__ dquat) {
  __shared__ float w[7][256];
  for (int e = threadIdx.x; e < 7 * 256; e += blockDim.x) w[e >> 8][e & 255] = (e >> 8) < 3 ? W9[e] : W11[e - 3 * 256];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int c = half * 128 + lane * 4;
    const float4 a = tn_tile_load4(hp, 8, (int)r, c >> 2);
    const float4 b = tn_tile_load4(hr, 8, (int)r, c >> 2);
#pragma unroll
    for (int j = 0; j < 3; ++j) acc[j] += a.x * w[j][c] + a.y * w[j][c + 1] + a.z * w[j][c + 2] + a.w * w[j][c + 3];
#pragma unroll
    for (int j = 3; j < 7; ++j) acc[j] += b.x * w[j][c] + b.y * w[j][c + 1] + b.z * w[j][c + 2] + b.w * w[j][c + 3];
  }
#pragma unroll
  for (int j = 0; j < 7; ++j)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  if (lane < 3) dxyz[r * 3 + lane] = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : acc[2]) + b9[lane];
  else if (lane < 7) dquat[r * 4 + (lane - 3)] = (lane == 3 ? acc[3] : lane == 4 ? acc[4] : lane == 5 ? acc[5] : acc[6]) + b11[lane - 3];
}

// backward: CTA = (32-row block, head) -- 4 x the CTAs of a 128-row split: the kernel is latency-bound (strided tile
// reads, scattered transposed stores) and 128 CTAs left most of the GPU idle.  dH = (g W) * [H > 0] -> split tiles +
// transposed tiles (the first data-gradient GEMM's A operand / the weight-gradient operand);
// dW[j, c] += sum_r g[r, j] H[r, c]; db[j] += sum_r g[r, j]
constexpr int HB_ROWS = 32;
__global__ void __launch_bounds__(256) tn_heads_bwd_kernel(int R, int Rp, const float* __restrict__ g_dxyz,
                                                           const float* __restrict__ g_dquat, const uint8_t* __restrict__ hp,
                                                           const uint8_t* __restrict__ hr, const float* __restrict__ W9,
                                                           const float* __restrict__ W11, uint8_t* __restrict__ out_p,
                                                           uint8_t* __restrict__ outT_p, uint8_t* __restrict__ out_r,
                                                           uint8_t* __restrict__ outT_r, float* __restrict__ dW9,
                                                           float* __restrict__ db9, float* __restrict__ dW11,
                                                           float* __restrict__ db11, float det) {
  const int head = blockIdx.y, tid = threadIdx.x;
  const int row0 = blockIdx.x * HB_ROWS;             // first row of this CTA (never straddles a 128-row tile)
  const int rb = row0 >> 7, rl0 = row0 & 127;
  const int No = head ? 4 : 3;
  const float* g = head ? g_dquat : g_dxyz;
  const uint8_t* H = head ? hr : hp;                 // split tiles of the head's hidden activation
  const float* W = head ? W11 : W9;
  uint8_t* out = head ? out_r : out_p;
  uint8_t* outT = head ? outT_r : outT_p;
  float* dW = head ? dW11 : dW9;
  float* db = head ? db11 : db9;
  __shared__ float sg[HB_ROWS][4];
  __shared__ float sw[4][256];
  for (int e = tid; e < HB_ROWS * 4; e += 256) {
    const int r = row0 + (e >> 2), j = e & 3;
    sg[e >> 2][j] = (r < R && j < No) ? g[(int64_t)r * No + j] : 0.f;
  }
  for (int e = tid; e < 4 * 256; e += 256) sw[e >> 8][e & 255] = (e >> 8) < No ? W[e] : 0.f;
  __syncthreads();
  // ---- masked data gradient -> tiles: item = (row, 16-byte chunk), consecutive threads take consecutive rows ----
  for (int it = tid; it < HB_ROWS * 64; it += 256) {
    const int rr = it & (HB_ROWS - 1), qq = it / HB_ROWS;    // qq: chunk of 4 columns, 0..63
    const int rl = rl0 + rr;
    const int r = row0 + rr;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < R) {
      const float4 h = tn_tile_load4(H, 8, r, qq);
      const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = qq * 4 + e;
        const float d = sg[rr][0] * sw[0][c] + sg[rr][1] * sw[1][c] + sg[rr][2] * sw[2][c] + sg[rr][3] * sw[3][c];
        v[e] = hv[e] > 0.f ? d : 0.f;
      }
    }
    uint4 hi, lo;
    tn_split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
    const int kt = qq >> 3, q = qq & 7;
    uint8_t* t = out + ((size_t)rb * 8 + kt) * TN_STAGE_A + tn_off(rl, q);
    *reinterpret_cast<uint4*>(t) = hi;
    *reinterpret_cast<uint4*>(t + TN_PLANE_A) = lo;
    const int rt = r >> 5, rq = (r & 31) >> 2, re = r & 3;
    const uint32_t hh[4] = {hi.x, hi.y, hi.z, hi.w}, ll[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = qq * 4 + e, fb = f >> 7, fl = f & 127;
      uint8_t* tt = outT + (((size_t)rt * 2) * 2 + fb) * TN_FB_BYTES + tn_off_t(fl, rq) + re * 4;
      *reinterpret_cast<uint32_t*>(tt) = hh[e];
      *reinterpret_cast<uint32_t*>(tt + (size_t)2 * TN_FB_BYTES) = ll[e];
    }
  }
  // ---- weight / bias gradient of the head: thread = column c ----
  {
    const int c = tid;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int rows = max(0, min(HB_ROWS, R - row0));
    for (int r0 = 0; r0 < rows; r0 += 16) {           // 16 independent loads in flight
      float h[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        h[u] = 0.f;
        if (r0 + u < rows) {
          const uint8_t* t = H + ((size_t)rb * 8 + (c >> 5)) * TN_STAGE_A + tn_off(rl0 + r0 + u, (c & 31) >> 2) + (c & 3) * 4;
          h[u] = *reinterpret_cast<const float*>(t) + *reinterpret_cast<const float*>(t + TN_PLANE_A);
        }
      }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (r0 + u < rows) {
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(sg[r0 + u][j], h[u], acc[j]);
        }
      }
    }
    if (rows > 0) {
      for (int j = 0; j < No; ++j) acc_add(dW, (int64_t)j * 256 + c, acc[j], det);
      if (tid < No) {
        float s = 0.f;
        for (int rr = 0; rr < rows; ++rr) s += sg[rr][tid];
        acc_add(db, tid, s, det);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// GEMM  D[128 rows, N] = sum over segments  A_seg[rows, k] * B_seg[N, k]^T
// ---------------------------------------------------------------------------------------------------------------
struct TnSeg { const uint8_t* a; int a_nkt, a_kt0; const uint8_t* b; int nkt; };
struct TnGemmArgs {
  int R, N;                 // valid rows; output columns: 128 or 256 (the weight tiles hold N rows)
  int nseg; TnSeg seg[2];
  const float* bias; int relu;
  const uint8_t* mask; int mask_nkt, mask_kt0;     // ReLU mask source: split tiles A[rb][kt] of the activation whose sign gates column block kt
  uint8_t* out; int out_nkt, out_kt0;              // split tiles of the result (next GEMM's A), or NULL
  uint8_t* outT; int t_nfb, t_f0;                  // transposed split tiles (weight-gradient operand), or NULL
  float* plain; int64_t ldp; int plain_cols, plain_acc;   // fp32 row-major copy of the first plain_cols columns, or NULL
  unsigned long long* dbg;                         // bring-up: [cta][8] %globaltimer stamps (dimo_timenet_debug_stamps), or NULL
};

__device__ __forceinline__ void tn_stamp(unsigned long long* dbg, int slot) {
  if (dbg != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + slot] = t;
  }
}

// One CTA = 128 rows x NB output columns (grid.y = N / NB).  10 warps: 0-7 epilogue (warp w: TMEM lanes 32 (w & 3),
// 32-column chunks (w >> 2), (w >> 2) + 2, ...), 8 producer, 9 MMA issuer.  The layers are bound by the L2 -> shared
// memory operand traffic (every operand is a (hi, lo) pair: 8 B per element), so the tile shape follows the row count:
//   NB =  64: two 48 KB stages, two CTAs per SM -- 128 CTAs for a 4096-row layer (one wave), 100 MB per 8192-row layer;
//   NB = 128: three 64 KB stages, one CTA per SM -- 128 CTAs for an 8192-row layer, 67 MB (the A tiles are re-read by
//             two column blocks instead of four).
// Phase times of a K = 256 layer (tools/tn_stamps.py, NB = 64): first stage lands at 1.8 us, MMAs done at ~6 us.
constexpr int TG_THREADS = 320;
template <int NB>
struct TgCfg {
  static constexpr int STAGE_B = 2 * NB * TN_KT * 4;               // 16 / 32 KB
  static constexpr int STAGE = TN_STAGE_A + STAGE_B;               // 48 / 64 KB
  static constexpr int STAGES = NB == 64 ? 2 : 3;
  static constexpr int SMEM = STAGES * STAGE + 1024;
  static constexpr int TSTAGE = 2 * (NB / 8) * TT_SBO;             // transposed staging per lane group: (hi, lo) x NB tile rows
  static_assert(4 * TSTAGE <= STAGES * STAGE && 128 * (NB + 1) * 4 <= STAGES * STAGE, "epilogue staging");
};

__device__ __forceinline__ void tn_named_bar(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// The three warp roles of ONE layer (shared by the single-layer kernel and the chain kernel).  `it_base`: k tiles this
// CTA has pushed through its stage ring before this layer (mbarrier phases continue across layers); `done_parity`:
// parity of the layers this CTA has finished.
template <int NB>
__device__ __forceinline__ void tn_layer_roles(const TnGemmArgs& p, uint8_t* smem, uint64_t* full_bar, uint64_t* empty_bar,
                                               uint64_t& done_bar, uint32_t tmem_d, int it_base, uint32_t done_parity,
                                               int rb, int cb) {
  using C = TgCfg<NB>;
  constexpr int TG_STAGES = C::STAGES, TG_STAGE = C::STAGE, TG_TSTAGE = C::TSTAGE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t planeB_full = (uint32_t)p.N * 128u;          // bytes of one plane of a packed weight tile (all N rows)
  constexpr uint32_t planeB = NB * 128u;                      // this CTA's NB rows of it
  if (warp == 8) {
    // ===== producer: three bulk copies per stage (A split tile, B hi rows, B lo rows) =====
    if (lane == 0) {
      int it = it_base;
      for (int sg = 0; sg < p.nseg; ++sg) {
        const TnSeg& s = p.seg[sg];
        for (int kt = 0; kt < s.nkt; ++kt, ++it) {
          const int st = it % TG_STAGES, round = it / TG_STAGES;
          if (round > 0) tn_mbar_wait(&empty_bar[st], (uint32_t)((round - 1) & 1));
          uint8_t* sa = smem + st * TG_STAGE;
          uint8_t* sb = sa + TN_STAGE_A;
          const uint8_t* bt = s.b + (size_t)kt * 2 * planeB_full + (size_t)cb * planeB;
          tn_mbar_expect_tx(&full_bar[st], TN_STAGE_A + 2 * planeB);
          tn_bulk_g2s(sa, s.a + ((size_t)rb * s.a_nkt + s.a_kt0 + kt) * TN_STAGE_A, TN_STAGE_A, &full_bar[st]);
          tn_bulk_g2s(sb, bt, planeB, &full_bar[st]);
          tn_bulk_g2s(sb + planeB, bt + planeB_full, planeB, &full_bar[st]);
        }
      }
    }
  } else if (warp == 9) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = tn_idesc(TN_BM, NB);
      int total = 0;
      for (int sg = 0; sg < p.nseg; ++sg) total += p.seg[sg].nkt;
      for (int k = 0; k < total; ++k) {
        const int it = it_base + k;
        const int st = it % TG_STAGES, round = it / TG_STAGES;
        tn_mbar_wait(&full_bar[st], (uint32_t)(round & 1));
        if (k == 0) tn_stamp(p.dbg, 2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = tn_smem_u32(smem + st * TG_STAGE), sb = sa + TN_STAGE_A;
#pragma unroll
        for (int ks = 0; ks < TN_KT / 8; ++ks) {
          const uint32_t koff = (uint32_t)ks * 256u;
          const uint64_t dah = tn_desc(sa + koff), dal = tn_desc(sa + TN_PLANE_A + koff);
          const uint64_t dbh = tn_desc(sb + koff), dbl = tn_desc(sb + planeB + koff);
          tn_mma(tmem_d, dah, dbh, idesc, (k > 0 || ks > 0) ? 1u : 0u);
          tn_mma(tmem_d, dah, dbl, idesc, 1u);
          tn_mma(tmem_d, dal, dbh, idesc, 1u);
        }
        tn_commit(&empty_bar[st]);
      }
      tn_commit(&done_bar);
      tn_stamp(p.dbg, 3);
    }
  } else {
    // ===== epilogue: 32-column chunks (w >> 2), (w >> 2) + 2, ... of the lane group's 32 rows =====
    const int lg = warp & 3;                       // TMEM lane group = tile rows [32 lg, 32 lg + 32)
    const int rl = lg * 32 + lane;
    const int gr = rb * TN_BM + rl;
    const bool valid = gr < p.R;
    const uint32_t row_off = tn_off(rl, 0);        // (row, chunk q) -> row_off + 128 q
    if (tid == 0) tn_stamp(p.dbg, 4);
    bool waited = false;
    for (int ch = warp >> 2; ch < NB / 32; ch += 2) {
      const int lc0 = ch * 32;                     // column within the CTA's NB
      const int n0 = cb * NB + lc0;                // column of the GEMM's N
      // everything that does not depend on the accumulator is fetched first (for the first chunk: while the MMAs run)
      float bias_r[32];
      uint32_t mbits = 0xFFFFFFFFu;
      if (p.bias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = b4[q];
          bias_r[4 * q] = bb.x; bias_r[4 * q + 1] = bb.y; bias_r[4 * q + 2] = bb.z; bias_r[4 * q + 3] = bb.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) bias_r[j] = 0.f;
      }
      if (p.mask != nullptr) {
        const uint8_t* mt = p.mask + ((size_t)rb * p.mask_nkt + p.mask_kt0 + (n0 >> 5)) * TN_STAGE_A + row_off;
        uint32_t bits = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 mh = *reinterpret_cast<const float4*>(mt + 128 * q);
          const float4 ml = *reinterpret_cast<const float4*>(mt + TN_PLANE_A + 128 * q);
          bits |= (uint32_t)(mh.x > 0.f || ml.x > 0.f) << (4 * q);
          bits |= (uint32_t)(mh.y > 0.f || ml.y > 0.f) << (4 * q + 1);
          bits |= (uint32_t)(mh.z > 0.f || ml.z > 0.f) << (4 * q + 2);
          bits |= (uint32_t)(mh.w > 0.f || ml.w > 0.f) << (4 * q + 3);
        }
        mbits = bits;
      }
      if (!waited) {
        tn_mbar_wait(&done_bar, done_parity);
        if (tid == 0) tn_stamp(p.dbg, 5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        waited = true;
      }
      uint32_t v[32];
      tn_ld32(tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)lc0, v);
      // per group of four columns: bias / ReLU / mask, split into (hi, lo), and every store that wants the values
      uint8_t* tT = smem + (size_t)lg * TG_TSTAGE + (size_t)(lc0 >> 3) * TT_SBO + (lane >> 2) * TT_LBO + (lane & 3) * 4;
      uint8_t* tO = p.out != nullptr ? p.out + ((size_t)rb * p.out_nkt + p.out_kt0 + (n0 >> 5)) * TN_STAGE_A + row_off : nullptr;
      float* sp = reinterpret_cast<float*>(smem) + rl * (NB + 1) + lc0;   // plain staging (jobs without transposed output)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 4 * q + e;
          float t = __uint_as_float(v[j]) + bias_r[j];
          if (p.relu) t = fmaxf(t, 0.f);
          const float y = (valid && ((mbits >> j) & 1u)) ? t : 0.f;
          hi[e] = tn_tf32(y);
          lo[e] = tn_tf32(y - __uint_as_float(hi[e]));
          if (p.outT != nullptr) {
            *reinterpret_cast<uint32_t*>(tT + (j >> 3) * TT_SBO + (j & 7) * 16) = hi[e];
            *reinterpret_cast<uint32_t*>(tT + TG_TSTAGE / 2 + (j >> 3) * TT_SBO + (j & 7) * 16) = lo[e];
          }
          if (p.plain != nullptr) sp[j] = y;
        }
        if (tO != nullptr) {
          *reinterpret_cast<uint4*>(tO + 128 * q) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(tO + TN_PLANE_A + 128 * q) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    if (p.outT != nullptr) {
      // transposed tiles: staged in the (now idle) operand stages as the exact image of this CTA's part of
      // T[rt = 4 rb + lg][plane][fb] -- NB consecutive tile rows, contiguous in global memory -- and written by one
      // bulk store per (lane group, plane) as soon as the group's two warps are through
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // staged image -> bulk-copy engine
      tn_named_bar(1 + lg, 64);
      if (warp < 4 && lane == 0) {
        const int f0 = p.t_f0 + cb * NB, fb = f0 >> 7, fl0 = f0 & 127;
#pragma unroll
        for (int pl = 0; pl < 2; ++pl) {
          const uint8_t* src = smem + (size_t)lg * TG_TSTAGE + (size_t)pl * (TG_TSTAGE / 2);
          uint8_t* dst = p.outT + ((((size_t)(rb * 4 + lg)) * 2 + pl) * p.t_nfb + fb) * TN_FB_BYTES + (size_t)(fl0 >> 3) * TT_SBO;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(tn_smem_u32(src)),
                       "r"((uint32_t)(TG_TSTAGE / 2))
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else if (p.plain != nullptr) {
      // fp32 row-major copy: through shared memory so that the global accesses are row-contiguous
      const float* sr = reinterpret_cast<const float*>(smem);
      tn_named_bar(5, 256);
      for (int r = warp; r < TN_BM; r += 8) {
        const int g2 = rb * TN_BM + r;
        if (g2 >= p.R) break;
#pragma unroll
        for (int h = 0; h < NB / 32; ++h) {
          const int col = cb * NB + h * 32 + lane;
          if (col < p.plain_cols) {
            float* dst = p.plain + (int64_t)g2 * p.ldp + col;
            const float val = sr[r * (NB + 1) + h * 32 + lane];
            *dst = p.plain_acc ? *dst + val : val;
          }
        }
      }
    }
    if (p.outT != nullptr && warp < 4 && lane == 0)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // shared memory may be released after the reads
    if (tid == 0) tn_stamp(p.dbg, 6);
  }
}

template <int NB>
__global__ void __launch_bounds__(TG_THREADS, NB == 64 ? 2 : 1) tn_gemm_kernel(const __grid_constant__ TnGemmArgs p) {
  using C = TgCfg<NB>;
  constexpr int TG_STAGES = C::STAGES, TG_STAGE = C::STAGE, TG_TSTAGE = C::TSTAGE;
  extern __shared__ uint8_t tn_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TG_STAGES], empty_bar[TG_STAGES], done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tn_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rb = blockIdx.x, cb = blockIdx.y;
  if (tid == 0) tn_stamp(p.dbg, 0);
  const uint32_t planeB_full = (uint32_t)p.N * 128u;          // bytes of one plane of a packed weight tile (all N rows)
  constexpr uint32_t planeB = NB * 128u;                      // this CTA's NB rows of it

  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tn_smem_u32(&tmem_slot)),
                 "r"((uint32_t)NB)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TG_STAGES; ++s) { tn_mbar_init(&full_bar[s], 1); tn_mbar_init(&empty_bar[s], 1); }
    tn_mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;
  if (tid == 0) tn_stamp(p.dbg, 1);

  tn_layer_roles<NB>(p, smem, full_bar, empty_bar, done_bar, tmem_d, 0, 0u, rb, cb);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) tn_stamp(p.dbg, 7);
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)NB) : "memory");
  }
}

// The whole chain of layers in ONE launch.  Rows are independent, so the only dependency between layer l and l + 1
// is inside a 128-row block: the four CTAs that own its four 64-column slabs.  They form a thread-block cluster
// (1 x 4 x 1); a layer's output tiles go to global memory as before (the backward needs them anyway) and a CLUSTER
// barrier -- not a kernel boundary -- separates the layers: ~1 us instead of the ~5 us of launch gap, prologue and
// teardown per layer.  Generic-proxy stores are fenced into the async proxy before the barrier because the next layer
// reads them with bulk copies.
struct TnChainArgs { int nlayers; TnGemmArgs layer[10]; };

__global__ void __cluster_dims__(1, 4, 1) __launch_bounds__(TG_THREADS, 2) tn_chain_kernel(const __grid_constant__ TnChainArgs c) {
  constexpr int NB = 64;
  using C = TgCfg<NB>;
  constexpr int TG_STAGES = C::STAGES;
  extern __shared__ uint8_t tn_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TG_STAGES], empty_bar[TG_STAGES], done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tn_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rb = blockIdx.x, cb = blockIdx.y;
  cg::cluster_group cluster = cg::this_cluster();
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tn_smem_u32(&tmem_slot)),
                 "r"((uint32_t)NB)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TG_STAGES; ++s) { tn_mbar_init(&full_bar[s], 1); tn_mbar_init(&empty_bar[s], 1); }
    tn_mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;
  int it_base = 0;
  uint32_t n_done = 0;
  for (int L = 0; L < c.nlayers; ++L) {
    const TnGemmArgs& p = c.layer[L];
    if (cb * NB < p.N) {                                   // (128-column layers use two of the four CTAs)
      tn_layer_roles<NB>(p, smem, full_bar, empty_bar, done_bar, tmem_d, it_base, n_done & 1u, rb, cb);
      for (int sg = 0; sg < p.nseg; ++sg) it_base += p.seg[sg].nkt;
      ++n_done;
    }
    if (L + 1 < c.nlayers) {
      asm volatile("fence.proxy.async;" ::: "memory");     // this thread's global stores -> visible to bulk copies
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      cluster.sync();                                      // release / acquire at cluster scope
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)NB) : "memory");
  }
}

// tile width by row count: see the comment above tn_gemm_kernel
static inline int tn_pick_nb(int nrb, int N) { return (nrb >= 64 && N % 128 == 0) ? 128 : 64; }

static int tn_launch_gemm(const TnGemmArgs& g, int nrb, cudaStream_t st) {
  if (tn_pick_nb(nrb, g.N) == 128)
    tn_gemm_kernel<128><<<dim3(nrb, g.N / 128), TG_THREADS, TgCfg<128>::SMEM, st>>>(g);
  else
    tn_gemm_kernel<64><<<dim3(nrb, g.N / 64), TG_THREADS, TgCfg<64>::SMEM, st>>>(g);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------------------------
// grouped weight gradient  dW[n, col0 + j] += sum_r dYm[r, n] * X[r, f0 + j]   (+ db[n] += sum_r dYm[r, n])
// A = dYm^T tiles (128-feature block a_fb), B = X^T tiles (feature blocks [b_fb0, b_fb0 + b_nfb_use)), reduction over the
// row index in 32-row stages [rt0, rt1)
// ---------------------------------------------------------------------------------------------------------------
struct TnWJob {
  const uint8_t* aT; int a_nfb, a_fb;
  const uint8_t* bT; int b_nfb, b_fb0, b_use;      // N = 128 * b_use (1 or 2)
  float* dW; int ldw, col0, ncols, nrows;          // output rows n < nrows, columns j < ncols
  float* db;                                       // NULL unless this job also owns the bias gradient of its n block
};
constexpr int TN_MAX_WJOBS = 32;
constexpr int TW_STAGE = 2 * TN_FB_BYTES + 4 * TN_FB_BYTES;      // A (hi, lo) + B (hi, lo) x up to 2 feature blocks = 108 KB
constexpr int TW_SMEM = 2 * TW_STAGE + 1024;
struct TnWTable { int n, splits, per, nrt; float det; TnWJob job[TN_MAX_WJOBS]; };    // CTA = (job, row split)

__global__ void __launch_bounds__(TN_THREADS, 1) tn_wgrad_kernel(const __grid_constant__ TnWTable tab) {
  extern __shared__ uint8_t tn_smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tn_smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TnWJob& p = tab.job[blockIdx.x / tab.splits];
  const int split = blockIdx.x % tab.splits;
  const int rt0 = split * tab.per, rt1 = min(tab.nrt, rt0 + tab.per);
  const int N = 128 * p.b_use;
  const uint32_t planeB = (uint32_t)p.b_use * TN_FB_BYTES;    // transposed tiles: 18 KB per 128 features and plane
  const int nst = max(0, rt1 - rt0);

  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tn_smem_u32(&tmem_slot)),
                 "r"((uint32_t)TN_MAXN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    tn_mbar_init(&full_bar[0], 1); tn_mbar_init(&full_bar[1], 1);
    tn_mbar_init(&empty_bar[0], 1 + 128); tn_mbar_init(&empty_bar[1], 1 + 128);   // MMA commit + the 128 bias-gradient readers
    tn_mbar_init(&done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int st = it & 1;
        if (it >= 2) tn_mbar_wait(&empty_bar[st], (uint32_t)(((it >> 1) - 1) & 1));
        uint8_t* sa = smem + st * TW_STAGE;
        uint8_t* sb = sa + 2 * TN_FB_BYTES;
        const size_t rt = (size_t)(rt0 + it);
        tn_mbar_expect_tx(&full_bar[st], 2 * TN_FB_BYTES + 2 * planeB);
        tn_bulk_g2s(sa, p.aT + ((rt * 2 + 0) * p.a_nfb + p.a_fb) * TN_FB_BYTES, TN_FB_BYTES, &full_bar[st]);
        tn_bulk_g2s(sa + TN_FB_BYTES, p.aT + ((rt * 2 + 1) * p.a_nfb + p.a_fb) * TN_FB_BYTES, TN_FB_BYTES, &full_bar[st]);
        tn_bulk_g2s(sb, p.bT + ((rt * 2 + 0) * p.b_nfb + p.b_fb0) * TN_FB_BYTES, planeB, &full_bar[st]);
        tn_bulk_g2s(sb + planeB, p.bT + ((rt * 2 + 1) * p.b_nfb + p.b_fb0) * TN_FB_BYTES, planeB, &full_bar[st]);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = tn_idesc(TN_BM, N);
      for (int it = 0; it < nst; ++it) {
        const int st = it & 1;
        tn_mbar_wait(&full_bar[st], (uint32_t)((it >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = tn_smem_u32(smem + st * TW_STAGE), sb = sa + 2 * TN_FB_BYTES;
#pragma unroll
        for (int ks = 0; ks < TN_KT / 8; ++ks) {
          const uint32_t koff = (uint32_t)ks * 2u * TT_LBO;
          const uint64_t dah = tn_desc_t(sa + koff), dal = tn_desc_t(sa + TN_FB_BYTES + koff);
          const uint64_t dbh = tn_desc_t(sb + koff), dbl = tn_desc_t(sb + planeB + koff);
          tn_mma(tmem_d, dah, dbh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          tn_mma(tmem_d, dah, dbl, idesc, 1u);
          tn_mma(tmem_d, dal, dbh, idesc, 1u);
        }
        tn_commit(&empty_bar[st]);
      }
      tn_commit(&done_bar);
    }
  } else {
    // epilogue warps: thread = output row n (feature of dY); bias gradient = row sums of the A tiles as they pass by
    const int nl = warp * 32 + lane;
    float dbacc = 0.f;
    for (int it = 0; it < nst; ++it) {
      const int st = it & 1;
      tn_mbar_wait(&full_bar[st], (uint32_t)((it >> 1) & 1));
      if (p.db != nullptr) {
        const uint8_t* sa = smem + st * TW_STAGE;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 h = *reinterpret_cast<const float4*>(sa + tn_off_t(nl, q));
          const float4 l = *reinterpret_cast<const float4*>(sa + TN_FB_BYTES + tn_off_t(nl, q));
          dbacc += (h.x + l.x) + (h.y + l.y) + (h.z + l.z) + (h.w + l.w);
        }
      }
      tn_mbar_arrive(&empty_bar[st]);
    }
    tn_mbar_wait(&done_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int n = p.a_fb * 128 + nl;
    if (nst > 0) {
      for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        tn_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
        if (n < p.nrows) {
          float* wrow = p.dW + (int64_t)n * p.ldw + p.col0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (tab.det != 0.f) {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c * 32 + j + e < p.ncols)
                  acc_add(p.dW, (int64_t)n * p.ldw + p.col0 + c * 32 + j + e, __uint_as_float(v[j + e]), tab.det);
            } else if (c * 32 + j + 3 < p.ncols) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(wrow + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                           : "memory");
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (c * 32 + j + e < p.ncols) atomicAdd(wrow + j + e, __uint_as_float(v[j + e]));
            }
          }
        }
      }
      if (p.db != nullptr && n < p.nrows) acc_add(p.db, n, dbacc, tab.det);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)TN_MAXN) : "memory");
  }
}

}  // namespace dimo

using namespace dimo;

// ---------------------------------------------------------------------------------------------------------------
// host side: workspace layout + launch sequences
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int HID = 256, NL = 12;             // layers: deformnet.0..7, pts.0, pts.2, rot.0, rot.2
struct TnLayout {
  int R, Rp, nrb, nrt, E, Ep;                 // E = embedding width (104), Ep = padded to 128
  size_t w_fwd[NL], w_bwd[NL], w_bwd5b;       // packed weight tiles (byte offsets)
  size_t h0, catT, cat_plain;                 // embedding: split tiles (4 kt), transposed [rt][2][3 fb], plain [R, E]
  size_t y[10], yT[10];                       // outputs of layers 0..7, 8 (hp), 10 (hr): split tiles (8 kt) / transposed (2 fb)
  size_t hp_plain, hr_plain;                  // plain [R, 256] inputs of the SIMT head layers
  size_t g[10], gT[10];                       // masked upstream gradients of the same ten outputs: split tiles / transposed
  size_t dhp_plain, dhr_plain, dcat_plain;    // plain scratch of the backward pass
  size_t total;
};

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

TnLayout tn_layout(int R, int L) {
  TnLayout o{};
  o.R = R; o.Rp = (R + 127) / 128 * 128; o.nrb = o.Rp / 128; o.nrt = o.Rp / 32;
  o.E = 72 + L; o.Ep = 128;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t r = off; off = al256(off + bytes); return r; };
  const size_t wtile = 2 * (size_t)HID * 128;                 // split weight tile, 256 rows
  const int fwd_kt[NL] = {4, 8, 8, 8, 8, 12, 8, 8, 8, 0, 8, 0};
  for (int l = 0; l < NL; ++l) o.w_fwd[l] = fwd_kt[l] ? take(fwd_kt[l] * wtile) : 0;
  for (int l = 0; l < NL; ++l) {
    if (l == 9 || l == 11) { o.w_bwd[l] = 0; continue; }
    const size_t rows_pad = l == 0 ? 128 : HID;              // output columns of the data-gradient GEMM
    o.w_bwd[l] = take(8 * 2 * rows_pad * 128);
  }
  o.w_bwd5b = take(8 * 2 * (size_t)128 * 128);
  o.h0 = take((size_t)o.nrb * 4 * TN_STAGE_A);
  o.catT = take((size_t)o.nrt * 2 * 3 * TN_FB_BYTES);
  o.cat_plain = take((size_t)R * o.E * 4);
  for (int i = 0; i < 10; ++i) {
    o.y[i] = take((size_t)o.nrb * 8 * TN_STAGE_A);
    o.yT[i] = i == 4 ? 0 : take((size_t)o.nrt * 2 * 2 * TN_FB_BYTES);       // layer 4's output lives in catT (fb 1, 2)
    o.g[i] = take((size_t)o.nrb * 8 * TN_STAGE_A);
    o.gT[i] = take((size_t)o.nrt * 2 * 2 * TN_FB_BYTES);
  }
  o.hp_plain = take((size_t)R * HID * 4); o.hr_plain = take((size_t)R * HID * 4);
  o.dhp_plain = take((size_t)R * HID * 4); o.dhr_plain = take((size_t)R * HID * 4);
  o.dcat_plain = take((size_t)R * o.E * 4);
  o.total = off;
  return o;
}

int tn_set_attrs() {
  static bool done = false;
  if (!done) {
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(tn_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TgCfg<64>::SMEM));
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(tn_gemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TgCfg<128>::SMEM));
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(tn_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TgCfg<64>::SMEM));
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(tn_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TW_SMEM));
    done = true;
  }
  return 0;
}

}  // namespace

extern "C" int dimo_linear_fwd(int R, int K, int No, const float* X, int64_t ldx, const float* Wt, const float* bias,
                               float* Y, int64_t ldy, int relu, void* stream);
extern "C" int dimo_linear_bwd_data(int R, int K, int No, const float* dY, int64_t lddy, const float* Y, int64_t ldy,
                                    const float* Wt, float* dX, int64_t lddx, int accumulate, void* stream);
extern "C" int dimo_linear_bwd_weight(int R, int K, int No, const float* dY, int64_t lddy, const float* Y, int64_t ldy,
                                      const float* X, int64_t ldx, float* dW, float* db, void* stream);
extern "C" int dimo_timenet_embed_fwd(int G, int rows_per_group, int L, const float* pts, const float* times,
                                      const float* latents, float* h0, int64_t ldh, void* stream);
extern "C" int dimo_timenet_embed_bwd(int G, int rows_per_group, int L, const float* h0, const float* dh0, int64_t ldh,
                                      float* dpts, float* dlatents, void* stream);

// all layers of one direction in one cluster launch (tn_chain_kernel) when the 64-column tiles are in use and nobody
// asked for per-layer phase stamps; DIMO_TN_CHAIN=0 falls back to one launch per layer
static bool tn_use_chain(int nrb) {
  static const bool on = [] { const char* e = getenv("DIMO_TN_CHAIN"); return !(e != nullptr && e[0] == '0'); }();
  return on && tn_pick_nb(nrb, 256) == 64;
}

static int g_tn_last_launches = 0;
/* kernel launches issued by the most recent dimo_timenet_fwd / dimo_timenet_bwd call of this process (5 / 4 with the
 * chained GEMMs, 14 / 13 with one launch per layer): bench.py's gpu_launches counts with it */
extern "C" int dimo_timenet_last_launches(void) { return g_tn_last_launches; }

static unsigned long long* g_tn_dbg = nullptr;
static int g_tn_dbg_launch = 0;
/* bring-up: device buffer of 32 x 128 x 8 u64; the following tn_gemm launches stamp [launch % 32][cta][8] with %globaltimer (NULL: off) */
extern "C" int dimo_timenet_debug_stamps(void* dev_buf) {
  g_tn_dbg = reinterpret_cast<unsigned long long*>(dev_buf);
  g_tn_dbg_launch = 0;
  return 0;
}

extern "C" size_t dimo_timenet_workspace_bytes(int G, int M, int L) {
  if (G <= 0 || M <= 0) return 256;
  return tn_layout(G * M, L).total;
}

/* byte offsets of the stored activations inside the workspace (tests / debugging): out[0] = padded rows, out[1] = offset
 * of the embedding's split tiles (4 k tiles), out[2..11] = split tiles (8 k tiles) of the outputs of deformnet.0..7,
 * pts_layers.0, rot_layers.0 */
extern "C" int dimo_timenet_layout(int G, int M, int L, int64_t* out12_host) {
  const TnLayout o = tn_layout(G * M, L);
  out12_host[0] = o.Rp; out12_host[1] = (int64_t)o.h0;
  for (int i = 0; i < 10; ++i) out12_host[2 + i] = (int64_t)o.y[i];
  return 0;
}

// index of a layer's output among the ten stored activations (0..7 trunk, 8 = pts_layers.0, 9 = rot_layers.0)
static inline int tn_slot(int layer) { return layer <= 7 ? layer : (layer == 8 ? 8 : 9); }

extern "C" int dimo_timenet_fwd(int G, int M, int L, const float* pts, const float* times, const float* latents,
                                const float* const* W_host, const float* const* b_host, void* workspace,
                                size_t workspace_bytes, float* dxyz, float* dquat, void* stream) {
  const int R = G * M;
  if (R == 0) return 0;
  DIMO_REQUIRE(L >= 0 && 72 + L <= 128 && (72 + L) % 4 == 0, "TimeNet: embedding width must be a multiple of 4 and <= 128");
  const TnLayout o = tn_layout(R, L);
  DIMO_REQUIRE(workspace_bytes >= o.total, "TimeNet workspace too small (dimo_timenet_workspace_bytes)");
  DIMO_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "TimeNet workspace must be 256-byte aligned");
  if (tn_set_attrs()) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int E = o.E;

  // ---- pack the weights: forward tiles W and data-gradient tiles W^T, one launch ----
  {
    PackArgs a{};
    int n = 0, chunks = 0;
    auto add = [&](const float* W, int ld, int rows, int rows_pad, int kk, int nkt, int transposed, int col0, uint8_t* dst) {
      PackJob& j = a.job[n++];
      j.W = W; j.ld = ld; j.rows = rows; j.rows_pad = rows_pad; j.kk = kk; j.nkt = nkt; j.transposed = transposed;
      j.col0 = col0; j.dst = dst; j.chunk0 = chunks;
      chunks += nkt * rows_pad * 8;
    };
    add(W_host[0], E, HID, HID, E, 4, 0, 0, ws + o.w_fwd[0]);
    for (int l = 1; l <= 8; ++l) {
      if (l == 5) {                       // K order of layer 5: [out of layer 4 (256) | embedding (E)]
        add(W_host[5], E + HID, HID, HID, HID, 8, 0, E, ws + o.w_fwd[5]);
        add(W_host[5], E + HID, HID, HID, E, 4, 0, 0, ws + o.w_fwd[5] + 8 * 2 * (size_t)HID * 128);
      } else {
        add(W_host[l], HID, HID, HID, HID, 8, 0, 0, ws + o.w_fwd[l]);
      }
    }
    add(W_host[10], HID, HID, HID, HID, 8, 0, 0, ws + o.w_fwd[10]);
    // data gradient: tile rows = input features, reduction over the layer's 256 outputs
    add(W_host[0], E, E, 128, HID, 8, 1, 0, ws + o.w_bwd[0]);
    for (int l = 1; l <= 8; ++l) {
      if (l == 5) {
        add(W_host[5], E + HID, HID, HID, HID, 8, 1, E, ws + o.w_bwd[5]);
        add(W_host[5], E + HID, E, 128, HID, 8, 1, 0, ws + o.w_bwd5b);
      } else {
        add(W_host[l], HID, HID, HID, HID, 8, 1, 0, ws + o.w_bwd[l]);
      }
    }
    add(W_host[10], HID, HID, HID, HID, 8, 1, 0, ws + o.w_bwd[10]);
    a.n = n; a.total_chunks = chunks;
    tn_pack_kernel<<<ceil_div(chunks, 256), 256, 0, st>>>(a);
    DIMO_CHECK_LAUNCH();
  }
  // ---- embedding (plain) -> split tiles + transposed tiles (feature block 0 of catT) ----
  float* cat_plain = reinterpret_cast<float*>(ws + o.cat_plain);
  int rc = dimo_timenet_embed_fwd(G, M, L, pts, times, latents, cat_plain, E, stream);
  if (rc) return rc;
  {
    TilesArgs t{};
    t.R = R; t.Rp = o.Rp; t.cols = E; t.nkt = 4; t.X = cat_plain; t.ldx = E;
    t.out = ws + o.h0; t.out_nkt = 4; t.out_kt0 = 0; t.outT = ws + o.catT; t.t_nfb = 3; t.t_f0 = 0;
    tn_tiles_kernel<<<ceil_div((int64_t)o.Rp * 32, 256), 256, 0, st>>>(t);
    DIMO_CHECK_LAUNCH();
  }
  // ---- the ten 256-wide layers ----
  const bool chained = tn_use_chain(o.nrb) && g_tn_dbg == nullptr;
  TnChainArgs chain{};
  auto gemm = [&](TnGemmArgs& g) {
    if (chained) { g.dbg = nullptr; chain.layer[chain.nlayers++] = g; return 0; }
    g.dbg = g_tn_dbg != nullptr ? g_tn_dbg + (size_t)(g_tn_dbg_launch++ % 32) * 128 * 8 : nullptr;
    return tn_launch_gemm(g, o.nrb, st);
  };
  const int order[10] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10};
  for (int oi = 0; oi < 10; ++oi) {
    const int l = order[oi], slot = tn_slot(l);
    TnGemmArgs g{};
    g.R = R; g.N = HID; g.bias = b_host[l]; g.relu = 1;
    if (l == 0) {
      g.nseg = 1; g.seg[0] = TnSeg{ws + o.h0, 4, 0, ws + o.w_fwd[0], 4};
    } else if (l == 5) {
      g.nseg = 2;
      g.seg[0] = TnSeg{ws + o.y[4], 8, 0, ws + o.w_fwd[5], 8};
      g.seg[1] = TnSeg{ws + o.h0, 4, 0, ws + o.w_fwd[5] + 8 * 2 * (size_t)HID * 128, 4};
    } else {
      const int src = l == 10 ? 7 : l - 1;        // rot_layers.0 reads the trunk output like pts_layers.0
      g.nseg = 1; g.seg[0] = TnSeg{ws + o.y[tn_slot(src)], 8, 0, ws + o.w_fwd[l], 8};
    }
    g.out = ws + o.y[slot]; g.out_nkt = 8; g.out_kt0 = 0;
    if (l == 4) { g.outT = ws + o.catT; g.t_nfb = 3; g.t_f0 = 128; }
    else if (l == 8 || l == 10) { g.outT = nullptr; }                 // hp / hr feed only the SIMT heads
    else { g.outT = ws + o.yT[slot]; g.t_nfb = 2; g.t_f0 = 0; }
    if (gemm(g)) { dimo::set_error("tn_gemm_kernel launch failed (layer %d)", l); return -1; }
  }
  if (chained) {
    tn_chain_kernel<<<dim3(o.nrb, 4), TG_THREADS, TgCfg<64>::SMEM, st>>>(chain);
    DIMO_CHECK_LAUNCH();
  }
  g_tn_last_launches = 4 + (chained ? 1 : 10);      // pack, embed, tiles, GEMMs, heads
  // ---- 3- and 4-wide heads (FP32 SIMT, one launch) ----
  tn_heads_fwd_kernel<<<ceil_div(R, 8), 256, 0, st>>>(R, ws + o.y[8], ws + o.y[9], W_host[9], b_host[9], W_host[11],
                                                      b_host[11], dxyz, dquat);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_timenet_bwd(int G, int M, int L, const float* const* W_host, void* workspace, size_t workspace_bytes,
                                const float* g_dxyz, const float* g_dquat, float* const* dW_host, float* const* db_host,
                                float* dpts, float* dlatents, void* stream) {
  const int R = G * M;
  if (R == 0) return 0;
  const TnLayout o = tn_layout(R, L);
  DIMO_REQUIRE(workspace_bytes >= o.total, "TimeNet workspace too small (dimo_timenet_workspace_bytes)");
  if (tn_set_attrs()) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  const int E = o.E;
  float* dcat = reinterpret_cast<float*>(ws + o.dcat_plain);
  // ---- heads (FP32 SIMT, one launch): weight / bias gradients + masked data gradients as split tiles ----
  tn_heads_bwd_kernel<<<dim3(o.Rp / HB_ROWS, 2), 256, 0, st>>>(R, o.Rp, g_dxyz, g_dquat, ws + o.y[8], ws + o.y[9], W_host[9], W_host[11], ws + o.g[8],
                                                      ws + o.gT[8], ws + o.g[9], ws + o.gT[9], dW_host[9], db_host[9],
                                                      dW_host[11], db_host[11], dimo::det_scale());
  DIMO_CHECK_LAUNCH();
  int rc = 0;
  const bool chained = tn_use_chain(o.nrb) && g_tn_dbg == nullptr;
  TnChainArgs chain{};
  auto gemm = [&](TnGemmArgs& g) {
    if (chained) { g.dbg = nullptr; chain.layer[chain.nlayers++] = g; return 0; }
    g.dbg = g_tn_dbg != nullptr ? g_tn_dbg + (size_t)(g_tn_dbg_launch++ % 32) * 128 * 8 : nullptr;
    return tn_launch_gemm(g, o.nrb, st);
  };
  // ---- d(out of layer 7) = dhp_m W8 + dhr_m W10, masked by the sign of layer 7's output ----
  {
    TnGemmArgs g{};
    g.R = R; g.N = HID; g.nseg = 2;
    g.seg[0] = TnSeg{ws + o.g[8], 8, 0, ws + o.w_bwd[8], 8};
    g.seg[1] = TnSeg{ws + o.g[9], 8, 0, ws + o.w_bwd[10], 8};
    g.mask = ws + o.y[7]; g.mask_nkt = 8; g.mask_kt0 = 0;
    g.out = ws + o.g[7]; g.out_nkt = 8; g.outT = ws + o.gT[7]; g.t_nfb = 2;
    if (gemm(g)) { dimo::set_error("tn_gemm_kernel launch failed (heads data gradient)"); return -1; }
  }
  // ---- trunk, layers 7 .. 1: d(out of layer l-1) = dY_l W_l masked ----
  for (int l = 7; l >= 1; --l) {
    TnGemmArgs g{};
    g.R = R; g.N = HID; g.nseg = 1;
    g.seg[0] = TnSeg{ws + o.g[l], 8, 0, ws + o.w_bwd[l], 8};
    g.mask = ws + o.y[l - 1]; g.mask_nkt = 8; g.mask_kt0 = 0;
    g.out = ws + o.g[l - 1]; g.out_nkt = 8; g.outT = ws + o.gT[l - 1]; g.t_nfb = 2;
    if (gemm(g)) { dimo::set_error("tn_gemm_kernel launch failed (data gradient, layer %d)", l); return -1; }
    if (l == 5) {                     // the embedding part of layer 5's input: plain, no mask
      TnGemmArgs e{};
      e.R = R; e.N = 128; e.nseg = 1;
      e.seg[0] = TnSeg{ws + o.g[5], 8, 0, ws + o.w_bwd5b, 8};
      e.plain = dcat; e.ldp = E; e.plain_cols = E; e.plain_acc = 0;
      if (gemm(e)) { dimo::set_error("tn_gemm_kernel launch failed (data gradient, layer 5 embedding part)"); return -1; }
    }
  }
  {                                   // layer 0: d(embedding) += dY_0 W_0
    TnGemmArgs e{};
    e.R = R; e.N = 128; e.nseg = 1;
    e.seg[0] = TnSeg{ws + o.g[0], 8, 0, ws + o.w_bwd[0], 8};
    e.plain = dcat; e.ldp = E; e.plain_cols = E; e.plain_acc = 1;
    if (gemm(e)) { dimo::set_error("tn_gemm_kernel launch failed (data gradient, layer 0)"); return -1; }
  }
  if (chained) {
    tn_chain_kernel<<<dim3(o.nrb, 4), TG_THREADS, TgCfg<64>::SMEM, st>>>(chain);
    DIMO_CHECK_LAUNCH();
  }
  g_tn_last_launches = 2 + (chained ? 1 : 10) + ((dpts != nullptr || dlatents != nullptr) ? 1 : 0);   // heads, GEMMs, wgrad, embed
  if (dpts != nullptr || dlatents != nullptr) {
    rc = dimo_timenet_embed_bwd(G, M, L, reinterpret_cast<float*>(ws + o.cat_plain), dcat, E, dpts, dlatents, stream);
    if (rc) return rc;
  }
  // ---- weight gradients of the ten 256-wide layers: one grouped launch over (job, row split) ----
  {
    TnWTable tab{};
    int n = 0;
    tab.splits = o.nrt >= 64 ? 8 : (o.nrt >= 16 ? 4 : (o.nrt >= 4 ? 2 : 1));
    tab.per = (o.nrt + tab.splits - 1) / tab.splits;
    tab.nrt = o.nrt;
    tab.det = dimo::det_scale();
    auto add = [&](const uint8_t* aT, int a_fb, const uint8_t* bT, int b_nfb, int b_fb0, int b_use, float* dW, int ldw,
                   int col0, int ncols, float* db) {
      TnWJob& j = tab.job[n++];
      j.aT = aT; j.a_nfb = 2; j.a_fb = a_fb; j.bT = bT; j.b_nfb = b_nfb; j.b_fb0 = b_fb0; j.b_use = b_use;
      j.dW = dW; j.ldw = ldw; j.col0 = col0; j.ncols = ncols; j.nrows = HID; j.db = db;
    };
    const int order[10] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 10};
    for (int oi = 0; oi < 10; ++oi) {
      const int l = order[oi], slot = tn_slot(l);
      for (int nb = 0; nb < 2; ++nb) {
        const uint8_t* aT = ws + o.gT[slot];
        if (l == 0) {
          add(aT, nb, ws + o.catT, 3, 0, 1, dW_host[0], E, 0, E, db_host[0]);
        } else if (l == 5) {
          add(aT, nb, ws + o.catT, 3, 0, 1, dW_host[5], E + HID, 0, E, db_host[5]);
          add(aT, nb, ws + o.catT, 3, 1, 2, dW_host[5], E + HID, E, HID, nullptr);
        } else {
          const int src = l == 10 ? 7 : l - 1;
          const uint8_t* bT = src == 4 ? ws + o.catT : ws + o.yT[tn_slot(src)];
          add(aT, nb, bT, src == 4 ? 3 : 2, src == 4 ? 1 : 0, 2, dW_host[l], HID, 0, HID, db_host[l]);
        }
      }
    }
    tab.n = n;
    tn_wgrad_kernel<<<n * tab.splits, TN_THREADS, TW_SMEM, st>>>(tab);
    DIMO_CHECK_LAUNCH();
  }
  return 0;
}
