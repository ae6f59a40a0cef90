// TimeNet linear layer on the 5th-generation tensor cores: Y = act(X * W^T + b) as a tcgen05 GEMM with
// 3xTF32 error compensation (renderer/latent_gs_renderer.py:223-232 runs these layers as cuBLAS FP32 SGEMM; the
// 1e-4 parity bound rules out plain TF32/BF16).
//
//   x = hi + lo  with hi = tf32(x), lo = tf32(x - hi)     (22 mantissa bits kept)
//   X*W^T ~= Xhi*Whi^T + Xhi*Wlo^T + Xlo*Whi^T            (three kind::tf32 MMAs per k-step, FP32 accumulate in TMEM)
//
// One CTA owns a 128-row output tile (64 columns, 256 threads for forward / data-gradient; 128 columns, 128 threads
// for the weight gradient):
//   * operands are split into hi/lo on the fly while they are copied global -> shared memory in the canonical
//     K-major no-swizzle UMMA layout (8-row x 16-byte core matrices, 128 B each; LBO = 128 B between the K-chunks of
//     a core-matrix row, SBO = 1024 B between 8-row groups);
//   * two shared-memory stages of 32 K-elements; one elected thread issues 4 k-steps x 3 tcgen05.mma (M=128,
//     K=8) per stage and commits them to an mbarrier, while every thread already holds the next K-tile in registers;
//   * the accumulator lives in TMEM; the epilogue reads it with tcgen05.ld (32 lanes x 32 columns per warp), adds the
//     bias, applies ReLU / the accumulate option and writes rows with 128-bit stores.
// The ReLU mask of the backward data-gradient (dX = (dY * [Y>0]) * W) is applied while loading the A operand.
#include "common.cuh"

namespace dimo {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32;
constexpr int TC_THREADS = 128;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;        // 16 KB (one of hi / lo)
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;        // 16 KB
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 64 KB
constexpr int TC_SMEM_BYTES = 2 * TC_STAGE_BYTES + 1024;          // + alignment slack

// debug knobs (dimo_tc_debug_set): 0 = swap LBO/SBO in the shared-memory descriptors, 1 = single-pass TF32 (no compensation)
static int h_tc_knob[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // [7]: ablation mask for timing experiments

struct TcArgs {
  int R, K, No;
  const float* X; int64_t ldx;
  const float* mask; int64_t ldm;     // optional: X(r,k) *= [mask(r,k) > 0]
  const float* Wt;                    // [No, K] row-major (K contiguous)
  const float* bias;
  float* Y; int64_t ldy;
  int relu, accumulate;
  int swap_lbo_sbo, single_pass;
  int ablate;      // timing experiments only (tools/linear_tc_bench.py): 1 = no global loads, 2 = no shared stores, 4 = no MMAs
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// byte offset of (row, 16-byte chunk) inside a [rows x 32 floats] canonical K-major tile
__device__ __forceinline__ uint32_t canon_off(int row, int chunk) {
  return (uint32_t)((((row >> 3) * (TC_BK / 4) + chunk) << 7) + ((row & 7) << 4));
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);               // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // leading dimension byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;     // stride dimension byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 (Blackwell), bits [46,48)
  // base_offset [49,52) = 0, lbo_mode [52] = 0, layout_type [61,64) = 0 (SWIZZLE_NONE)
  return d;
}

// instruction descriptor for kind::tf32, FP32 accumulate, both operands K-major
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(tc_smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// Forward / data-gradient kernel: 128-row x 64-column output tile, 256 threads (8 warps).
constexpr int LIN_BN = 64, LIN_BK = 64, LIN_THREADS = 256;     // 64 K-elements per stage: 8 k-steps, half the barrier rounds
constexpr int LIN_A_BYTES = TC_BM * LIN_BK * 4;                       // 16 KB (one of hi / lo)
constexpr int LIN_B_BYTES = LIN_BN * LIN_BK * 4;                      //  8 KB
constexpr int LIN_STAGE_BYTES = 2 * LIN_A_BYTES + 2 * LIN_B_BYTES;   // 48 KB
constexpr int LIN_SMEM_BYTES = 2 * LIN_STAGE_BYTES + 1024;
constexpr int LIN_NA = TC_BM * (LIN_BK / 4) / LIN_THREADS, LIN_NB = LIN_BN * (LIN_BK / 4) / LIN_THREADS;   // 16-byte chunks per thread

__device__ __forceinline__ uint32_t canon_off_n(int row, int chunk, int nchunks) {
  return (uint32_t)((((row >> 3) * nchunks + chunk) << 7) + ((row & 7) << 4));
}

__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, uint32_t off, const float4 v) {
  const uint4 hi = make_uint4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  const uint4 lo = make_uint4(to_tf32(v.x - __uint_as_float(hi.x)), to_tf32(v.y - __uint_as_float(hi.y)),
                              to_tf32(v.z - __uint_as_float(hi.z)), to_tf32(v.w - __uint_as_float(hi.w)));
  *reinterpret_cast<uint4*>(hi_base + off) = hi;
  *reinterpret_cast<uint4*>(lo_base + off) = lo;
}

__global__ void __launch_bounds__(LIN_THREADS, 1) linear_tc_kernel(TcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar[2];
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * LIN_BN;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)),
                 "r"((uint32_t)LIN_BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mma_bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mma_bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;

  const uint32_t idesc = make_idesc_tf32(TC_BM, LIN_BN);
  constexpr uint32_t LIN_SBO = (LIN_BK / 4) * 128;    // 8-row groups are LIN_BK/4 core matrices apart
  const uint32_t lbo = p.swap_lbo_sbo ? LIN_SBO : 128u, sbo = p.swap_lbo_sbo ? 128u : LIN_SBO;
  const int nk = (p.K + LIN_BK - 1) / LIN_BK;

  // Per-thread tile coordinates.  Quarter-warps write the 8 rows of ONE core matrix (128 contiguous bytes ->
  // conflict-free STS.128); neighbouring quarters take the adjacent 16-byte chunk of the same rows, so every global
  // request covers full 32-byte sectors.  A: 128 rows x 16 chunks = 8 per thread, B: 64 rows x 16 chunks = 4 per thread.
  int a_row[LIN_NA], a_chunk[LIN_NA], b_row[LIN_NB], b_chunk[LIN_NB];
#pragma unroll
  for (int it = 0; it < LIN_NA; ++it) {
    const int e = tid + it * LIN_THREADS;
    a_row[it] = (((e >> 4) & 15) << 3) | (e & 7);
    a_chunk[it] = ((e >> 8) << 1) | ((e >> 3) & 1);
  }
#pragma unroll
  for (int it = 0; it < LIN_NB; ++it) {
    const int e = tid + it * LIN_THREADS;
    b_row[it] = (((e >> 4) & 7) << 3) | (e & 7);
    b_chunk[it] = ((e >> 7) << 1) | ((e >> 3) & 1);
  }
  float4 va[LIN_NA], vb[LIN_NB];
  // the operands of K-tile kt+1 are fetched into registers while tile kt is converted, stored and multiplied
  auto fetch = [&](int kt) {
    const int k0 = kt * LIN_BK;
#pragma unroll
    for (int it = 0; it < LIN_NA; ++it) {
      const int gr = m0 + a_row[it], gk = k0 + a_chunk[it] * 4;
      va[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < p.R && gk < p.K) {
        va[it] = *reinterpret_cast<const float4*>(p.X + gr * p.ldx + gk);
        if (p.mask != nullptr) {
          const float4 m = *reinterpret_cast<const float4*>(p.mask + gr * p.ldm + gk);
          if (!(m.x > 0.f)) va[it].x = 0.f;
          if (!(m.y > 0.f)) va[it].y = 0.f;
          if (!(m.z > 0.f)) va[it].z = 0.f;
          if (!(m.w > 0.f)) va[it].w = 0.f;
        }
      }
    }
#pragma unroll
    for (int it = 0; it < LIN_NB; ++it) {
      const int gn = n0 + b_row[it], gk = k0 + b_chunk[it] * 4;
      vb[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gn < p.No && gk < p.K) vb[it] = *reinterpret_cast<const float4*>(p.Wt + (int64_t)gn * p.K + gk);
    }
  };

  if (!(p.ablate & 1)) fetch(0);
  else {
#pragma unroll
    for (int it = 0; it < LIN_NA; ++it) va[it] = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
    for (int it = 0; it < LIN_NB; ++it) vb[it] = make_float4(1.f, 2.f, 3.f, 4.f);
  }
  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt & 1;
    uint8_t* stage = smem + s * LIN_STAGE_BYTES;
    uint8_t* a_hi = stage, *a_lo = stage + LIN_A_BYTES, *b_hi = stage + 2 * LIN_A_BYTES,
             *b_lo = stage + 2 * LIN_A_BYTES + LIN_B_BYTES;
    if (kt >= 2) tc_mbar_wait(&mma_bar[s], (uint32_t)(((kt >> 1) - 1) & 1));   // MMAs of tile kt-2 have read this stage
    if (!(p.ablate & 2)) {
#pragma unroll
      for (int it = 0; it < LIN_NA; ++it) split_store(a_hi, a_lo, canon_off_n(a_row[it], a_chunk[it], LIN_BK / 4), va[it]);
#pragma unroll
      for (int it = 0; it < LIN_NB; ++it) split_store(b_hi, b_lo, canon_off_n(b_row[it], b_chunk[it], LIN_BK / 4), vb[it]);
    }
    if (kt + 1 < nk && !(p.ablate & 1)) fetch(kt + 1);          // in flight during the barrier, the MMA issue and the next wait
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async-proxy (MMA) reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa_hi = tc_smem_u32(a_hi), sa_lo = tc_smem_u32(a_lo), sb_hi = tc_smem_u32(b_hi),
                     sb_lo = tc_smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < LIN_BK / 8; ++ks) {
        const uint32_t koff = (uint32_t)ks * 256u;     // 8 tf32 = two 16-byte chunks = two 128-byte core-matrix blocks
        const uint64_t dah = make_smem_desc(sa_hi + koff, lbo, sbo), dal = make_smem_desc(sa_lo + koff, lbo, sbo);
        const uint64_t dbh = make_smem_desc(sb_hi + koff, lbo, sbo), dbl = make_smem_desc(sb_lo + koff, lbo, sbo);
        if (p.ablate & 4) continue;
        tc_mma(tmem_d, dah, dbh, idesc, (kt > 0 || ks > 0) ? 1u : 0u);
        if (!p.single_pass) {
          tc_mma(tmem_d, dah, dbl, idesc, 1u);
          tc_mma(tmem_d, dal, dbh, idesc, 1u);
        }
      }
      tc_commit(&mma_bar[s]);
    }
  }
  // all MMAs done <=> the last commit has arrived
  {
    const int last = nk - 1;
    tc_mbar_wait(&mma_bar[last & 1], (uint32_t)((last >> 1) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: warp w reads TMEM lanes [32 (w&3), +32) (its quarter) and columns [32 (w>>2), +32) ----
  const int row = m0 + (warp & 3) * 32 + lane;
  const int c0 = (warp >> 2) * 32;
  {
    uint32_t v[32];
    tc_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, v);
    if (row < p.R) {
      float* yrow = p.Y + row * p.ldy + n0 + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int gn = n0 + c0 + j;
        if (gn >= p.No) break;
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float val = __uint_as_float(v[j + q]);
          if (p.bias != nullptr && gn + q < p.No) val += p.bias[gn + q];
          if (p.relu) val = fmaxf(val, 0.f);
          o[q] = val;
        }
        const bool vec = (gn + 3 < p.No) && ((reinterpret_cast<uintptr_t>(yrow + j) & 15) == 0);
        if (vec) {
          float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.accumulate) prev = *reinterpret_cast<float4*>(yrow + j);
          *reinterpret_cast<float4*>(yrow + j) = make_float4(o[0] + prev.x, o[1] + prev.y, o[2] + prev.z, o[3] + prev.w);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (gn + q < p.No) yrow[j + q] = o[q] + (p.accumulate ? yrow[j + q] : 0.f);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)LIN_BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Weight gradient  dW[n,k] += sum_r (dY[r,n] * [Y[r,n] > 0]) * X[r,k],  db[n] += sum_r dY[r,n] * [Y[r,n] > 0]
// as D[M = n, N = k] with the reduction over rows r.  In memory both operands have the NON-reduction index
// contiguous (dY[r][n], X[r][k]); every thread transposes 4x4 blocks in registers while filling shared memory, so the
// tensor core sees the same K-major canonical layout as the forward kernel (reduction index r = UMMA K).
// grid = (ceil(No/128), ceil(K/128), row splits); partial tiles are added to dW with fp32 atomics.
// ---------------------------------------------------------------------------------------------------------------
struct TcWgradArgs {
  int R, K, No;
  const float* dY; int64_t lddy;
  const float* mask; int64_t ldm;
  const float* X; int64_t ldx;
  float* dW;   // [No, K]
  float* db;   // [No] or NULL
  int rows_per_split;
  int single_pass;
  int gx, gy;          // tiles along No and K
};

// Grouped launch: the weight gradients of all TimeNet layers are independent once the data-gradient chain has
// produced every dY, so they run as ONE launch (1-D grid over [problem][split][k-tile][n-tile]) instead of twelve
// sub-wave launches of ~128 CTAs each.
constexpr int TC_MAX_GROUP = 12;
struct TcWgradGroup {
  int n;
  int cta_start[TC_MAX_GROUP + 1];
  TcWgradArgs p[TC_MAX_GROUP];
};

__global__ void __launch_bounds__(TC_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ TcWgradGroup grp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t mma_bar[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_db[TC_BM];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int gi = 0;
  while (gi + 1 < grp.n && (int)blockIdx.x >= grp.cta_start[gi + 1]) ++gi;
  const TcWgradArgs& p = grp.p[gi];
  const int local = (int)blockIdx.x - grp.cta_start[gi];
  const int bx = local % p.gx, by = (local / p.gx) % p.gy, bz = local / (p.gx * p.gy);
  const int n0 = bx * TC_BM, k0 = by * TC_BN;
  const int r_begin = bz * p.rows_per_split;
  const int r_end = min(p.R, r_begin + p.rows_per_split);
  const bool do_db = p.db != nullptr && by == 0;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)),
                 "r"((uint32_t)TC_BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mma_bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&mma_bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  s_db[tid] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = make_idesc_tf32(TC_BM, TC_BN);
  const int nrows = max(0, r_end - r_begin);
  const int nk = (nrows + TC_BK - 1) / TC_BK;

  float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t lbo = 128u, sbo = 1024u;

  for (int kt = 0; kt < nk; ++kt) {
    const int s = kt & 1;
    uint8_t* stage = smem + s * TC_STAGE_BYTES;
    uint8_t* a_hi = stage, *a_lo = stage + TC_A_BYTES, *b_hi = stage + 2 * TC_A_BYTES,
             *b_lo = stage + 2 * TC_A_BYTES + TC_B_BYTES;
    if (kt >= 2) tc_mbar_wait(&mma_bar[s], (uint32_t)(((kt >> 1) - 1) & 1));
    const int r0 = r_begin + kt * TC_BK;
    // item = (column group u of 4 columns, row chunk c of 4 rows): 32 x 8 items, 2 per thread (same u, c and c+4)
    const int u = tid & 31;
    float4 va[2][4], vb[2][4];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c = (tid >> 5) + 4 * it;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gr = r0 + 4 * c + j;
        const int gn = n0 + 4 * u, gk = k0 + 4 * u;
        va[it][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        vb[it][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < r_end && gn < p.No) va[it][j] = *reinterpret_cast<const float4*>(p.dY + gr * p.lddy + gn);
        if (gr < r_end && gk < p.K) vb[it][j] = *reinterpret_cast<const float4*>(p.X + gr * p.ldx + gk);
      }
    }
    if (p.mask != nullptr) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int c = (tid >> 5) + 4 * it;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int gr = r0 + 4 * c + j, gn = n0 + 4 * u;
          if (gr < r_end && gn < p.No) {
            const float4 m = *reinterpret_cast<const float4*>(p.mask + gr * p.ldm + gn);
            if (!(m.x > 0.f)) va[it][j].x = 0.f;
            if (!(m.y > 0.f)) va[it][j].y = 0.f;
            if (!(m.z > 0.f)) va[it][j].z = 0.f;
            if (!(m.w > 0.f)) va[it][j].w = 0.f;
          }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c = (tid >> 5) + 4 * it;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dbacc.x += va[it][j].x; dbacc.y += va[it][j].y; dbacc.z += va[it][j].z; dbacc.w += va[it][j].w;
      }
      // transposed 4x4 blocks: output row (4u + i) holds rows r..r+3 of column i
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float a0 = i == 0 ? va[it][0].x : i == 1 ? va[it][0].y : i == 2 ? va[it][0].z : va[it][0].w;
        const float a1 = i == 0 ? va[it][1].x : i == 1 ? va[it][1].y : i == 2 ? va[it][1].z : va[it][1].w;
        const float a2 = i == 0 ? va[it][2].x : i == 1 ? va[it][2].y : i == 2 ? va[it][2].z : va[it][2].w;
        const float a3 = i == 0 ? va[it][3].x : i == 1 ? va[it][3].y : i == 2 ? va[it][3].z : va[it][3].w;
        const float b0 = i == 0 ? vb[it][0].x : i == 1 ? vb[it][0].y : i == 2 ? vb[it][0].z : vb[it][0].w;
        const float b1 = i == 0 ? vb[it][1].x : i == 1 ? vb[it][1].y : i == 2 ? vb[it][1].z : vb[it][1].w;
        const float b2 = i == 0 ? vb[it][2].x : i == 1 ? vb[it][2].y : i == 2 ? vb[it][2].z : vb[it][2].w;
        const float b3 = i == 0 ? vb[it][3].x : i == 1 ? vb[it][3].y : i == 2 ? vb[it][3].z : vb[it][3].w;
        const uint32_t off = canon_off(4 * u + i, c);
        const uint4 ah = make_uint4(to_tf32(a0), to_tf32(a1), to_tf32(a2), to_tf32(a3));
        const uint4 al = make_uint4(to_tf32(a0 - __uint_as_float(ah.x)), to_tf32(a1 - __uint_as_float(ah.y)),
                                    to_tf32(a2 - __uint_as_float(ah.z)), to_tf32(a3 - __uint_as_float(ah.w)));
        const uint4 bh = make_uint4(to_tf32(b0), to_tf32(b1), to_tf32(b2), to_tf32(b3));
        const uint4 bl = make_uint4(to_tf32(b0 - __uint_as_float(bh.x)), to_tf32(b1 - __uint_as_float(bh.y)),
                                    to_tf32(b2 - __uint_as_float(bh.z)), to_tf32(b3 - __uint_as_float(bh.w)));
        *reinterpret_cast<uint4*>(a_hi + off) = ah;
        *reinterpret_cast<uint4*>(a_lo + off) = al;
        *reinterpret_cast<uint4*>(b_hi + off) = bh;
        *reinterpret_cast<uint4*>(b_lo + off) = bl;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sa_hi = tc_smem_u32(a_hi), sa_lo = tc_smem_u32(a_lo), sb_hi = tc_smem_u32(b_hi),
                     sb_lo = tc_smem_u32(b_lo);
#pragma unroll
      for (int ks = 0; ks < TC_BK / 8; ++ks) {
        const uint32_t koff = (uint32_t)ks * 256u;
        const uint64_t dah = make_smem_desc(sa_hi + koff, lbo, sbo), dal = make_smem_desc(sa_lo + koff, lbo, sbo);
        const uint64_t dbh = make_smem_desc(sb_hi + koff, lbo, sbo), dbl = make_smem_desc(sb_lo + koff, lbo, sbo);
        tc_mma(tmem_d, dah, dbh, idesc, (kt > 0 || ks > 0) ? 1u : 0u);
        if (!p.single_pass) {
          tc_mma(tmem_d, dah, dbl, idesc, 1u);
          tc_mma(tmem_d, dal, dbh, idesc, 1u);
        }
      }
      tc_commit(&mma_bar[s]);
    }
  }
  if (nk > 0) {
    const int last = nk - 1;
    tc_mbar_wait(&mma_bar[last & 1], (uint32_t)((last >> 1) & 1));
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (nk > 0) {
    const int nrow = n0 + warp * 32 + lane;          // output row = weight row n
#pragma unroll 1
    for (int c0 = 0; c0 < TC_BN; c0 += 32) {
      uint32_t v[32];
      tc_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (nrow < p.No) {
        float* wrow = p.dW + (int64_t)nrow * p.K + k0 + c0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (k0 + c0 + j < p.K) atomicAdd(wrow + j, __uint_as_float(v[j]));
      }
    }
    if (do_db) {
      const int u = tid & 31;
      atomicAdd(&s_db[4 * u + 0], dbacc.x); atomicAdd(&s_db[4 * u + 1], dbacc.y);
      atomicAdd(&s_db[4 * u + 2], dbacc.z); atomicAdd(&s_db[4 * u + 3], dbacc.w);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (do_db && nk > 0 && n0 + tid < p.No) atomicAdd(p.db + n0 + tid, s_db[tid]);
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)TC_BN) : "memory");
  }
}

}  // namespace dimo

using namespace dimo;

namespace dimo { extern int g_blend_gather_mode; extern int g_blend_bwd_chunk; extern int g_blend_fwd_chunk; }   // raster_blend.cu
namespace dimo { extern int g_disable_packed_instances; }                           // raster_bin.cu

extern "C" int dimo_tc_debug_set(int key, int value) {
  if (key == 6) {
    dimo::g_disable_packed_instances = value != 0;
    return 0;
  }
  if (key == 7) {
    h_tc_knob[7] = value;
    return 0;
  }
  if (key == 5) {
    if (value != 64 && value != 128) return -2;
    dimo::g_blend_fwd_chunk = value;
    return 0;
  }
  if (key < 0 || key >= 5) return -2;
  if (key == 4) {
    if (value != 64 && value != 128) return -2;
    dimo::g_blend_bwd_chunk = value;
    return 0;
  }
  h_tc_knob[key] = value;
  if (key == 3) dimo::g_blend_gather_mode = value != 0;
  return 0;
}

extern "C" int dimo_linear_tc(int R, int K, int No, const float* X, int64_t ldx, const float* mask, int64_t ldm,
                              const float* Wt, const float* bias, float* Y, int64_t ldy, int relu, int accumulate,
                              void* stream) {
  if (R == 0 || No == 0) return 0;
  DIMO_REQUIRE(K % 4 == 0 && ldx % 4 == 0, "tensor-core linear: K and ldx must be multiples of 4 floats");
  DIMO_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Wt) & 15) == 0,
               "tensor-core linear: X and W must be 16-byte aligned");
  DIMO_REQUIRE(mask == nullptr || (ldm % 4 == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0),
               "tensor-core linear: mask must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LIN_SMEM_BYTES));
    attr_set = true;
  }
  TcArgs p{};
  p.R = R; p.K = K; p.No = No; p.X = X; p.ldx = ldx; p.mask = mask; p.ldm = ldm; p.Wt = Wt; p.bias = bias;
  p.Y = Y; p.ldy = ldy; p.relu = relu; p.accumulate = accumulate;
  p.swap_lbo_sbo = h_tc_knob[0]; p.single_pass = h_tc_knob[1]; p.ablate = h_tc_knob[7];
  dim3 grid(ceil_div(R, TC_BM), ceil_div(No, LIN_BN));
  linear_tc_kernel<<<grid, LIN_THREADS, LIN_SMEM_BYTES, (cudaStream_t)stream>>>(p);
  DIMO_CHECK_LAUNCH();
  return 0;
}

static int wgrad_fill(TcWgradArgs& p, int R, int K, int No, const float* dY, int64_t lddy, const float* mask,
                      int64_t ldm, const float* X, int64_t ldx, float* dW, float* db, int target_ctas) {
  DIMO_REQUIRE(No % 4 == 0 && K % 4 == 0 && lddy % 4 == 0 && ldx % 4 == 0,
               "tensor-core wgrad: No, K, lddy, ldx must be multiples of 4 floats");
  DIMO_REQUIRE((reinterpret_cast<uintptr_t>(dY) & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
               "tensor-core wgrad: dY and X must be 16-byte aligned");
  DIMO_REQUIRE(mask == nullptr || (ldm % 4 == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0),
               "tensor-core wgrad: mask must be 16-byte aligned");
  p = TcWgradArgs{};
  p.R = R; p.K = K; p.No = No; p.dY = dY; p.lddy = lddy; p.mask = mask; p.ldm = ldm; p.X = X; p.ldx = ldx;
  p.dW = dW; p.db = db; p.single_pass = h_tc_knob[1];
  p.gx = ceil_div(No, TC_BM); p.gy = ceil_div(K, TC_BN);
  const int tiles = p.gx * p.gy;
  int splits = max(1, min(ceil_div(R, 2 * TC_BK), ceil_div(target_ctas, tiles)));
  p.rows_per_split = ceil_div(ceil_div(R, splits), TC_BK) * TC_BK;
  splits = ceil_div(R, p.rows_per_split);
  return tiles * splits;       // CTAs of this problem
}

static int wgrad_launch(const TcWgradGroup& grp, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    attr_set = true;
  }
  const int total = grp.cta_start[grp.n];
  if (total == 0) return 0;
  wgrad_tc_kernel<<<total, TC_THREADS, TC_SMEM_BYTES, st>>>(grp);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_linear_wgrad_tc(int R, int K, int No, const float* dY, int64_t lddy, const float* mask, int64_t ldm,
                                    const float* X, int64_t ldx, float* dW, float* db, void* stream) {
  if (R == 0 || No == 0 || K == 0) return 0;
  TcWgradGroup grp{};
  grp.n = 1;
  const int target_ctas = h_tc_knob[2] > 0 ? h_tc_knob[2] : 148;     // knob 2: CTA budget for the row split (tuning)
  const int ctas = wgrad_fill(grp.p[0], R, K, No, dY, lddy, mask, ldm, X, ldx, dW, db, target_ctas);
  if (ctas < 0) return ctas;
  grp.cta_start[0] = 0; grp.cta_start[1] = ctas;
  return wgrad_launch(grp, (cudaStream_t)stream);
}

extern "C" int dimo_linear_wgrad_tc_grouped(int n, int R, const int* K, const int* No, const float* const* dY,
                                            const int64_t* lddy, const float* const* mask, const int64_t* ldm,
                                            const float* const* X, const int64_t* ldx, float* const* dW,
                                            float* const* db, void* stream) {
  DIMO_REQUIRE(n >= 0 && n <= TC_MAX_GROUP, "at most 12 problems per grouped weight-gradient launch");
  if (n == 0 || R == 0) return 0;
  TcWgradGroup grp{};
  grp.n = n;
  // every problem gets enough row splits to keep ~2 waves of CTAs in flight over the whole group
  const int target_ctas = h_tc_knob[2] > 0 ? h_tc_knob[2] : max(32, 2 * 148 / n);
  int start = 0;
  for (int i = 0; i < n; ++i) {
    grp.cta_start[i] = start;
    const int ctas = wgrad_fill(grp.p[i], R, K[i], No[i], dY[i], lddy[i], mask[i], ldm[i], X[i], ldx[i], dW[i], db[i],
                                target_ctas);
    if (ctas < 0) return ctas;
    start += ctas;
  }
  grp.cta_start[n] = start;
  return wgrad_launch(grp, (cudaStream_t)stream);
}
