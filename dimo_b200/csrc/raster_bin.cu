// Binning stage of the tile rasteriser: depth order of the splats, per-tile instance lists, per-tile ranges.
//
// Upstream shape (diff_gauss / diff_gaussian_rasterization, call site
// renderer/latent_gs_renderer.py:1256-1277): InclusiveSum -> duplicateWithKeys -> SortPairs over
// 64-bit (tile | depth) keys (6 radix passes at 512^2 x 16 frames = 144 B per instance) -> identifyTileRanges.
//
// Here the same total order -- per tile: ascending view depth, ties in ascending Gaussian index -- comes from two
// hand-written stages that never materialise a (tile | depth) key:
//
//   1. depth_sort_kernel   the N splats of every frame are sorted by their 32-bit depth bits (culled ones last) by
//      ONE kernel: a thread-block CLUSTER of 8 CTAs per frame runs the four 8-bit passes of a stable LSD radix sort;
//      digit counts are exchanged through distributed shared memory and the passes are separated by cluster barriers,
//      so there are no per-pass launches, no global histograms and no look-back spinning.  Inside a warp, ranks
//      come from match.any (lanes that hold the same digit) -- no shared-memory atomics.
//   2. a STABLE COUNTING SORT of the instances by tile, one pass, exploiting that the depth order is frame-major
//      (instances of frame b can only land in tiles of frame b):
//        tile_hist_kernel      per chunk of consecutive splats: how many of them touch each tile -- a 2-D difference
//                              array (4 shared-memory increments per splat) + prefix sums, no per-instance work;
//        tile_chunk_scan / tile_base_scan   exclusive prefix over chunks per tile, then over tiles: every tile's
//                              [begin, end) range and every chunk's first slot in every tile (+ overflow flag);
//        tile_scatter_kernel   every warp walks its splats in depth order; the lanes take the tiles of one splat's
//                              rectangle and write each instance to its tile's next free slot (per-warp cursors).
//      HBM: 8 B read per splat (tile rectangle) twice + 4 B written per instance; the 2-pass radix sort it replaces
//      wrote the unsorted instances, read them for the histogram and moved them twice (28 B per instance), plus a
//      4 B sentinel fill and a 4 B read for the ranges.
// The blend kernels gather the 64-byte blend records of a tile's list by index (`vals_sorted` -> record table
// written by the preprocess kernel).  Integer-only, bit-exact against the oracle's stable 64-bit sort.
#include "common.cuh"
#include <cooperative_groups.h>
#include <stdarg.h>

namespace cg = cooperative_groups;

namespace dimo {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static float g_det_scale = 0.f;
float det_scale() { return g_det_scale; }

__global__ void __launch_bounds__(256) fixed_to_float_kernel(int64_t n, const long long* __restrict__ src,
                                                             float* __restrict__ dst, int accumulate, double inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = (float)((double)src[i] * inv);
  dst[i] = accumulate ? dst[i] + v : v;
}

int g_disable_packed_instances = 0;     // dimo_tc_debug_set key 6: force the (key, value) pair format (tests, A/B)

static inline int bits_for(int64_t n) {
  int bits = 1;
  while (((int64_t)1 << bits) < n) ++bits;
  return bits;
}

// Lanes that hold the same `key` (its low `nbits` bits) among the lanes with `valid`: one ballot per bit.  (The
// hardware MATCH.ANY serialises on a slow unit -- ~900 cycles per use with 16 resident warps, profiles/r2e -- while
// ballots issue at full rate.)  Lanes with !valid get an empty set.
__device__ __forceinline__ uint32_t warp_match(uint32_t key, int nbits, bool valid) {
  uint32_t peers = __ballot_sync(0xffffffffu, valid);
  for (int bit = 0; bit < nbits; ++bit) {
    const bool p = (key >> bit) & 1u;
    const uint32_t b = __ballot_sync(0xffffffffu, p);
    peers &= p ? b : ~b;
  }
  return valid ? peers : 0u;
}

// ---------------------------------------------------------------------------------------------------------------
// 1. per-frame depth sort
// ---------------------------------------------------------------------------------------------------------------
constexpr int DS_CL = 8;              // CTAs per cluster = per frame
constexpr int DS_BINS = 256;
constexpr int DS_BATCH = 8;           // keys in flight per lane (independent loads)

// keys0 [B*N] (input, depth bits); kB, kC, vB, vC [B*N] ping-pong buffers.  After the four passes vC holds, per
// frame, the indices b*N + i in ascending (depth, i) order.
// DS_THREADS = 1024: one CTA per SM, at most 15 clusters in flight on a B200 (measured, dimo_debug_max_sort_clusters);
// 512: two CTAs per SM -- chosen when more frames than that have to be sorted at once, so that no frame waits for a
// second wave.
template <int DS_THREADS>
__global__ void __cluster_dims__(DS_CL, 1, 1) __launch_bounds__(DS_THREADS, DS_THREADS == 1024 ? 1 : 2)
depth_sort_kernel(int N, const uint32_t* __restrict__ keys0, uint32_t* __restrict__ kB, uint32_t* __restrict__ kC,
                  uint32_t* __restrict__ vB, uint32_t* __restrict__ vC) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int b = blockIdx.x / DS_CL;
  constexpr int DS_WARPS = DS_THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ uint32_t wc[DS_WARPS][DS_BINS];     // per-warp digit counts, then per-warp cursors
  __shared__ uint32_t cta_tot[DS_BINS];          // this CTA's digit counts (read by the whole cluster)
  __shared__ uint32_t scan[DS_BINS];
  const int64_t fbase = (int64_t)b * N;
  // contiguous ranges: CTA `crank` of the frame, warp `warp` of the CTA (multiples of 32 keep the loads aligned)
  const int per_cta = ((N + DS_CL - 1) / DS_CL + 31) & ~31;
  const int clo = min(N, crank * per_cta), chi = min(N, clo + per_cta);
  const int per_warp = (((chi - clo) + DS_WARPS - 1) / DS_WARPS + 31) & ~31;
  const int wlo = min(chi, clo + warp * per_warp), whi = min(chi, wlo + per_warp);
  const uint32_t lt = (1u << lane) - 1u;

  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 8 * pass;
    const uint32_t* in_k = pass == 0 ? keys0 : ((pass & 1) ? kB : kC);
    const uint32_t* in_v = (pass & 1) ? vB : vC;
    uint32_t* out_k = (pass & 1) ? kC : kB;
    uint32_t* out_v = (pass & 1) ? vC : vB;
    for (int k = tid; k < DS_WARPS * DS_BINS; k += DS_THREADS) (&wc[0][0])[k] = 0;
    __syncthreads();
    // ---- count: native shared-memory integer atomics on the warp's own counters (ranks are not needed yet) ----
    for (int i0 = wlo; i0 < whi; i0 += 32 * DS_BATCH) {
      uint32_t key[DS_BATCH];
#pragma unroll
      for (int u = 0; u < DS_BATCH; ++u) {
        const int i = i0 + 32 * u + lane;
        key[u] = i < whi ? in_k[fbase + i] : 0u;
      }
#pragma unroll
      for (int u = 0; u < DS_BATCH; ++u)
        if (i0 + 32 * u + lane < whi) atomicAdd(&wc[warp][(key[u] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < DS_BINS) {                          // exclusive prefix over the CTA's warps, CTA total per digit
      uint32_t run = 0;
#pragma unroll
      for (int w = 0; w < DS_WARPS; ++w) { const uint32_t c = wc[w][tid]; wc[w][tid] = run; run += c; }
      cta_tot[tid] = run;
    }
    cluster.sync();                               // every CTA's totals are published
    uint32_t lower = 0, all = 0;
    if (tid < DS_BINS) {
#pragma unroll
      for (int c = 0; c < DS_CL; ++c) {
        const uint32_t t = *cluster.map_shared_rank(&cta_tot[tid], c);
        all += t;
        if (c < crank) lower += t;
      }
      scan[tid] = all;
    }
    __syncthreads();
    for (int off = 1; off < DS_BINS; off <<= 1) {  // inclusive Hillis-Steele scan over the 256 digit totals
      uint32_t v = 0;
      if (tid < DS_BINS && tid >= off) v = scan[tid - off];
      __syncthreads();
      if (tid < DS_BINS) scan[tid] += v;
      __syncthreads();
    }
    if (tid < DS_BINS) {
      const uint32_t base = scan[tid] - all + lower;   // keys with a smaller digit + same digit in earlier CTAs
#pragma unroll
      for (int w = 0; w < DS_WARPS; ++w) wc[w][tid] += base;
    }
    __syncthreads();
    // ---- scatter: rank among the lanes with the same digit = position after the warp's cursor ----
    for (int i0 = wlo; i0 < whi; i0 += 32 * DS_BATCH) {
      uint32_t key[DS_BATCH], val[DS_BATCH];
#pragma unroll
      for (int u = 0; u < DS_BATCH; ++u) {
        const int i = i0 + 32 * u + lane;
        key[u] = i < whi ? in_k[fbase + i] : 0u;
        val[u] = i < whi ? (pass == 0 ? (uint32_t)(fbase + i) : in_v[fbase + i]) : 0u;
      }
#pragma unroll
      for (int u = 0; u < DS_BATCH; ++u) {
        if (i0 + 32 * u >= whi) break;
        const bool valid = i0 + 32 * u + lane < whi;
        const uint32_t d = (key[u] >> shift) & 255u;
        const uint32_t peers = warp_match(d, 8, valid);
        const uint32_t pos = valid ? wc[warp][d] + __popc(peers & lt) : 0u;
        __syncwarp();
        if (valid && (peers & lt) == 0) wc[warp][d] += __popc(peers);
        __syncwarp();
        if (valid) {
          if (pass < 3) out_k[fbase + pos] = key[u];
          out_v[fbase + pos] = val[u];
        }
      }
    }
    cluster.sync();   // the pass's writes are visible to the whole cluster; cta_tot may be overwritten again
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. stable counting sort of the instances by tile
// ---------------------------------------------------------------------------------------------------------------
// Chunk c of frame b = sorted positions [c * per_chunk, (c+1) * per_chunk) of that frame (same split in the
// histogram and the scatter kernel).
struct BinGeom {
  int B, N, gx, gy, T, nchunk, per_chunk, wpc;   // T = gx * gy tiles per frame; wpc = warps per scatter CTA
};

// counts[(y, x)] over a (gy+1) x (gx+1) difference array: += the number of rectangles covering tile (x, y)
__device__ __forceinline__ void diff_add(int* diff, int gw, uint2 r) {
  const int x0 = r.x & 0xFFFF, y0 = r.x >> 16, x1 = r.y & 0xFFFF, y1 = r.y >> 16;
  if (x1 <= x0 || y1 <= y0) return;
  atomicAdd(&diff[y0 * gw + x0], 1);
  atomicAdd(&diff[y0 * gw + x1], -1);
  atomicAdd(&diff[y1 * gw + x0], -1);
  atomicAdd(&diff[y1 * gw + x1], 1);
}

// in-place 2-D inclusive prefix of a (gy+1) x (gx+1) array by `nthr` cooperating threads (thread `t` of them);
// the caller synchronises the group before and after and between the two phases through `sync`
template <typename Sync>
__device__ __forceinline__ void prefix2d(int* a, int gw, int gh, int t, int nthr, Sync sync) {
  for (int y = t; y < gh; y += nthr) {
    int run = 0;
    for (int x = 0; x < gw; ++x) { run += a[y * gw + x]; a[y * gw + x] = run; }
  }
  sync();
  for (int x = t; x < gw; x += nthr) {
    int run = 0;
    for (int y = 0; y < gh; ++y) { run += a[y * gw + x]; a[y * gw + x] = run; }
  }
  sync();
}

// also writes the rectangles in sorted order (rects_sorted [B*N]) so that the scatter kernel reads them coalesced
__global__ void __launch_bounds__(256) tile_hist_kernel(BinGeom g, const uint2* __restrict__ rects,
                                                        const uint32_t* __restrict__ perm,
                                                        uint32_t* __restrict__ hist, uint2* __restrict__ rects_sorted) {
  extern __shared__ int diff[];                  // (gy+1) x (gx+1)
  const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int gw = g.gx + 1, gh = g.gy + 1;
  for (int k = tid; k < gw * gh; k += blockDim.x) diff[k] = 0;
  __syncthreads();
  const int lo = min(g.N, chunk * g.per_chunk), hi = min(g.N, lo + g.per_chunk);
  const int64_t fbase = (int64_t)b * g.N;
  for (int i = lo + tid; i < hi; i += blockDim.x) {
    const uint2 r = rects[perm[fbase + i]];
    rects_sorted[fbase + i] = r;
    diff_add(diff, gw, r);
  }
  __syncthreads();
  prefix2d(diff, gw, gh, tid, (int)blockDim.x, [] { __syncthreads(); });
  uint32_t* out = hist + ((int64_t)b * g.nchunk + chunk) * g.T;
  for (int t = tid; t < g.T; t += blockDim.x) {
    const int y = t / g.gx, x = t - y * g.gx;
    out[t] = (uint32_t)diff[y * gw + x];
  }
}

// hist[b][c][t] -> exclusive prefix over c (in place); tile_total[b*T + t] = sum over c
__global__ void __launch_bounds__(256) tile_chunk_scan_kernel(BinGeom g, uint32_t* __restrict__ hist,
                                                              uint32_t* __restrict__ tile_total) {
  const int64_t bt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bt >= (int64_t)g.B * g.T) return;
  const int b = (int)(bt / g.T), t = (int)(bt - (int64_t)b * g.T);
  uint32_t* col = hist + (int64_t)b * g.nchunk * g.T + t;
  uint32_t run = 0;
  for (int c0 = 0; c0 < g.nchunk; c0 += 8) {        // 8 independent loads in flight
    uint32_t v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = c0 + u < g.nchunk ? col[(int64_t)(c0 + u) * g.T] : 0u;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (c0 + u < g.nchunk) col[(int64_t)(c0 + u) * g.T] = run;
      run += v[u];
    }
  }
  tile_total[bt] = run;
}

// exclusive scan over the B*T tile totals (frame-major, tile order) by ONE CTA, 4096 tiles per round (one 128-bit
// load per thread): ranges[tile] = [begin, end), clamped to `slots`; count_overflow (may be NULL): [0] = true
// instance count, [1] |= count > slots
constexpr int TS_THREADS = 1024;
__global__ void __launch_bounds__(TS_THREADS) tile_base_scan_kernel(int64_t n, int64_t slots,
                                                                    const uint32_t* __restrict__ tile_total,
                                                                    uint32_t* __restrict__ tile_base,
                                                                    uint2* __restrict__ ranges,
                                                                    int32_t* __restrict__ count_overflow) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t cap = (uint32_t)min(slots, (int64_t)0xFFFFFFFFll);
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 4 * TS_THREADS) {
    const int64_t k0 = base + 4 * (int64_t)tid;
    uint32_t c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c[u] = k0 + u < n ? tile_total[k0 + u] : 0u;
    const uint32_t mine = c[0] + c[1] + c[2] + c[3];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = wsum[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += v;
      }
      wsum[lane] = wi - w;                           // exclusive prefix of the warp sums
    }
    __syncthreads();
    const uint32_t carry = carry_s;
    uint32_t run = carry + wsum[warp] + incl - mine;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (k0 + u < n) {
        tile_base[k0 + u] = run;
        ranges[k0 + u] = make_uint2(min(run, cap), min(run + c[u], cap));
      }
      run += c[u];
    }
    __syncthreads();
    if (tid == TS_THREADS - 1) carry_s = run;        // the round's total
    __syncthreads();
  }
  if (tid == 0 && count_overflow != nullptr) {
    count_overflow[0] = (int32_t)carry_s;
    if ((int64_t)carry_s > slots) count_overflow[1] = 1;
  }
}

// Every warp of a CTA owns a contiguous run of the chunk's splats (depth order) and walks their instances in order.
//   cnt[w][(gy+1) x (gx+1)]: difference array -> per-tile count of warp w -> warp w's next free slot per tile.
template <bool PACKED>
__global__ void __launch_bounds__(256) tile_scatter_kernel(BinGeom g, int64_t slots, int vbits,
                                                           const uint2* __restrict__ rects_sorted,
                                                           const uint32_t* __restrict__ perm,
                                                           const uint32_t* __restrict__ hist,
                                                           const uint32_t* __restrict__ tile_base, int tile_bits,
                                                           uint32_t* __restrict__ keys_out,
                                                           uint32_t* __restrict__ vals_out) {
  extern __shared__ int cnt_all[];
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gw = g.gx + 1, gh = g.gy + 1, stride = gw * gh;
  int* cnt = cnt_all + warp * stride;
  for (int k = tid; k < g.wpc * stride; k += blockDim.x) cnt_all[k] = 0;
  __syncthreads();
  const int clo = min(g.N, chunk * g.per_chunk), chi = min(g.N, clo + g.per_chunk);
  const int per_warp = (((chi - clo) + g.wpc - 1) / g.wpc + 31) & ~31;
  const int wlo = min(chi, clo + warp * per_warp), whi = min(chi, wlo + per_warp);
  const int64_t fbase = (int64_t)b * g.N;
  // ---- per-warp tile counts ----
  for (int i = wlo + lane; i < whi; i += 32) diff_add(cnt, gw, rects_sorted[fbase + i]);
  __syncwarp();
  prefix2d(cnt, gw, gh, lane, 32, [] { __syncwarp(); });
  __syncthreads();
  // ---- cursors: tile base + earlier chunks + earlier warps of this chunk ----
  const uint32_t* hx = hist + ((int64_t)b * g.nchunk + chunk) * g.T;
  const uint32_t* tb = tile_base + (int64_t)b * g.T;
  for (int t = tid; t < g.T; t += blockDim.x) {
    const int y = t / g.gx, x = t - y * g.gx;
    uint32_t run = tb[t] + hx[t];
    for (int w = 0; w < g.wpc; ++w) {
      int* c = cnt_all + w * stride + y * gw + x;
      const uint32_t v = (uint32_t)*c;
      *c = (int)run;
      run += v;
    }
  }
  __syncthreads();
  // ---- scatter: the warp's splats one after the other in depth order; the lanes take the tiles of the current
  // splat's rectangle (distinct cells, so the cursor update needs no atomics and no lane matching), and the per-tile
  // order is the order in which the splats are visited.  (An earlier version packed the instance stream into full
  // 32-lane windows and recovered the per-tile ranks with ballot matching: 230 instructions per window, profiles/r2u;
  // this loop needs ~35 per non-empty splat, i.e. less than half per instance at ~10 tiles per splat.) ----
  const uint32_t key_base = (uint32_t)b * (uint32_t)g.T;
  uint32_t nidx = 0;                                 // next group's splat (prefetched while this group is walked)
  uint2 nrect = make_uint2(0u, 0u);
  if (wlo + lane < whi) { nidx = perm[fbase + wlo + lane]; nrect = rects_sorted[fbase + wlo + lane]; }
  const uint32_t cap = (uint32_t)(slots < (int64_t)0xFFFFFFFFll ? slots : (int64_t)0xFFFFFFFFll);
  for (int g0 = wlo; g0 < whi; g0 += 32) {
    const int i = g0 + lane;
    // per lane, for its own splat: origin (x0 | y0 << 16), width | count << 9, 1 / width -- shuffled to the whole warp
    // when the splat's turn comes (width <= 511 tiles, count < 2^23: images up to 8176 pixels wide)
    uint32_t idx = nidx, rx = 0, wn = 0;
    float inv_w = 0.f;
    const uint2 r = nrect;
    if (i + 32 < whi) { nidx = perm[fbase + i + 32]; nrect = rects_sorted[fbase + i + 32]; }
    if (i < whi) {
      const uint32_t x0 = r.x & 0xFFFF, y0 = r.x >> 16, x1 = r.y & 0xFFFF, y1 = r.y >> 16;
      if (x1 > x0 && y1 > y0) {
        rx = r.x; wn = (x1 - x0) | (((x1 - x0) * (y1 - y0)) << 9);
        inv_w = 1.0f / (float)(x1 - x0);
        if (PACKED) idx -= (uint32_t)fbase;          // index within the frame
      }
    }
    uint32_t todo = __ballot_sync(0xffffffffu, wn != 0u);
    while (todo) {                                   // warp-uniform: the non-empty splats of the group, in order
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t s_rx = __shfl_sync(0xffffffffu, rx, src);
      const uint32_t s_wn = __shfl_sync(0xffffffffu, wn, src);
      const uint32_t s_idx = __shfl_sync(0xffffffffu, idx, src);
      const float s_inv = __shfl_sync(0xffffffffu, inv_w, src);
      const uint32_t w = s_wn & 511u, n = s_wn >> 9;
      const uint32_t bx = s_rx & 0xFFFF, by = s_rx >> 16;
      const uint32_t cell0 = by * (uint32_t)gw + bx, key0 = key_base + by * (uint32_t)g.gx + bx;
      // k -> (row, col) = (k / w, k % w): (k + 0.5) * (1 / w) truncates to the exact quotient while n <= 4096
      // (|error| <= 4096 * 2^-22 < 0.5 / 64 <= distance of (k + 0.5) / w to the next integer); bigger rectangles divide
      const bool small = n <= 4096u && w <= 64u;     // warp-uniform
      for (uint32_t k = lane; k < n; k += 32) {
        const uint32_t row = small ? (uint32_t)(((float)k + 0.5f) * s_inv) : k / w;
        const uint32_t col = k - row * w;
        const uint32_t cell = cell0 + row * (uint32_t)gw + col;
        const uint32_t slot = (uint32_t)cnt[cell];
        cnt[cell] = (int)(slot + 1u);
        if (slot < cap) {
          const uint32_t key = key0 + row * (uint32_t)g.gx + col;
          if (PACKED) {
            vals_out[slot] = (key << vbits) | s_idx;
          } else {
            keys_out[slot] = key;
            vals_out[slot] = s_idx;
          }
        }
      }
      __syncwarp();                                  // the next splat reads the cursors this one advanced
    }
  }
}


// launch geometry of the counting sort for a launch set
static inline BinGeom bin_geom(int B, int N, int W, int H) {
  BinGeom g;
  g.B = B; g.N = N;
  g.gx = (W + TILE - 1) / TILE; g.gy = (H + TILE - 1) / TILE;
  g.T = g.gx * g.gy;
  int nchunk = (6 * 148) / (B > 0 ? B : 1);            // <= 6 resident scatter CTAs per SM: rounded DOWN, so that the
                                                       // B x nchunk CTAs are one wave (rounding up left 8 CTAs of a
                                                       // 896-CTA grid for a second wave that doubled the kernel's time)
  nchunk = nchunk < 4 ? 4 : (nchunk > 128 ? 128 : nchunk);
  const int by_size = (N + 255) / 256;                 // at least 256 splats per chunk
  if (nchunk > by_size) nchunk = by_size < 1 ? 1 : by_size;
  g.nchunk = nchunk;
  g.per_chunk = (((N + nchunk - 1) / nchunk) + 31) & ~31;
  const int64_t cell_bytes = (int64_t)(g.gx + 1) * (g.gy + 1) * 4;
  int wpc = (int)(96 * 1024 / cell_bytes);
  g.wpc = wpc > 8 ? 8 : wpc;                            // < 1: rejected by dimo_raster_bin
  return g;
}

}  // namespace dimo

using namespace dimo;

namespace dimo {
int preprocess_launch(int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags,
                      const float* cams, const int32_t* frame_src, const float* means3D, int64_t means3D_bstride,
                      const float* scales,
                      int64_t scales_bstride, const float* rotations, int64_t rotations_bstride,
                      const float* opacities, int64_t opacities_bstride, const float* shs, int64_t shs_bstride,
                      const float* colors_precomp, int64_t colors_bstride, float* splats, int32_t* radii,
                      uint32_t* tiles_touched, uint32_t* rects, uint32_t* depth_keys,
                      unsigned long long* total_count, cudaStream_t st);
}  // namespace dimo

extern "C" {

int dimo_abi_version(void) { return DIMO_ABI_VERSION; }

/* debugging: how many depth-sort clusters (8 CTAs x 1024 threads) the device can hold at once */
int dimo_debug_max_sort_clusters(void) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(DS_CL * 64); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = DS_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = -1;
  if (cudaOccupancyMaxActiveClusters(&n, depth_sort_kernel<1024>, &cfg) != cudaSuccess) { cudaGetLastError(); return -1; }
  return n;
}

int dimo_set_deterministic(int on) {
  dimo::g_det_scale = on ? DET_SCALE : 0.f;
  return 0;
}
int dimo_get_deterministic(void) { return dimo::g_det_scale != 0.f; }

int dimo_fixed_to_float(int64_t n, const void* src_i64, float* dst, int accumulate, void* stream) {
  if (n <= 0) return 0;
  fixed_to_float_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      n, reinterpret_cast<const long long*>(src_i64), dst, accumulate, 1.0 / (double)DET_SCALE);
  DIMO_CHECK_LAUNCH();
  return 0;
}
const char* dimo_last_error(void) { return dimo::get_error(); }

int dimo_device_info(int* out3_host) {
  int dev = 0;
  DIMO_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  DIMO_CHECK_CUDA(cudaGetDeviceProperties(&p, dev));
  out3_host[0] = p.multiProcessorCount;
  out3_host[1] = (int)p.sharedMemPerBlockOptin;
  out3_host[2] = p.major * 10 + p.minor;
  return 0;
}

/* bytes of the counting sort's scratch (per-chunk tile histograms, tile totals and bases) */
size_t dimo_raster_bin_temp_bytes(int B, int N, int W, int H) {
  if (B <= 0 || N <= 0 || W <= 0 || H <= 0) return 256;
  const BinGeom g = bin_geom(B, N, W, H);
  return ((size_t)B * g.nchunk * g.T + 2 * (size_t)B * g.T + 2 + 2 * (size_t)B * N) * sizeof(uint32_t) + 256;
}

// Packed instances: when the tile key (bits_for(B*tiles + 1) bits) and the index of a Gaussian within its frame
// (bits_for(N) bits) fit one 32-bit word, instances are single words (key << vbits) | index: half the bytes.
int dimo_raster_packed_value_bits(int B, int N, int W, int H) {
  if (dimo::g_disable_packed_instances) return 0;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int64_t ntiles = (int64_t)B * gx * gy;
  if (B <= 0 || N <= 0 || ntiles <= 0) return 0;
  const int vbits = bits_for(N), kbits = bits_for(ntiles + 1);
  return vbits + kbits <= 32 ? vbits : 0;
}

int dimo_raster_preprocess(int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier,
                           int act_flags, const float* cams, const int32_t* frame_src, const float* means3D,
                           int64_t means3D_bstride, const float* scales,
                           int64_t scales_bstride, const float* rotations, int64_t rotations_bstride,
                           const float* opacities, int64_t opacities_bstride, const float* shs,
                           int64_t shs_bstride, const float* colors_precomp, int64_t colors_bstride,
                           float* splats, int32_t* radii, uint32_t* tiles_touched, uint32_t* rects,
                           uint32_t* sort_scratch, uint32_t* perm, uint64_t* total_count,
                           int64_t* R_host, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t BN = (int64_t)B * N;
  DIMO_REQUIRE(B >= 0 && N >= 0 && W > 0 && H > 0, "bad sizes");
  DIMO_REQUIRE(BN < ((int64_t)1 << 31), "B*N must fit int32");
  DIMO_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
  DIMO_REQUIRE((shs != nullptr) != (colors_precomp != nullptr), "exactly one of shs / colors_precomp");
  DIMO_REQUIRE(shs == nullptr || sh_coeffs >= (sh_degree + 1) * (sh_degree + 1), "sh_coeffs < (deg+1)^2");
  DIMO_REQUIRE((W + TILE - 1) / TILE <= 1023 && (H + TILE - 1) / TILE <= 1023, "image larger than 1023 tiles per side");
  if (BN == 0) {
    if (R_host) *R_host = 0;
    return 0;
  }
  DIMO_CHECK_CUDA(cudaMemsetAsync(total_count, 0, sizeof(uint64_t), st));
  // sort_scratch: [3*BN] u32 = depth keys | ping | pong; perm: [2*BN] u32 = ping | pong (result = perm + BN)
  int rc = preprocess_launch(B, N, W, H, sh_degree, sh_coeffs, scale_modifier, act_flags, cams, frame_src, means3D,
                             means3D_bstride, scales, scales_bstride, rotations, rotations_bstride, opacities,
                             opacities_bstride, shs, shs_bstride, colors_precomp, colors_bstride, splats, radii,
                             tiles_touched, rects, sort_scratch, reinterpret_cast<unsigned long long*>(total_count),
                             st);
  if (rc) return rc;
  if (B <= 15)
    depth_sort_kernel<1024><<<B * DS_CL, 1024, 0, st>>>(N, sort_scratch, sort_scratch + BN, sort_scratch + 2 * BN, perm,
                                                       perm + BN);
  else
    depth_sort_kernel<512><<<B * DS_CL, 512, 0, st>>>(N, sort_scratch, sort_scratch + BN, sort_scratch + 2 * BN, perm,
                                                     perm + BN);
  DIMO_CHECK_LAUNCH();
  if (R_host) {
    unsigned long long total = 0;
    DIMO_CHECK_CUDA(cudaMemcpyAsync(&total, total_count, sizeof(total), cudaMemcpyDeviceToHost, st));
    DIMO_CHECK_CUDA(cudaStreamSynchronize(st));
    *R_host = (int64_t)total;
  }
  return 0;
}

int dimo_raster_bin(int B, int N, int W, int H, int64_t R, const uint32_t* rects, const uint32_t* perm_sorted,
                    uint32_t* keys_sorted, uint32_t* vals_sorted, void* temp, size_t temp_bytes, uint32_t* ranges,
                    int32_t* count_overflow, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int64_t ntiles = (int64_t)B * gx * gy;
  const int64_t BN = (int64_t)B * N;
  DIMO_REQUIRE(R < ((int64_t)1 << 31), "instance count must fit int32");
  DIMO_REQUIRE(ntiles < ((int64_t)1 << 31) - 1, "B*tiles must fit int32");
  DIMO_REQUIRE(gx <= 1023 && gy <= 1023, "image larger than 1023 tiles per side");
  if (BN == 0 || ntiles == 0) {
    if (ntiles > 0) DIMO_CHECK_CUDA(cudaMemsetAsync(ranges, 0, sizeof(uint32_t) * 2 * ntiles, st));
    return 0;
  }
  const BinGeom g = bin_geom(B, N, W, H);
  DIMO_REQUIRE(g.wpc >= 1, "image too large for the tile counting sort (more than ~24k tiles per frame)");
  DIMO_REQUIRE(temp_bytes >= dimo_raster_bin_temp_bytes(B, N, W, H), "dimo_raster_bin: temp buffer too small");
  const int vbits = dimo_raster_packed_value_bits(B, N, W, H);
  DIMO_REQUIRE(vbits > 0 || keys_sorted != nullptr, "keys_sorted required for the (key, value) instance format");
  uint32_t* hist = reinterpret_cast<uint32_t*>(temp);
  uint32_t* tile_total = hist + (size_t)B * g.nchunk * g.T;
  uint32_t* tile_base = tile_total + (size_t)B * g.T;
  const size_t words = ((size_t)B * g.nchunk * g.T + 2 * (size_t)B * g.T + 1) & ~(size_t)1;   // keep uint2 alignment
  DIMO_REQUIRE((reinterpret_cast<uintptr_t>(temp) & 7) == 0, "dimo_raster_bin: temp must be 8-byte aligned");
  uint2* rects_sorted = reinterpret_cast<uint2*>(hist + words);
  const size_t cells = (size_t)(g.gx + 1) * (g.gy + 1) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(tile_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(tile_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DIMO_CHECK_CUDA(cudaFuncSetAttribute(tile_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const uint2* r2 = reinterpret_cast<const uint2*>(rects);
  dim3 grid(g.nchunk, B);
  tile_hist_kernel<<<grid, 256, cells, st>>>(g, r2, perm_sorted, hist, rects_sorted);
  DIMO_CHECK_LAUNCH();
  tile_chunk_scan_kernel<<<ceil_div(ntiles, 256), 256, 0, st>>>(g, hist, tile_total);
  DIMO_CHECK_LAUNCH();
  tile_base_scan_kernel<<<1, TS_THREADS, 0, st>>>(ntiles, R, tile_total, tile_base, reinterpret_cast<uint2*>(ranges),
                                                  count_overflow);
  DIMO_CHECK_LAUNCH();
  if (R == 0) return 0;
  if (vbits > 0)
    tile_scatter_kernel<true><<<grid, g.wpc * 32, g.wpc * cells, st>>>(g, R, vbits, rects_sorted, perm_sorted, hist,
                                                                      tile_base, bits_for(g.T), keys_sorted, vals_sorted);
  else
    tile_scatter_kernel<false><<<grid, g.wpc * 32, g.wpc * cells, st>>>(g, R, 0, rects_sorted, perm_sorted, hist,
                                                                       tile_base, bits_for(g.T), keys_sorted, vals_sorted);
  DIMO_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
