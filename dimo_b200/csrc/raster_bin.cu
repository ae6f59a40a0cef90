// Binning stage of the tile rasteriser: inclusive scan of per-splat tile counts, key emission,
// stable radix sort by tile id, per-tile ranges.
//
// Upstream shape (diff_gauss / diff_gaussian_rasterization, call site
// renderer/latent_gs_renderer.py:1256-1277): InclusiveSum -> duplicateWithKeys -> SortPairs over
// 64-bit (tile | depth) keys (6 radix passes at 512^2 x 16 frames = 144 B per instance) -> identifyTileRanges.
//
// Here the same total order is produced with ~4.5x less traffic:
//   1. the B*N splats are sorted ONCE by (frame, depth) -- 64-bit keys, but only B*N of them;
//   2. tile counts are scanned in that order and instances are emitted front-to-back;
//   3. the R instances are STABLE-sorted by the tile id alone: 32-bit keys, ceil(log2(B*tiles)) bits
//      (14 bits = 2 passes at 512^2 x 16 frames, 32 B per instance).
// A stable sort keeps the emission order inside a tile, i.e. ascending depth with ties in ascending Gaussian
// index -- bit-for-bit the order of the 64-bit sort (tests compare against the oracle's stable 64-bit sort).
// The blend kernels gather the 64-byte blend records of a tile's list by index (`vals_sorted` -> record table
// written by the preprocess kernel): the table (B*N records) is far smaller than the instance list and mostly
// L2-resident, and only the part of a list in front of the saturation depth is ever fetched -- materialising all R
// records in sorted order (the first version of this file) cost more than both sorts together.  Scan and sorts
// are CUB device primitives compiled into this library (integer-only, HBM-bound).
#include "common.cuh"
#include <cub/cub.cuh>
#include <stdarg.h>

namespace dimo {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// Per-tile [begin, end) ranges of the sorted instance list: every slot's key is compared with its neighbours.
//
// `R` is the number of SLOTS: with an exact instance count every slot is valid; in capacity mode (no host
// read-back of the count) the unused tail carries sentinel keys (>= ntiles) which sort last and are skipped here.
// Four slots per thread (one 128-bit load + the two neighbouring keys): 16 B in flight per thread instead of 4.
__global__ void __launch_bounds__(256) tile_ranges_kernel(int64_t R, uint32_t ntiles, const uint32_t* __restrict__ keys,
                                                          int shift, uint2* __restrict__ ranges) {
  const int64_t j0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j0 >= R) return;
  constexpr uint32_t NONE = 0xFFFFFFFFu;      // never a valid tile id (ntiles < 2^31)
  uint32_t k[6];
  // `shift` > 0: packed instances, the tile key is the word's upper part
  k[0] = j0 > 0 ? keys[j0 - 1] >> shift : NONE;
  if (j0 + 4 <= R) {
    const uint4 v = *reinterpret_cast<const uint4*>(keys + j0);
    k[1] = v.x >> shift; k[2] = v.y >> shift; k[3] = v.z >> shift; k[4] = v.w >> shift;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) k[1 + i] = j0 + i < R ? keys[j0 + i] >> shift : NONE;
  }
  k[5] = j0 + 4 < R ? keys[j0 + 4] >> shift : NONE;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t tile = k[1 + i];
    if (tile >= ntiles) continue;             // sentinel slot (capacity mode) or past the end
    if (k[i] != tile) ranges[tile].x = (uint32_t)(j0 + i);
    if (k[2 + i] != tile) ranges[tile].y = (uint32_t)(j0 + i + 1);
  }
}

// capacity mode: flags an instance count larger than the number of slots (the surplus instances were dropped)
__global__ void overflow_check_kernel(const uint32_t* __restrict__ offsets, int64_t BN, int64_t slots,
                                      int32_t* __restrict__ flag) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const uint32_t total = offsets[BN - 1];
    flag[0] = (int32_t)total;
    if ((int64_t)total > slots) flag[1] = 1;
  }
}

int g_disable_packed_instances = 0;     // dimo_tc_debug_set key 6: force the (key, value) pair format (tests, A/B)

static inline int bits_for(int64_t n) {
  int bits = 1;
  while (((int64_t)1 << bits) < n) ++bits;
  return bits;
}

// gathers tile counts in depth-sorted order for the scan
struct PermutedCount {
  const uint32_t* counts;
  const uint32_t* perm;
  __host__ __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return counts[perm[i]]; }
};

}  // namespace dimo

using namespace dimo;

namespace dimo {
int preprocess_launch(int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier,
                      const float* cams, const int32_t* frame_src, const float* means3D, int64_t means3D_bstride,
                      const float* scales,
                      int64_t scales_bstride, const float* rotations, int64_t rotations_bstride,
                      const float* opacities, int64_t opacities_bstride, const float* shs, int64_t shs_bstride,
                      const float* colors_precomp, int64_t colors_bstride, float* splats, int32_t* radii,
                      uint32_t* tiles_touched, uint64_t* depth_keys, uint32_t* iota, cudaStream_t st);
int emit_keys_launch(int B, int N, int W, int H, int64_t R, const float* splats, const int32_t* radii,
                     const uint32_t* perm, const uint32_t* offsets, uint32_t* tile_keys, uint32_t* vals, int vbits,
                     cudaStream_t st);
}  // namespace dimo

extern "C" {

int dimo_abi_version(void) { return DIMO_ABI_VERSION; }
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

size_t dimo_raster_scan_temp_bytes(int64_t BN) {
  size_t scan = 0, sort = 0;
  cub::DeviceScan::InclusiveSum(nullptr, scan, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)BN);
  cub::DeviceRadixSort::SortPairs(nullptr, sort, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)BN, 0, 64);
  return (scan > sort ? scan : sort) + 256;
}

// Packed instances: when the tile key (bits_for(B*tiles + 1) bits, one spare code for the capacity-mode sentinel)
// and the index of a Gaussian within its frame (bits_for(N) bits) fit one 32-bit word, instances are single words
// (key << vbits) | index and the tile sort is a keys-only radix sort over the key bits: half the bytes per pass.
int dimo_raster_packed_value_bits(int B, int N, int W, int H) {
  if (dimo::g_disable_packed_instances) return 0;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int64_t ntiles = (int64_t)B * gx * gy;
  if (B <= 0 || N <= 0 || ntiles <= 0) return 0;
  const int vbits = bits_for(N), kbits = bits_for(ntiles + 1);
  return vbits + kbits <= 32 ? vbits : 0;
}

size_t dimo_raster_sort_temp_bytes(int64_t R) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)R, 0, 32);
  return bytes + 256;
}

int dimo_raster_preprocess(int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier,
                           const float* cams, const int32_t* frame_src, const float* means3D,
                           int64_t means3D_bstride, const float* scales,
                           int64_t scales_bstride, const float* rotations, int64_t rotations_bstride,
                           const float* opacities, int64_t opacities_bstride, const float* shs,
                           int64_t shs_bstride, const float* colors_precomp, int64_t colors_bstride,
                           float* splats, int32_t* radii, uint32_t* tiles_touched, uint32_t* offsets,
                           uint64_t* depth_keys, uint32_t* perm, void* scan_temp, size_t scan_temp_bytes,
                           int64_t* R_host, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t BN = (int64_t)B * N;
  DIMO_REQUIRE(B >= 0 && N >= 0 && W > 0 && H > 0, "bad sizes");
  DIMO_REQUIRE(BN < ((int64_t)1 << 31), "B*N must fit int32");
  DIMO_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "sh_degree must be 0..3");
  DIMO_REQUIRE((shs != nullptr) != (colors_precomp != nullptr), "exactly one of shs / colors_precomp");
  DIMO_REQUIRE(shs == nullptr || sh_coeffs >= (sh_degree + 1) * (sh_degree + 1), "sh_coeffs < (deg+1)^2");
  if (BN == 0) {
    if (R_host) *R_host = 0;
    return 0;
  }
  // depth_keys: [2*BN] (unsorted | sorted), perm: [2*BN] (iota | sorted permutation = perm + BN)
  int rc = preprocess_launch(B, N, W, H, sh_degree, sh_coeffs, scale_modifier, cams, frame_src, means3D,
                             means3D_bstride, scales,
                             scales_bstride, rotations, rotations_bstride, opacities, opacities_bstride, shs,
                             shs_bstride, colors_precomp, colors_bstride, splats, radii, tiles_touched, depth_keys,
                             perm, st);
  if (rc) return rc;
  size_t need = scan_temp_bytes;
  DIMO_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(scan_temp, need, depth_keys, depth_keys + BN, perm, perm + BN,
                                                  (int)BN, 0, 32 + bits_for(B), st));
  need = scan_temp_bytes;
  cub::TransformInputIterator<uint32_t, PermutedCount, cub::CountingInputIterator<uint32_t>> counts_sorted(
      cub::CountingInputIterator<uint32_t>(0), PermutedCount{tiles_touched, perm + BN});
  DIMO_CHECK_CUDA(cub::DeviceScan::InclusiveSum(scan_temp, need, counts_sorted, offsets, (int)BN, st));
  if (R_host) {
    uint32_t last = 0;
    DIMO_CHECK_CUDA(cudaMemcpyAsync(&last, offsets + (BN - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    DIMO_CHECK_CUDA(cudaStreamSynchronize(st));
    *R_host = (int64_t)last;
  }
  return 0;
}

int dimo_raster_bin(int B, int N, int W, int H, int64_t R, const float* splats, const int32_t* radii,
                    const uint32_t* perm_sorted, const uint32_t* offsets, uint32_t* keys_unsorted,
                    uint32_t* vals_unsorted, uint32_t* keys_sorted, uint32_t* vals_sorted, void* sort_temp,
                    size_t sort_temp_bytes, uint32_t* ranges, int32_t* count_overflow, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
  const int64_t ntiles = (int64_t)B * gx * gy;
  const int64_t BN = (int64_t)B * N;
  DIMO_REQUIRE(R < ((int64_t)1 << 31), "instance count must fit int32");
  DIMO_REQUIRE(ntiles < ((int64_t)1 << 31) - 1, "B*tiles must fit int32");
  const int vbits = dimo_raster_packed_value_bits(B, N, W, H);
  uint32_t* const first_unsorted = vbits > 0 ? vals_unsorted : keys_unsorted;   // buffer that carries the key bits
  DIMO_REQUIRE(((uintptr_t)(vbits > 0 ? vals_sorted : keys_sorted) & 15) == 0, "sorted instance buffer must be 16-byte aligned");
  DIMO_CHECK_CUDA(cudaMemsetAsync(ranges, 0, sizeof(uint32_t) * 2 * ntiles, st));
  if (R == 0 || BN == 0) return 0;
  if (count_overflow != nullptr) {
    // capacity mode: R is a slot count chosen by the caller; unused slots keep the all-ones sentinel key
    DIMO_CHECK_CUDA(cudaMemsetAsync(first_unsorted, 0xFF, sizeof(uint32_t) * (size_t)R, st));
    overflow_check_kernel<<<1, 32, 0, st>>>(offsets, BN, R, count_overflow);
    DIMO_CHECK_LAUNCH();
  }
  int rc = emit_keys_launch(B, N, W, H, R, splats, radii, perm_sorted, offsets, keys_unsorted, vals_unsorted, vbits, st);
  if (rc) return rc;
  size_t need = sort_temp_bytes;
  // one spare code above the last tile id so that the sentinel sorts behind every real key
  const int kbits = bits_for(ntiles + 1);
  if (vbits > 0) {
    DIMO_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(sort_temp, need, vals_unsorted, vals_sorted, (int)R, vbits,
                                                   vbits + kbits, st));
    tile_ranges_kernel<<<ceil_div(R, 1024), 256, 0, st>>>(R, (uint32_t)ntiles, vals_sorted, vbits,
                                                         reinterpret_cast<uint2*>(ranges));
  } else {
    DIMO_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(sort_temp, need, keys_unsorted, keys_sorted, vals_unsorted,
                                                    vals_sorted, (int)R, 0, kbits, st));
    tile_ranges_kernel<<<ceil_div(R, 1024), 256, 0, st>>>(R, (uint32_t)ntiles, keys_sorted, 0,
                                                         reinterpret_cast<uint2*>(ranges));
  }
  DIMO_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
