// Nearest-neighbour terms.
//   dimo_knn      replaces knn_cuda.KNN(k, transpose_mode=True)(ref[1,M,3], query[1,N,3])
//                 (main_train_dimo.py:502-509): Gaussian -> k nearest control points.
//   dimo_dist3nn  replaces simple_knn._C.distCUDA2 (renderer/latent_gs_renderer.py:426).
//
// COMPILED WITH -fmad=false: d2 = (dx*dx + dy*dy) + dz*dz as separately rounded fp32 operations, the
// same sequence oracle/knn.py uses, so neighbour indices are bit-exact (ties -> lower index).
//
// B200 design: the reference points are staged through shared memory in tiles; every thread owns one
// query and keeps its k best in registers (no M x N distance matrix is ever materialised -- upstream
// KNN_CUDA builds one: 512 x 500k x 4 B = 1 GB at config c5).  Output traffic 48 B/query (k=4).
#include "common.cuh"

namespace dimo {

constexpr int KNN_TILE = 2048;   // reference points per smem tile (24 KB)
constexpr int KNN_MAXK = 8;

template <int K>
__global__ void __launch_bounds__(256) knn_kernel(int M, int N, const float* __restrict__ ref,
                                                  const float* __restrict__ query, float* __restrict__ dist,
                                                  int64_t* __restrict__ idx) {
  __shared__ float sref[KNN_TILE * 3];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (i < N) { qx = query[3 * (int64_t)i]; qy = query[3 * (int64_t)i + 1]; qz = query[3 * (int64_t)i + 2]; }
  float bd[K];
  int bi[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { bd[k] = INFINITY; bi[k] = -1; }

  for (int base = 0; base < M; base += KNN_TILE) {
    const int cnt = min(KNN_TILE, M - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += blockDim.x) sref[e] = ref[3 * (int64_t)base + e];
    __syncthreads();
    if (i < N) {
      for (int j = 0; j < cnt; ++j) {
        const float dx = qx - sref[3 * j], dy = qy - sref[3 * j + 1], dz = qz - sref[3 * j + 2];
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 < bd[K - 1]) {      // strict: equal distances keep the earlier (lower) index
          bd[K - 1] = d2; bi[K - 1] = base + j;
#pragma unroll
          for (int k = K - 1; k > 0; --k) {
            if (bd[k] < bd[k - 1]) {
              const float td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
              const int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
            }
          }
        }
      }
    }
  }
  if (i < N) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      dist[(int64_t)i * K + k] = sqrtf(bd[k]);
      idx[(int64_t)i * K + k] = (int64_t)bi[k];
    }
  }
}

// mean squared distance to the 3 nearest other points; tiled brute force (exact).
__global__ void __launch_bounds__(256) dist3nn_kernel(int N, const float* __restrict__ pts, float* __restrict__ out) {
  __shared__ float sp[KNN_TILE * 3];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (i < N) { qx = pts[3 * (int64_t)i]; qy = pts[3 * (int64_t)i + 1]; qz = pts[3 * (int64_t)i + 2]; }
  float b0 = INFINITY, b1 = INFINITY, b2 = INFINITY;
  for (int base = 0; base < N; base += KNN_TILE) {
    const int cnt = min(KNN_TILE, N - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += blockDim.x) sp[e] = pts[3 * (int64_t)base + e];
    __syncthreads();
    if (i < N) {
      for (int j = 0; j < cnt; ++j) {
        const float dx = qx - sp[3 * j], dy = qy - sp[3 * j + 1], dz = qz - sp[3 * j + 2];
        float d2 = (dx * dx + dy * dy) + dz * dz;
        if (base + j == i) d2 = INFINITY;
        if (d2 < b2) {
          b2 = d2;
          if (b2 < b1) { const float t = b1; b1 = b2; b2 = t; }
          if (b1 < b0) { const float t = b0; b0 = b1; b1 = t; }
        }
      }
    }
  }
  if (i < N) out[i] = ((b0 + b1) + b2) / 3.0f;
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_knn(int M, int N, int k, const float* ref, const float* query, float* dist, int64_t* idx,
                        void* stream) {
  DIMO_REQUIRE(k >= 1 && k <= KNN_MAXK, "k must be 1..8");
  DIMO_REQUIRE(M >= k, "need at least k reference points");
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ceil_div(N, 256);
  switch (k) {
#define DIMO_KNN_CASE(KK) \
  case KK: knn_kernel<KK><<<grid, 256, 0, st>>>(M, N, ref, query, dist, idx); break;
    DIMO_KNN_CASE(1) DIMO_KNN_CASE(2) DIMO_KNN_CASE(3) DIMO_KNN_CASE(4)
    DIMO_KNN_CASE(5) DIMO_KNN_CASE(6) DIMO_KNN_CASE(7) DIMO_KNN_CASE(8)
#undef DIMO_KNN_CASE
  }
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_dist3nn(int N, const float* points, float* out, void* stream) {
  if (N == 0) return 0;
  dist3nn_kernel<<<ceil_div(N, 256), 256, 0, (cudaStream_t)stream>>>(N, points, out);
  DIMO_CHECK_LAUNCH();
  return 0;
}
