// ARAP term of the step (Renderer.arap_loss_v2, renderer/latent_gs_renderer.py:1081-1094 ->
// utils/deform_utils.py:115-150 connectivity, :152-232 rotation fit + energy) in TWO launches:
//
//   dimo_arap_connectivity   thread per vertex: ball query (first K+1 hits in index order, first one dropped) in every
//                            frame, lists intersected over the T frames -> neighbour table nbr [M,K] (-1 padded)
//   dimo_arap_energy         thread per (target frame, vertex): edge fans, 3x3 covariance, Kabsch rotation (fp64 Jacobi),
//                            energy and its gradient w.r.t. every node position (R constant, as in the reference)
//
// The reference spends ~40 torch launches per call on M = 512 points (one_hot over [T,M,K,M+1], top-k, a Python loop
// over frames with a batched LAPACK-style SVD each) and calls it once per motion of the batch; here the arithmetic is
// csrc/arap_math.h, which the CPU tests run through a host build against the reference-pinned fixture.
// Compiled with -fmad=false: the ball-query distances are the same separately rounded fp32 sequence as csrc/points.cu.
#include "arap_math.h"
#include "common.cuh"

namespace dimo {

__global__ void __launch_bounds__(128) arap_connectivity_kernel(int T, int M, int Kq, int K, float r2,
                                                                const float* __restrict__ nodes,
                                                                int64_t* __restrict__ nbr, int32_t* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int c = arap::common_neighbours(T, M, Kq, K, r2, nodes, i, nbr + (int64_t)i * K);
  if (count) count[i] = c;
}

struct AtomicAdd {
  __device__ __forceinline__ void operator()(float* p, float v) const { atomicAdd(p, v); }
};

__global__ void __launch_bounds__(128) arap_energy_kernel(int T, int M, int K, const float* __restrict__ nodes,
                                                          const int64_t* __restrict__ nbr, const float* __restrict__ mult,
                                                          float* __restrict__ energy, float* __restrict__ grad) {
  __shared__ float s_part[4];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  float v = 0.0f;
  if (e < (T - 1) * M) {
    const int t = 1 + e / M, i = e % M;
    const float m = mult ? mult[i] : 1.0f;
    if (m != 0.0f)
      v = arap::vertex_term(K, nodes, nodes + (int64_t)t * M * 3, nbr + (int64_t)i * K, i, m, grad,
                            grad ? grad + (int64_t)t * M * 3 : nullptr, AtomicAdd());
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(energy, (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]));
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_arap_connectivity(int T, int M, int K, float radius, const float* nodes, int64_t* nbr,
                                      int32_t* count, void* stream) {
  DIMO_REQUIRE(T >= 1 && M >= 1, "need at least one frame and one vertex");
  DIMO_REQUIRE(K >= 1 && K < arap::MAXK, "K must be 1..15");
  arap_connectivity_kernel<<<ceil_div(M, 128), 128, 0, (cudaStream_t)stream>>>(T, M, K + 1, K, radius * radius, nodes,
                                                                                nbr, count);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_arap_energy(int T, int M, int K, const float* nodes, const int64_t* nbr, const float* mult,
                                float* energy, float* grad, void* stream) {
  DIMO_REQUIRE(T >= 1 && M >= 1, "need at least one frame and one vertex");
  DIMO_REQUIRE(K >= 1 && K <= arap::MAXK, "K must be 1..16");
  cudaStream_t st = (cudaStream_t)stream;
  DIMO_CHECK_CUDA(cudaMemsetAsync(energy, 0, sizeof(float), st));
  if (grad) DIMO_CHECK_CUDA(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)T * M * 3, st));
  if (T == 1) return 0;
  arap_energy_kernel<<<ceil_div((int64_t)(T - 1) * M, 128), 128, 0, st>>>(T, M, K, nodes, nbr, mult, energy, grad);
  DIMO_CHECK_LAUNCH();
  return 0;
}
