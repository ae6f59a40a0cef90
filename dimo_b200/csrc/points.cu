// Small point-set kernels on the control points / Gaussian centres that the step's regularisers and the key-point
// annealing use:
//   dimo_fps          replaces pytorch3d.ops.sample_farthest_points(points[1,N,3], K)      (GUI.FPS, main_train_dimo.py:511-515)
//   dimo_ball_query   replaces pytorch3d.ops.ball_query(p, p, K=11, radius=0.1)            (utils/deform_utils.py:128, ARAP connectivity)
//   dimo_chamfer_fwd / _bwd  replace chamferdist.ChamferDistance()(cpts, cpts_ori)         (main_train_dimo.py:298-299)
//
// COMPILED WITH -fmad=false: d2 = (dx*dx + dy*dy) + dz*dz as separately rounded fp32 operations, the sequence
// oracle/points.py uses, so the selected indices / neighbour lists are bit-exact (ties -> lower index).
//
// pytorch3d and chamferdist are pip dependencies that are absent from /root/reference ("parity unpinned"): semantics
// restated from their published behaviour -- FPS starts at index 0 and keeps, per point, the squared distance to the
// nearest selected point; ball_query lists the first K points in INDEX order with squared distance < radius^2, padding
// idx with -1 and dists with 0; ChamferDistance's default is the one-directional sum over source points of the
// squared distance to the nearest target point.
#include "common.cuh"

namespace dimo {

// ------------------------------------------------------------------------------------------------------------------
// Farthest point sampling: one CTA per cloud (the K selections are sequential; each is a min-update + arg-max over N).
// N = 512..1e5 here, run once per 1000 steps; `mind` ([B,N] fp32 scratch) stays L1/L2 resident.
// ------------------------------------------------------------------------------------------------------------------
constexpr int FPS_THREADS = 1024;

__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(int N, int K, int start, const float* __restrict__ pts_all,
                                                          float* __restrict__ mind_all, int64_t* __restrict__ out_all) {
  const float* pts = pts_all + (int64_t)blockIdx.x * N * 3;
  float* mind = mind_all + (int64_t)blockIdx.x * N;
  int64_t* out = out_all + (int64_t)blockIdx.x * K;
  __shared__ float s_d[32];
  __shared__ int s_i[32];
  __shared__ int s_sel;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < N; i += FPS_THREADS) mind[i] = INFINITY;
  int sel = start;
  if (tid == 0) out[0] = sel;
  __syncthreads();
  for (int k = 1; k < K; ++k) {
    const float sx = pts[3 * (int64_t)sel], sy = pts[3 * (int64_t)sel + 1], sz = pts[3 * (int64_t)sel + 2];
    float best = -1.0f;
    int bi = 0x7fffffff;
    for (int i = tid; i < N; i += FPS_THREADS) {
      const float dx = pts[3 * (int64_t)i] - sx, dy = pts[3 * (int64_t)i + 1] - sy, dz = pts[3 * (int64_t)i + 2] - sz;
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      const float m = fminf(mind[i], d2);
      mind[i] = m;
      if (m > best) { best = m; bi = i; }          // i ascends within a thread: the first maximum is kept
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (od > best || (od == best && oi < bi)) { best = od; bi = oi; }
    }
    if (lane == 0) { s_d[warp] = best; s_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      best = s_d[lane]; bi = s_i[lane];             // FPS_THREADS / 32 == 32 warps
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od > best || (od == best && oi < bi)) { best = od; bi = oi; }
      }
      if (lane == 0) { s_sel = bi; out[k] = bi; }
    }
    __syncthreads();
    sel = s_sel;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Ball query: thread per query point, candidates staged through shared memory, walked in index order.
// ------------------------------------------------------------------------------------------------------------------
constexpr int BQ_TILE = 2048;

__global__ void __launch_bounds__(128) ball_query_kernel(int P1, int P2, int K, float radius2,
                                                         const float* __restrict__ p1_all,
                                                         const float* __restrict__ p2_all, int64_t* __restrict__ idx_all,
                                                         float* __restrict__ dist_all) {
  __shared__ float sp[BQ_TILE * 3];
  const int b = blockIdx.y;
  const float* p1 = p1_all + (int64_t)b * P1 * 3;
  const float* p2 = p2_all + (int64_t)b * P2 * 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int64_t* idx = idx_all + ((int64_t)b * P1 + i) * K;
  float* dist = dist_all + ((int64_t)b * P1 + i) * K;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (i < P1) { qx = p1[3 * (int64_t)i]; qy = p1[3 * (int64_t)i + 1]; qz = p1[3 * (int64_t)i + 2]; }
  int count = 0;
  for (int base = 0; base < P2; base += BQ_TILE) {
    const int cnt = min(BQ_TILE, P2 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += blockDim.x) sp[e] = p2[3 * (int64_t)base + e];
    __syncthreads();
    if (i < P1) {
      for (int j = 0; j < cnt && count < K; ++j) {
        const float dx = sp[3 * j] - qx, dy = sp[3 * j + 1] - qy, dz = sp[3 * j + 2] - qz;
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 < radius2) { idx[count] = base + j; dist[count] = d2; ++count; }
      }
    }
  }
  if (i < P1)
    for (int c = count; c < K; ++c) { idx[c] = -1; dist[c] = 0.0f; }
}

// ------------------------------------------------------------------------------------------------------------------
// One-directional chamfer term: for every source point the nearest target point (squared distance, index), the sum
// over the source points accumulated (times `lw`) into a device loss scalar.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) chamfer_fwd_kernel(int N, int M, const float* __restrict__ src,
                                                          const float* __restrict__ tgt, float* __restrict__ d2_out,
                                                          int32_t* __restrict__ nn_out, float* __restrict__ sum_out,
                                                          float* __restrict__ loss_acc, float lw) {
  __shared__ float sp[BQ_TILE * 3];
  __shared__ float s_part[4];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (i < N) { qx = src[3 * (int64_t)i]; qy = src[3 * (int64_t)i + 1]; qz = src[3 * (int64_t)i + 2]; }
  float best = INFINITY;
  int bi = -1;
  for (int base = 0; base < M; base += BQ_TILE) {
    const int cnt = min(BQ_TILE, M - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += blockDim.x) sp[e] = tgt[3 * (int64_t)base + e];
    __syncthreads();
    if (i < N) {
      for (int j = 0; j < cnt; ++j) {
        const float dx = qx - sp[3 * j], dy = qy - sp[3 * j + 1], dz = qz - sp[3 * j + 2];
        const float d2 = (dx * dx + dy * dy) + dz * dz;
        if (d2 < best) { best = d2; bi = base + j; }
      }
    }
  }
  float v = 0.f;
  if (i < N) { d2_out[i] = best; nn_out[i] = bi; v = best; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float t = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
    if (sum_out) atomicAdd(sum_out, t);
    if (loss_acc) atomicAdd(loss_acc, lw * t);
  }
}

// d src_i = 2 g (src_i - tgt_nn(i));  d tgt_j = -sum_{i: nn(i)=j} 2 g (src_i - tgt_j)   (optional)
__global__ void __launch_bounds__(128) chamfer_bwd_kernel(int N, const float* __restrict__ src,
                                                          const float* __restrict__ tgt, const int32_t* __restrict__ nn,
                                                          const float* __restrict__ g_scalar, float gw,
                                                          float* __restrict__ d_src, float* __restrict__ d_tgt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float g = 2.0f * gw * (g_scalar ? g_scalar[0] : 1.0f);
  const int j = nn[i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float d = g * (src[3 * (int64_t)i + c] - tgt[3 * (int64_t)j + c]);
    if (d_src) d_src[3 * (int64_t)i + c] = d;
    if (d_tgt) atomicAdd(d_tgt + 3 * (int64_t)j + c, -d);
  }
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_fps(int B, int N, int K, int start, const float* points, float* min_dist_scratch, int64_t* idx,
                        void* stream) {
  DIMO_REQUIRE(B >= 1 && N >= 1, "need at least one point");
  DIMO_REQUIRE(K >= 1 && K <= N, "K must be 1..N");
  DIMO_REQUIRE(start >= 0 && start < N, "start index out of range");
  fps_kernel<<<B, FPS_THREADS, 0, (cudaStream_t)stream>>>(N, K, start, points, min_dist_scratch, idx);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_ball_query(int B, int P1, int P2, int K, float radius, const float* p1, const float* p2,
                               int64_t* idx, float* dists, void* stream) {
  DIMO_REQUIRE(K >= 1, "K must be positive");
  DIMO_REQUIRE(B >= 1 && B <= 65535, "batch must be 1..65535");
  if (P1 == 0) return 0;
  dim3 grid(ceil_div(P1, 128), B);
  ball_query_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P1, P2, K, radius * radius, p1, p2, idx, dists);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_chamfer_fwd(int N, int M, const float* src, const float* tgt, float* d2, int32_t* nn, float* sum,
                                float* loss_acc, float lw, void* stream) {
  DIMO_REQUIRE(M >= 1, "need at least one target point");
  if (N == 0) return 0;
  chamfer_fwd_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(N, M, src, tgt, d2, nn, sum, loss_acc, lw);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_chamfer_bwd(int N, const float* src, const float* tgt, const int32_t* nn, const float* g_scalar,
                                float gw, float* d_src, float* d_tgt, void* stream) {
  if (N == 0) return 0;
  chamfer_bwd_kernel<<<ceil_div(N, 128), 128, 0, (cudaStream_t)stream>>>(N, src, tgt, nn, g_scalar, gw, d_src, d_tgt);
  DIMO_CHECK_LAUNCH();
  return 0;
}
