// Optimizer step of the loop: torch.optim.Adam(l, lr=0.0, eps=1e-15) over the twelve parameter
// groups of GaussianModel.training_setup (renderer/latent_gs_renderer.py:453-476), stepped at
// main_train_dimo.py:416-417 (optimizer.step(); optimizer.zero_grad()).
//
// B200 design: every parameter is a view into ONE flat fp32 buffer and so is its gradient (the buffer the
// all-reduce runs on, dimo_b200/dist.py), so the whole optimizer is ONE launch: a grid-stride pass of 128-bit
// loads/stores over (param, grad, exp_avg, exp_avg_sq) that also clears the gradient for the next step
// (zero_grad folded in: 28 B read+written per element -> 32 B with the clear, instead of two more passes).
// Per-group learning rates and the step counter live in device memory, so a captured CUDA graph replays the same
// launch while the host changes learning rates (update_learning_rate, :502-520) between replays.
// Bound: HBM (the buffers total 4 x 8.2 MB at the c3 shape and mostly sit in the 126 MB L2).
//
// Also here: the grouped transpose that produces W^T for the tensor-core data-gradient GEMMs in one launch.
#include "common.cuh"

namespace dimo {

constexpr int ADAM_THREADS = 256;
constexpr int ADAM_MAX_SEGS = 64;

struct AdamSegs {
  int n;
  int64_t begin[ADAM_MAX_SEGS + 1];   // element offsets, multiples of 4; begin[n] = total
};

// state (device, 4 x i32): [0] step count t (number of updates applied so far), [1] ticket counter
__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(int64_t n4, float4* __restrict__ p, float4* __restrict__ g,
                                                            float4* __restrict__ m, float4* __restrict__ v,
                                                            const float* __restrict__ seg_lr, AdamSegs segs,
                                                            double beta1d, double beta2d, float eps,
                                                            int zero_grads, int* __restrict__ state,
                                                            const float* __restrict__ skip_flag) {
  __shared__ int64_t s_begin[ADAM_MAX_SEGS + 1];
  __shared__ float s_lr[ADAM_MAX_SEGS];
  __shared__ float s_bc[2];
  for (int i = threadIdx.x; i <= segs.n; i += ADAM_THREADS) s_begin[i] = segs.begin[i];
  for (int i = threadIdx.x; i < segs.n; i += ADAM_THREADS) s_lr[i] = seg_lr[i];
  const int t = state[0] + 1;
  // skip_flag (device float, may be NULL; lives OUTSIDE the buffers this kernel writes): non-zero = the step that
  // produced these gradients dropped work (the rasteriser's instance capacity overflowed on some rank) -> no update,
  // the step counter stays, the gradients are still cleared, state[3] counts the skipped steps
  const bool skip = skip_flag != nullptr && skip_flag[0] != 0.f;
  if (threadIdx.x == 0) {
    // bias corrections in double, as torch.optim.Adam forms them on the host (1 - beta ** step): float(0.999) alone
    // is already 1.3e-5 (relative) away from 1 - 0.999 at step 1.  One thread per CTA, two pow() calls.
    s_bc[0] = (float)(1.0 / (1.0 - pow(beta1d, (double)t)));
    s_bc[1] = (float)sqrt(1.0 - pow(beta2d, (double)t));
  }
  __syncthreads();
  const float beta2 = (float)beta2d;
  const float omb1 = (float)(1.0 - beta1d), omb2 = (float)(1.0 - beta2d);
  const float inv_bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  int seg = 0;
  for (int64_t i = (int64_t)blockIdx.x * ADAM_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * ADAM_THREADS) {
    const int64_t e = i * 4;
    // segments are few and a thread walks the buffer monotonically: advance linearly
    while (seg + 1 < segs.n && e >= s_begin[seg + 1]) ++seg;
    if (skip) {
      if (zero_grads) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float step_size = s_lr[seg] * inv_bc1;
    float4 P = p[i], G = g[i], M = m[i], V = v[i];
    float* pp = &P.x; float* gp = &G.x; float* mp = &M.x; float* vp = &V.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gp[k];
      mp[k] = mp[k] + omb1 * (gk - mp[k]);                      // exp_avg.lerp_(grad, 1 - beta1)
      vp[k] = beta2 * vp[k] + omb2 * gk * gk;                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      pp[k] = pp[k] - step_size * (mp[k] / denom);               // param.addcdiv_(exp_avg, denom, value=-step_size)
    }
    p[i] = P; m[i] = M; v[i] = V;
    if (zero_grads) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // the last CTA to finish publishes the new step count (every CTA read state[0] before taking its ticket)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int ticket = atomicAdd(&state[1], 1);
    if (ticket == (int)gridDim.x - 1) {
      if (skip) state[3] += 1; else state[0] = t;
      state[1] = 0;
      __threadfence();
    }
  }
}

constexpr int TR_TILE = 32, TR_MAX = 16;
struct TransposeArgs {
  int n;
  int rows[TR_MAX], cols[TR_MAX], tile0[TR_MAX + 1];   // tile0: first CTA of each matrix
  const float* src[TR_MAX];
  float* dst[TR_MAX];
};

// dst[c][r] = src[r][c] for n row-major matrices; one 32x32 tile per CTA (32x8 threads), padded shared tile
__global__ void __launch_bounds__(256) transpose_grouped_kernel(TransposeArgs a) {
  __shared__ float tile[TR_TILE][TR_TILE + 1];
  int k = 0;
  while (k + 1 < a.n && (int)blockIdx.x >= a.tile0[k + 1]) ++k;
  const int rows = a.rows[k], cols = a.cols[k];
  const int tcols = (cols + TR_TILE - 1) / TR_TILE;
  const int tidx = blockIdx.x - a.tile0[k];
  const int r0 = (tidx / tcols) * TR_TILE, c0 = (tidx % tcols) * TR_TILE;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* __restrict__ src = a.src[k];
  float* __restrict__ dst = a.dst[k];
#pragma unroll
  for (int j = ty; j < TR_TILE; j += 8) {
    const int r = r0 + j, c = c0 + tx;
    if (r < rows && c < cols) tile[j][tx] = src[(int64_t)r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = ty; j < TR_TILE; j += 8) {
    const int c = c0 + j, r = r0 + tx;
    if (r < rows && c < cols) dst[(int64_t)c * rows + r] = tile[tx][j];
  }
}

}  // namespace dimo

using namespace dimo;

extern "C" int dimo_adam_step(int64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int nseg,
                              const int64_t* seg_begin_host, const float* seg_lr, double beta1, double beta2, float eps,
                              int zero_grads, int* state, const float* skip_flag, void* stream) {
  DIMO_REQUIRE(n >= 0 && (n & 3) == 0, "n must be a multiple of 4 (pad the flat buffer)");
  DIMO_REQUIRE(nseg >= 1 && nseg <= ADAM_MAX_SEGS, "1..64 learning-rate segments");
  if (n == 0) return 0;
  AdamSegs segs;
  segs.n = nseg;
  for (int i = 0; i <= nseg; ++i) {
    segs.begin[i] = seg_begin_host[i];
    DIMO_REQUIRE((segs.begin[i] & 3) == 0, "segment offsets must be multiples of 4");
    DIMO_REQUIRE(i == 0 || segs.begin[i] >= segs.begin[i - 1], "segment offsets must be ascending");
  }
  DIMO_REQUIRE(segs.begin[0] == 0 && segs.begin[nseg] == n, "segments must cover [0, n)");
  const int64_t n4 = n / 4;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n4 + ADAM_THREADS - 1) / ADAM_THREADS;
  const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
  adam_kernel<<<grid, ADAM_THREADS, 0, (cudaStream_t)stream>>>(
      n4, reinterpret_cast<float4*>(params), reinterpret_cast<float4*>(grads), reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), seg_lr, segs, beta1, beta2, eps, zero_grads, state, skip_flag);
  DIMO_CHECK_LAUNCH();
  return 0;
}

extern "C" int dimo_transpose_grouped(int n, const int* rows_host, const int* cols_host, const float* const* src_host,
                                      float* const* dst_host, void* stream) {
  DIMO_REQUIRE(n >= 0 && n <= TR_MAX, "at most 16 matrices per call");
  if (n == 0) return 0;
  TransposeArgs a;
  a.n = n;
  int tiles = 0;
  for (int k = 0; k < n; ++k) {
    DIMO_REQUIRE(rows_host[k] > 0 && cols_host[k] > 0, "empty matrix");
    a.rows[k] = rows_host[k]; a.cols[k] = cols_host[k];
    a.src[k] = src_host[k]; a.dst[k] = dst_host[k];
    a.tile0[k] = tiles;
    tiles += ceil_div(rows_host[k], TR_TILE) * ceil_div(cols_host[k], TR_TILE);
  }
  a.tile0[n] = tiles;
  transpose_grouped_kernel<<<tiles, 256, 0, (cudaStream_t)stream>>>(a);
  DIMO_CHECK_LAUNCH();
  return 0;
}
