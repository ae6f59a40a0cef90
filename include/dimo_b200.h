/* libdimo_b200.so -- C ABI of the B200-native DIMO deform -> raster -> loss hot path.
 *
 * The reference has no C ABI: its boundary for this path is five Python extension modules
 * (pybind11 / torch extensions) plus PyTorch glue.  Each entry point below names the
 * reference interface it replaces (file:line relative to the reference tree); the Python
 * shims in dimo_b200/shims/ bind these with ctypes and re-expose the reference's own module
 * names and call signatures.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the caller owns every buffer (inputs, outputs, workspace, saved-for-backward); the
 *    library never allocates device memory and never touches the default stream;
 *  - `stream` is a cudaStream_t passed as void*;
 *  - return 0 on success, <0 on error (message via dimo_last_error(), thread-local);
 *  - fp32 everywhere unless noted; integers called out per argument;
 *  - `*_bstride` is the element stride between consecutive frames of a batched argument;
 *    0 means "shared by all B frames".
 */
#ifndef DIMO_B200_H
#define DIMO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIMO_ABI_VERSION 2

/* Per-frame camera block, DIMO_CAM_FLOATS floats on the device:
 *   [0:16)  viewmatrix   (world_view_transform, row-vector convention)   renderer/latent_gs_renderer.py:960
 *   [16:32) projmatrix   (full_proj_transform)                            :969
 *   [32:35) campos       (camera_center)                                  :970
 *   [35]    tanfovx  [36] tanfovy                                         :1129-1130
 *   [37:40) bg colour                                                     :1139
 * i.e. the fields of GaussianRasterizationSettings (:1133-1146) that vary per frame. */
#define DIMO_CAM_FLOATS 40

/* Per-(frame, Gaussian) projected record ("splat" / blend record), DIMO_SPLAT_FLOATS floats (64 B = one bulk
 * async copy into shared memory):
 *   x, y (pixel centre), a2, b2 | c2, opacity, pthr2, r | g, b, depth, nx | ny, nz, own index (u32 bits), 0
 * with a2 = -0.5*log2(e)*conic_a, b2 = -log2(e)*conic_b, c2 = -0.5*log2(e)*conic_c (alpha = opacity * 2^p2) and
 * pthr2 = -log2(255*opacity) - 0.01 (pairs below it cannot reach alpha >= 1/255).
 * Gradient records (dL_dsplats) use: x, y, conic_a, conic_b | conic_c, opacity, r, g | b, depth, nx, ny | nz,0,0,0 */
#define DIMO_SPLAT_FLOATS 16

int         dimo_abi_version(void);
const char* dimo_last_error(void);
/* device properties the host side sizes grids with: out[0]=SM count, out[1]=max smem/block optin,
 * out[2]=compute capability major*10+minor */
int         dimo_device_info(int* out3_host);

/* ---------------------------------------------------------------------------------------------
 * Rasteriser  (replaces diff_gauss.GaussianRasterizer.forward/backward, call site
 * renderer/latent_gs_renderer.py:1147,1256-1266, and diff_gaussian_rasterization
 * .GaussianRasterizer, :1163,1268-1277).  Batched over B frames that share W,H.
 * ------------------------------------------------------------------------------------------- */

/* bytes of stage 2's scratch buffer (per-chunk tile histograms, tile totals, tile bases) */
size_t dimo_raster_bin_temp_bytes(int B, int N, int W, int H);
/* Instance format of a launch set.  Returns value_bits > 0 when a tile key (B*tiles + 1 codes) and the index of a
 * Gaussian within its frame (N codes) fit one 32-bit word: stage 2 then writes SINGLE words
 * (key << value_bits) | index into vals_sorted (keys_sorted is not touched and may be NULL).  0: separate 32-bit keys
 * and B*N indices.
 * Pass the same value to dimo_raster_blend_fwd / _bwd, which decode  record = (word & mask) + frame * N. */
int dimo_raster_packed_value_bits(int B, int N, int W, int H);

/* Stage 1: per-Gaussian projection + tile rectangles, total instance count, per-frame depth sort of the splats
 * (one thread-block cluster per frame; hand-written stable LSD radix sort on the 32-bit depth bits).
 *   splats [B*N,16] f32, radii [B*N] i32, tiles_touched [B*N] u32, rects [B*N,2] u32 (x0 | y0 << 16, x1 | y1 << 16:
 *   the tile rectangle [x0,x1) x [y0,y1) of every splat; empty for culled ones), all in Gaussian order;
 *   sort_scratch [3*B*N] u32, perm [2*B*N] u32 scratch: perm + B*N is, per frame, the splat indices b*N + i in
 *   ascending (view depth, i) order, culled splats last (input of stage 2);
 *   total_count [1] u64 (device): number of tile instances R of the launch set.
 *   shs [N,sh_coeffs,3] (or NULL) / colors_precomp [N,3] (or NULL): exactly one non-NULL.
 *   R_host: if non-NULL the stream is synchronised and the total instance count is stored there.
 *   frame_src [B] i32 (device, or NULL = identity): frame b reads means3D / rotations block frame_src[b] -- the
 *   deformation depends on (motion, t) only, so the frames of a step that differ in the view alone share one
 *   block (renderer/latent_gs_renderer.py:1191-1219 is recomputed per render upstream).
 *   act_flags: bit 0 = `scales` are the model's raw log-scales (exp applied in the kernel), bit 1 = `opacities` are
 *   logits (sigmoid applied in the kernel): GaussianModel.get_scaling / get_opacity (:257-265, 340-355) folded in;
 *   the backward then returns gradients w.r.t. the raw parameters. */
int dimo_raster_preprocess(
    int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags,
    const float* cams, const int32_t* frame_src,
    const float* means3D, int64_t means3D_bstride,
    const float* scales, int64_t scales_bstride,
    const float* rotations, int64_t rotations_bstride,
    const float* opacities, int64_t opacities_bstride,
    const float* shs, int64_t shs_bstride,
    const float* colors_precomp, int64_t colors_bstride,
    float* splats, int32_t* radii, uint32_t* tiles_touched, uint32_t* rects,
    uint32_t* sort_scratch, uint32_t* perm, uint64_t* total_count,
    int64_t* R_host, void* stream);

/* Stage 2: per-tile instance lists in depth order (stable counting sort by tile) + per-tile ranges.
 *   rects, perm_sorted (= perm + B*N) from stage 1; vals_sorted [R] u32 (index into the B*N splat records, or packed
 *   words), keys_sorted [R] u32 (frame*tiles + tile; only in the unpacked format), ranges [B*tiles,2] u32 = [begin,
 *   end) of every tile's list; temp: dimo_raster_bin_temp_bytes() bytes.
 *   R is the number of instance SLOTS.  count_overflow == NULL: R is the exact count read back from stage 1.
 *   count_overflow != NULL (i32[2], device, [1] zeroed by the caller once): "capacity mode" for sync-free /
 *   CUDA-graph use -- R is a capacity, [0] receives the true count and [1] is set to 1 if it exceeded R (instances
 *   whose slot lies beyond R were dropped and the ranges are clamped to R: the caller must re-run with a larger
 *   capacity). */
int dimo_raster_bin(
    int B, int N, int W, int H, int64_t R,
    const uint32_t* rects, const uint32_t* perm_sorted,
    uint32_t* keys_sorted, uint32_t* vals_sorted,
    void* temp, size_t temp_bytes,
    uint32_t* ranges, int32_t* count_overflow, void* stream);

/* Stage 3: per-tile front-to-back blend; tile t walks vals_sorted[ranges[t].x .. ranges[t].y) and gathers the
 * splat records by index.
 *   out_color [B,3,H,W], out_depth [B,1,H,W], out_normal [B,3,H,W], out_alpha [B,1,H,W],
 *   final_T [B,H,W] f32, n_contrib [B,H,W] i32. */
int dimo_raster_blend_fwd(
    int B, int N, int W, int H, int value_bits, const float* cams, const float* splats, const uint32_t* vals_sorted,
    const uint32_t* ranges, float* out_color, float* out_depth, float* out_normal, float* out_alpha,
    float* final_T, int32_t* n_contrib, void* stream);

/* Backward of stage 3: dL_dsplats [B*N,16] (zeroed here, then accumulated), gradient-record layout above.
 * dL_ddepth and dL_dnormal may BOTH be NULL (no loss term reads depth / normal, e.g. the MSE + SSIM + mask step):
 * the four channels are then compiled out of the pixel loop and of the warp reduction. */
int dimo_raster_blend_bwd(
    int B, int N, int W, int H, int value_bits, const float* cams, const float* splats, const uint32_t* vals_sorted,
    const uint32_t* ranges, const float* final_T, const int32_t* n_contrib,
    const float* dL_dcolor, const float* dL_ddepth, const float* dL_dnormal, const float* dL_dalpha,
    float* dL_dsplats, void* stream);

/* Backward of stage 1: per-frame gradients (dense, no atomics):
 *   dL_dmeans3D [B,N,3], dL_dmeans2D [B,N,3] (NDC units, z=0; may be NULL: not written), dL_dscales [B,N,3],
 *   dL_drotations [B,N,4], dL_dopacities [B,N], dL_dshs [B,N,sh_coeffs,3] or dL_dcolors [B,N,3].
 * reduce_shared != 0 (the training step: scales, opacities and shs shared by all frames, i.e. batch strides 0, SH
 * colours): dL_dscales [N,3], dL_dopacities [N] and dL_dshs [N,sh_coeffs,3] are the SUMS over the B frames, formed in
 * registers / shared memory in a fixed order (deterministic) instead of being written per frame and folded by
 * dimo_segment_sum afterwards.  reduce_shared == 2: the three sums are ADDED to what the buffers hold (the caller's
 * gradient buffers) instead of overwriting them. */
int dimo_raster_preprocess_bwd(
    int B, int N, int W, int H, int sh_degree, int sh_coeffs, float scale_modifier, int act_flags,
    const float* cams, const int32_t* frame_src,
    const float* means3D, int64_t means3D_bstride,
    const float* scales, int64_t scales_bstride,
    const float* rotations, int64_t rotations_bstride,
    const float* opacities, int64_t opacities_bstride,
    const float* shs, int64_t shs_bstride,
    const int32_t* radii, const float* dL_dsplats,
    float* dL_dmeans3D, float* dL_dmeans2D, float* dL_dscales, float* dL_drotations,
    float* dL_dopacities, float* dL_dshs, float* dL_dcolors, int reduce_shared, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Nearest neighbours (replaces knn_cuda.KNN(k, transpose_mode=True), main_train_dimo.py:505-506,
 * and simple_knn._C.distCUDA2, renderer/latent_gs_renderer.py:426).
 * ------------------------------------------------------------------------------------------- */
/* ref [M,3], query [N,3] -> dist [N,k] f32 (Euclidean, ascending), idx [N,k] i64.  k <= 8. */
int dimo_knn(int M, int N, int k, const float* ref, const float* query,
             float* dist, int64_t* idx, void* stream);
/* points [N,3] -> mean squared distance to the 3 nearest other points, [N]. */
int dimo_dist3nn(int N, const float* points, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Point-set ops of the key-point annealing and the step regularisers.  The
 * reference reaches them through pip packages that are not part of its tree (pytorch3d, chamferdist:
 * parity unpinned, semantics restated in csrc/points.cu).
 * ------------------------------------------------------------------------------------------- */
/* replaces pytorch3d.ops.sample_farthest_points(points[B,N,3], K=K)[1] (GUI.FPS, main_train_dimo.py:511-515,
 * main_test_dimo.py:166-170): idx [B,K] i64, first pick = `start` (pytorch3d: 0), then repeatedly the point with
 * the largest squared distance to the picked set (ties -> lower index).  min_dist_scratch [B,N] f32. */
int dimo_fps(int B, int N, int K, int start, const float* points, float* min_dist_scratch, int64_t* idx,
             void* stream);
/* replaces pytorch3d.ops.ball_query(p1[B,P1,3], p2[B,P2,3], K=K, radius=radius) (utils/deform_utils.py:128):
 * idx [B,P1,K] i64 = the first K points of p2 IN INDEX ORDER with squared distance < radius^2 (-1 padded),
 * dists [B,P1,K] f32 = those squared distances (0 padded). */
int dimo_ball_query(int B, int P1, int P2, int K, float radius, const float* p1, const float* p2,
                    int64_t* idx, float* dists, void* stream);
/* replaces chamferdist.ChamferDistance()(src[1,N,3], tgt[1,M,3]) (main_train_dimo.py:298-299; defaults: forward
 * direction only, squared distances summed over the source points):
 *   d2 [N] f32, nn [N] i32 = nearest target per source point; *sum += sum_i d2_i (may be NULL);
 *   *loss_acc += lw * sum_i d2_i (may be NULL: the step's loss scalar, like dimo_ssim_fwd).
 * bwd: d_src [N,3] = 2 * gw * g * (src_i - tgt_nn(i)) (g = *g_scalar on the device, NULL = 1);
 *      d_tgt [M,3] (may be NULL; must be zeroed by the caller) accumulates the opposite sign. */
int dimo_chamfer_fwd(int N, int M, const float* src, const float* tgt, float* d2, int32_t* nn, float* sum,
                     float* loss_acc, float lw, void* stream);
int dimo_chamfer_bwd(int N, const float* src, const float* tgt, const int32_t* nn, const float* g_scalar,
                     float gw, float* d_src, float* d_tgt, void* stream);
/* ARAP term (Renderer.arap_loss_v2, renderer/latent_gs_renderer.py:1081-1094; utils/deform_utils.py:115-232).
 *   nodes [T,M,3] f32: frame 0 is the source, frames 1..T-1 the targets.
 * connectivity: nbr [M,K] i64 (-1 padded, ascending) = vertices inside the ball of `radius` around i in EVERY frame
 *   (per frame the first K+1 hits in index order, the first one dropped -- cal_connectivity_from_points_v2), count [M]
 *   i32 (may be NULL).  K <= 15.
 * energy: *energy = sum_t sum_i mult_i sum_k |(p^t_i - p^t_j) - R^t_i (p^0_i - p^0_j)|^2 with R the Kabsch rotation of
 *   the vertex's edge fan (cal_arap_error with unit edge weights); mult [M] f32 = how often vertex i is in the
 *   reference's random vertex sample (NULL = 1 each); grad [T,M,3] (may be NULL) = d energy / d nodes with R held
 *   constant.  Both outputs are overwritten. */
int dimo_arap_connectivity(int T, int M, int K, float radius, const float* nodes, int64_t* nbr, int32_t* count,
                           void* stream);
int dimo_arap_energy(int T, int M, int K, const float* nodes, const int64_t* nbr, const float* mult, float* energy,
                     float* grad, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Deformation: positional encoding + TimeNet MLP (renderer/latent_gs_renderer.py:184-235,
 * src/pos_enc.py:6-54) and K-neighbour linear-blend skinning (:1191-1219).
 * ------------------------------------------------------------------------------------------- */

/* Generic fused linear layer, the building block of TimeNet:
 *   Y[R,No] = act( X[R,K] * W[No,K]^T + bias[No] ),  row strides ldx/ldy in floats, relu 0/1. */
int dimo_linear_fwd(int R, int K, int No, const float* X, int64_t ldx, const float* Wt,
                    const float* bias, float* Y, int64_t ldy, int relu, void* stream);
/* dX[R,K] (=|+=) (dY * [Y>0]) * W ; accumulate 0/1 ; Y == NULL -> no ReLU mask (head layers). */
int dimo_linear_bwd_data(int R, int K, int No, const float* dY, int64_t lddy, const float* Y, int64_t ldy,
                         const float* Wt, float* dX, int64_t lddx, int accumulate, void* stream);
/* dW[No,K] += (dY * [Y>0])^T * X ; db[No] += column sums of (dY * [Y>0]).  dW/db must be zeroed (or hold
 * prior gradients) by the caller. */
int dimo_linear_bwd_weight(int R, int K, int No, const float* dY, int64_t lddy, const float* Y, int64_t ldy,
                           const float* X, int64_t ldx, float* dW, float* db, void* stream);

/* Tensor-core (tcgen05, 3xTF32-compensated) variant of the fused linear layer:
 *   Y[R,No] (=|+=) act( (X * [mask>0]) * W[No,K]^T + bias ).  mask may be NULL (same shape/stride rules as X, ldm);
 *   K, ldx (and ldm) multiples of 4 floats, X/W/mask 16-byte aligned.  Used for TimeNet forward and, with the
 *   transposed weights, for the data gradient. */
int dimo_linear_tc(int R, int K, int No, const float* X, int64_t ldx, const float* mask, int64_t ldm,
                   const float* Wt, const float* bias, float* Y, int64_t ldy, int relu, int accumulate,
                   void* stream);
/* Tensor-core weight gradient: dW[No,K] += (dY * [mask>0])^T * X, db[No] += column sums of (dY * [mask>0]);
 *   No, K, lddy, ldx (ldm) multiples of 4 floats; 16-byte aligned operands.  dW/db accumulate (caller zeroes). */
int dimo_linear_wgrad_tc(int R, int K, int No, const float* dY, int64_t lddy, const float* mask, int64_t ldm,
                         const float* X, int64_t ldx, float* dW, float* db, void* stream);
/* The same for n <= 12 independent layers (same R) in ONE launch: every argument is a host array of length n. */
int dimo_linear_wgrad_tc_grouped(int n, int R, const int* K, const int* No, const float* const* dY,
                                 const int64_t* lddy, const float* const* mask, const int64_t* ldm,
                                 const float* const* X, const int64_t* ldx, float* const* dW, float* const* db,
                                 void* stream);
/* ---------------------------------------------------------------------------------------------
 * TimeNet as ONE call per direction (renderer/latent_gs_renderer.py:205-235: pos-enc + 12 F.linear; autograd backward):
 * TMA-fed tcgen05 GEMMs on operands that are stored pre-split (3xTF32 hi / lo) in the tensor core's shared-memory
 * layout (csrc/timenet_tc.cu).  G (motion, t) groups x M points = R rows; L = latent width (72 + L <= 128).
 *   W_host / b_host / dW_host / db_host: HOST arrays of 12 DEVICE pointers in the order deformnet.0..7, pts_layers.0,
 *   pts_layers.2, rot_layers.0, rot_layers.2 (weights [out, in] row-major).
 *   workspace: dimo_timenet_workspace_bytes(G, M, L) bytes, 256-byte aligned; written by _fwd (packed weights, every
 *   activation as split tiles + transposed split tiles) and consumed by _bwd -- keep it alive in between.
 *   _fwd: dxyz [R,3], dquat [R,4].
 *   _bwd: g_dxyz / g_dquat upstream gradients; dW / db are ACCUMULATED into (caller zeroes; they may be views of the
 *   flat gradient buffer); dpts [M,3] / dlatents [G,L] (may be NULL) are accumulated into as well (caller zeroes).
 * ------------------------------------------------------------------------------------------- */
size_t dimo_timenet_workspace_bytes(int G, int M, int L);
int dimo_timenet_layout(int G, int M, int L, int64_t* out12_host);
int dimo_timenet_last_launches(void);           /* kernel launches of the latest dimo_timenet_fwd / _bwd call (5 / 4 chained, 14 / 13 per layer) */
int dimo_timenet_debug_stamps(void* dev_buf);   /* bring-up: 64 x 8 u64 of per-CTA phase time stamps, NULL = off */
int dimo_timenet_fwd(int G, int M, int L, const float* pts, const float* times, const float* latents,
                     const float* const* W_host, const float* const* b_host, void* workspace, size_t workspace_bytes,
                     float* dxyz, float* dquat, void* stream);
int dimo_timenet_bwd(int G, int M, int L, const float* const* W_host, void* workspace, size_t workspace_bytes,
                     const float* g_dxyz, const float* g_dquat, float* const* dW_host, float* const* db_host,
                     float* dpts, float* dlatents, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Deterministic gradient accumulation (hard part 4).  Off (default): gradients that several CTAs
 * add into are summed with fp32 atomics / vector reductions -- run-to-run differences of ~1e-7 of the tensor scale.
 * On: every such accumulation target of dimo_raster_blend_bwd (dL_dsplats), dimo_lbs_bwd (its six outputs),
 * dimo_timenet_bwd (dW / db / dpts / dlatents), dimo_timenet_embed_bwd and dimo_linear_bwd_weight must be an INT64
 * buffer with the same element count (zeroed by the caller); addends are rounded to multiples of 2^-28 and summed with
 * native 64-bit integer reductions (associative: bit-identical results in any arrival order; range +-3.4e10).
 * dimo_fixed_to_float converts such a buffer to fp32 (dst = or += src * 2^-28).  The scalar loss sums keep fp32
 * atomics (they feed no gradient).
 * ------------------------------------------------------------------------------------------- */
int dimo_debug_max_sort_clusters(void);   /* debugging: co-resident depth-sort clusters */
int dimo_set_deterministic(int on);
int dimo_get_deterministic(void);
int dimo_fixed_to_float(int64_t n, const void* src_i64, float* dst, int accumulate, void* stream);

/* bring-up knobs (0: swap LBO/SBO, 1: single-pass TF32, 2: wgrad CTA target, 3: blend gather via 16-byte
 * cp.async instead of 64-byte bulk copies, 4 / 5: records per stage of the blend backward / forward, 64 or 128, 6: 1 = never pack
 * instances into single words); not part of the stable ABI */
int dimo_tc_debug_set(int key, int value);

/* TimeNet input embedding h0[R,104] = [posenc(x,10) | posenc(t,6) | latent]  (pos_enc.py:35-36,
 * latent_gs_renderer.py:223-225).  Row r belongs to group g = r / rows_per_group and reads
 * times[g], latents[g*L..].  pts [rows_per_group,3] shared by all groups. */
int dimo_timenet_embed_fwd(int G, int rows_per_group, int L, const float* pts, const float* times,
                           const float* latents, float* h0, int64_t ldh, void* stream);
/* backward of the embedding: dpts [rows_per_group,3] += , dlatents [G,L] += .  h0 is the forward output (its
 * sin/cos columns are reused instead of being recomputed); h0 and dh0 share the row stride ldh. */
int dimo_timenet_embed_bwd(int G, int rows_per_group, int L, const float* h0, const float* dh0, int64_t ldh,
                           float* dpts, float* dlatents, void* stream);

/* LBS skinning + activations for B (motion,t) frames over N Gaussians with K neighbours.
 *   xyz [N,3], rot [N,4], idx [N,K] i64, dist [N,K], c_xyz [M,3], c_radius_raw [M] (log-radius),
 *   dxyz [B,M,3], dquat [B,M,4]  ->  means3D [B,N,3], rotations [B,N,4] (normalised). */
int dimo_lbs_fwd(int B, int N, int M, int K, const float* xyz, const float* rot, const int64_t* idx,
                 const float* dist, const float* c_xyz, const float* c_radius_raw,
                 const float* dxyz, const float* dquat, float* means3D, float* rotations, void* stream);
/* backward: accumulates (+=) into dxyz_c [N,3], drot_c [N,4], dc_xyz [M,3], dc_radius_raw [M],
 * ddxyz [B,M,3], ddquat [B,M,4].  All outputs must be zeroed (or hold prior grads) by the caller. */
int dimo_lbs_bwd(int B, int N, int M, int K, const float* xyz, const float* rot, const int64_t* idx,
                 const float* dist, const float* c_xyz, const float* c_radius_raw,
                 const float* dxyz, const float* dquat, const float* dL_dmeans3D, const float* dL_drotations,
                 float* dxyz_c, float* drot_c, float* dc_xyz, float* dc_radius_raw,
                 float* ddxyz, float* ddquat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Image loss (replaces fused_ssim.fused_ssim, main_test_dimo.py:979, and src/loss.py:144-178 ssim /
 * l1_loss + F.mse_loss, main_train_dimo.py:333-344).
 *   img1,img2 [B,C,H,W]; sums[3] (f32, zeroed by callee) = { sum ssim_map, sum |a-b|, sum (a-b)^2 }.
 *   dm [3,B,C,H,W]: the three partial-derivative maps kept for backward (may be NULL for eval).
 *   clamp01: img1 is clamped to [0,1] on load (the reference clamps the render before the losses,
 *   renderer/latent_gs_renderer.py:1279) and the backward zeroes the gradient outside [0,1].
 * ------------------------------------------------------------------------------------------- */
int dimo_ssim_fwd(int B, int C, int H, int W, int clamp01, const float* img1, const float* img2,
                  float* sums, float* dm, const float* mse_frame_w, float* loss_acc, float lw_ssim, float lw_l1,
                  float lw_mse, void* stream);
/*   mse_frame_w [B] (device, or NULL = 1): per-frame weight of the squared-error sum -- the reference weights the MSE
 *   of non-reference views/frames by 0.5 (main_train_dimo.py:331-336); sums[2] is then the weighted sum.
 *   loss_acc (device scalar, or NULL): loss_acc += lw_ssim*sums[0] + lw_l1*sums[1] + lw_mse*sums[2], accumulated by
 *   the kernel itself so the step's scalar loss needs no elementwise launches; NOT zeroed by the callee. */

/* dL_dimg1 [B,C,H,W] = g * (w_ssim * d(sum ssim_map)/dimg1 + w_l1 * d(sum|a-b|)/dimg1
 *                           + w_mse * mse_frame_w[b] * d(sum (a-b)^2)/dimg1).
 * Weights are host floats: the caller folds 1/numel and the loss weights in.  g = *g_dev, the upstream gradient of
 * the scalar loss as a DEVICE scalar (NULL = 1).  w_ssim == 0 skips the convolutions (dm may then be NULL). */
int dimo_ssim_bwd(int B, int C, int H, int W, int clamp01, const float* img1, const float* img2, const float* dm,
                  float w_ssim, float w_l1, float w_mse, const float* mse_frame_w, const float* g_dev,
                  float* dL_dimg1, void* stream);

/* out[u,:] = sum of the rows s of in [S,n] with seg[s] == u (seg [S] i32 device; NULL: all rows -> segment 0), u < U,
 * rows added in ascending order.  Folds per-frame gradients ([B,N,*], dimo_raster_preprocess_bwd) onto shared inputs:
 * the (motion, t) block named by frame_src, or a per-Gaussian parameter shared by all frames.  S <= 1024. */
int dimo_segment_sum(int S, int U, int64_t n, const int32_t* seg, const float* in, float* out, void* stream);

/* sum (a-b)^2 over n floats -> sum[0] (zeroed by the callee); optional
 * loss_acc += lw * sum.  The mask term F.mse_loss(render_alpha, gt_mask), main_train_dimo.py:350; its gradient is
 * dimo_ssim_bwd with w_ssim = 0. */
int dimo_sqdiff_sum(int64_t n, const float* a, const float* b, float* sum, float* loss_acc, float lw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Step regularisers on the rendered maps: compute_edge_aware_smoothness_loss and
 * compute_bilateral_normal_smoothness_loss, src/loss.py:64-107, as main_train_dimo.py:363-372 applies them.
 *   rgb [B,3,H,W] (clamp01: clamped to [0,1] on load, gradient masked outside), depth [B,1,H,W], normal [B,3,H,W].
 *   sums4 (zeroed by the callee) = { sum over x-pairs of |dd| e^-gi, same over y-pairs,
 *                                    sum over x-pairs and channels of sqrt(1 + (|dn| e^-3gi)^2), same over y-pairs }
 *   with gi = mean_c |rgb_p - rgb_q|; the reference's means divide by B*H*(W-1) / B*(H-1)*W (x3 for the normals).
 *   wdx, wdy, wnx, wny: weight of each sum in the loss (the caller folds lambda and the normalisers in);
 *   loss_acc (device scalar or NULL) += the weighted total.
 * Backward: d_depth, d_normal are overwritten; d_rgb is overwritten or (accumulate_rgb != 0) added to -- the
 * exp(-gi) factor depends on the rendered image, so the image receives a gradient too; g_dev = upstream gradient of
 * the scalar loss (device scalar, NULL = 1).
 * ------------------------------------------------------------------------------------------- */
int dimo_smooth_fwd(int B, int H, int W, int clamp01, const float* rgb, const float* depth, const float* normal,
                    float* sums4, float* loss_acc, float wdx, float wdy, float wnx, float wny, void* stream);
int dimo_smooth_bwd(int B, int H, int W, int clamp01, const float* rgb, const float* depth, const float* normal,
                    float wdx, float wdy, float wnx, float wny, const float* g_dev, float* d_rgb, int accumulate_rgb,
                    float* d_depth, float* d_normal, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Optimizer: torch.optim.Adam(groups, lr=0.0, eps=1e-15) of GaussianModel.training_setup
 * (renderer/latent_gs_renderer.py:453-476), stepped + zero_grad'ed at main_train_dimo.py:416-417.
 *   params / grads / exp_avg / exp_avg_sq: flat fp32 buffers of n floats (n % 4 == 0) with identical layout;
 *   nseg learning-rate segments: seg_begin_host [nseg+1] (host array, element offsets, multiples of 4, covering
 *   [0,n)), seg_lr [nseg] (DEVICE array, so learning-rate schedules work under CUDA-graph replay);
 *   state [4] i32 (device): [0] = number of updates applied so far (bias correction uses state[0]+1 and the
 *   kernel increments it), [1] scratch (must start 0), [3] number of skipped steps;  zero_grads != 0: grads are
 *   cleared in the same pass.
 *   skip_flag (device float, may be NULL, must not alias the four buffers): non-zero => this step's gradients are
 *   discarded (cleared, no update, step count unchanged).  The training step points it at the all-reduced
 *   "instance capacity overflowed" word, so a CUDA-graph replay whose rasteriser dropped instances on any rank never
 *   reaches the parameters (dimo_raster_bin, count_overflow).
 * ------------------------------------------------------------------------------------------- */
int dimo_adam_step(int64_t n, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int nseg,
                   const int64_t* seg_begin_host, const float* seg_lr, double beta1, double beta2, float eps,
                   int zero_grads, int* state, const float* skip_flag, void* stream);

/* dst_k[c][r] = src_k[r][c] for n <= 16 row-major matrices in ONE launch (W^T operands of the tensor-core
 * data-gradient GEMMs); all four arguments are host arrays of length n. */
int dimo_transpose_grouped(int n, const int* rows_host, const int* cols_host, const float* const* src_host,
                           float* const* dst_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Ground truth resident in HBM: one gather + convert + bilinear-resample launch per step instead
 * of the reference's per-frame upload + F.interpolate(mode="bilinear", align_corners=False)
 * (main_train_dimo.py:283-284, 305-313).
 *   store [F,4,Hs,Ws] (u8: value / 255, or f32; channels R,G,B,mask), slots [S] i32 (device) = frame slots to fetch,
 *   rgb [S,3,Ho,Wo] f32, mask [S,1,Ho,Wo] f32.
 * ------------------------------------------------------------------------------------------- */
int dimo_gt_fetch(int S, int Hs, int Ws, int Ho, int Wo, int store_is_u8, const void* store,
                  const int32_t* slots, float* rgb, float* mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIMO_B200_H */
