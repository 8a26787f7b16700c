/*
 * bdm_b200.h -- C-ABI of libbdm_b200.so: the B200 (sm_100a) implementation of BDM's per-step
 * denoising hot path (PVCNN/PVD point-voxel ops, PC^2 projection conditioning, evaluation kNN).
 *
 * This is the drop-in boundary.  The reference reaches its CUDA kernels through one pybind module
 * `_pvcnn_backend` (experiments/model/pvcnn/modules/functional/src/bindings.cpp:10-37, byte-identical
 * copy under experiments/pvd/...) whose C++ wrappers call one plain launcher per op, declared in the
 * reference's *.cuh files.  Each entry point below replaces exactly one of those launchers: same
 * argument meaning and order (sizes first, then pointers), plus a stream, plus -- for ops that
 * sort -- a caller-owned workspace.  No torch types, no ownership transfer, no hidden allocation,
 * no host synchronisation: everything is enqueued on `stream` and returns immediately.
 *
 * Conventions (identical to the reference, src/utils.hpp:7-18 and the wrappers' layouts):
 *   - all pointers are DEVICE pointers to contiguous fp32 / int32 arrays on the current device;
 *   - channel-first layouts [B,C,N]; coordinate planes [B,3,N]; int32 index arithmetic;
 *   - outputs are fully written by the kernels: the caller does NOT need to zero them
 *     (the reference wrappers pre-zero every output with torch::zeros; we do not need that);
 *   - return value: 0 on success, >0 a cudaError_t raised by the launch, <0 a BDM_ERR_* argument
 *     error.  The library never calls exit() (the reference does: src/cuda_utils.cuh:28-37) and
 *     never returns success without having enqueued the work.
 */
#ifndef BDM_B200_H
#define BDM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *bdm_stream_t; /* == cudaStream_t */

#define BDM_OK 0
#define BDM_ERR_NULL_POINTER (-1)
#define BDM_ERR_BAD_SIZE (-2)
#define BDM_ERR_WORKSPACE_TOO_SMALL (-3)
#define BDM_ERR_MISALIGNED (-4)

#define BDM_ABI_VERSION 1

int bdm_abi_version(void);
/* human-readable text for a return code of any function below (static storage) */
const char *bdm_error_string(int code);

/* ---- avg_voxelize ----------------------------------------------------------------------------
 * replaces  void avg_voxelize(int b,int c,int n,int r,int r2,int r3,const int*coords,const float*feat,
 *                             int*ind,int*cnt,float*out)               (src/voxelization/vox.cuh:5-6)
 *   coords i32[b,3,n] in [0,r)   feat f32[b,c,n]
 *   ind i32[b,n]   cnt i32[b,r^3]   out f32[b,c,r^3]
 * workspace: bdm_avg_voxelize_workspace_bytes(b,n,r) bytes, 16-byte aligned, contents irrelevant. */
size_t bdm_avg_voxelize_workspace_bytes(int b, int n, int r);
int bdm_avg_voxelize(int b, int c, int n, int r, const int *coords, const float *feat, int *ind,
                     int *cnt, float *out, void *workspace, size_t workspace_bytes,
                     bdm_stream_t stream);
/* The coordinate half of Voxelization.forward (modules/voxelization.py:17-24) in one kernel: per shape, centre on the
 * mean, scale by twice the largest point norm (+ eps) when `normalize` (else (c + 1) / 2), shift, stretch to the grid,
 * clamp to [0, r-1] -> norm_coords f32[b,3,n]; vox_coords i32[b,3,n] = round-half-even of it.  Each elementwise step
 * is the torch op's fp32 arithmetic; the mean is the correctly rounded one (double accumulation). */
int bdm_voxelize_coords(int b, int n, int r, int normalize, float eps, const float *coords,
                        float *norm_coords, int *vox_coords, bdm_stream_t stream);
/* The two halves of bdm_avg_voxelize, for callers that voxelize several feature tensors over the
 * same coordinates (consecutive PVConv blocks of one stage do: modules/pvconv.py:91-97 is called 2-3
 * times per stage with unchanged coords).  bdm_voxel_plan does everything that depends only on the
 * coordinates (ind, cnt, the sorted plan kept in `workspace`); bdm_avg_voxelize_fill produces the
 * dense grid for one feature tensor from that plan (same b, n, r and workspace). */
int bdm_voxel_plan(int b, int n, int r, const int *coords, int *ind, int *cnt, void *workspace,
                   size_t workspace_bytes, bdm_stream_t stream);
int bdm_avg_voxelize_fill(int b, int c, int n, int r, const int *ind, const int *cnt,
                          const float *feat, float *out, const void *workspace,
                          size_t workspace_bytes, bdm_stream_t stream);
/* Sparse consumers of a voxelization (the first Conv3d of a PVConv block, modules/pvconv.py:75-76,91-97,
 * reads a grid that is ~95 % zeros at r=32):
 *   bdm_avg_voxelize_compact   out f32[b,c,n]: column j = average of the j-th occupied voxel of shape b
 *                              (ascending voxel id; bit-identical to the dense grid's entry), columns past
 *                              the shape's occupied count are zero.  Needs the plan (r^3 <= 32768, n <= 16384).
 *   bdm_sparse_conv3_gather    the 3x3x3 / stride 1 / zero-padded convolution's dense output from per-
 *                              occupied-voxel tap products taps f32[b,n,27,cout]
 *                              (taps[b,j,k,co] = sum_ci W[co,ci,kd,kh,kw] * compact[b,ci,j], k=(kd*3+kh)*3+kw,
 *                              one GEMM on the host side): out[b,co,x,y,z] = bias[co] + sum_k taps[b,
 *                              slot(x+kd-1,y+kh-1,z+kw-1),k,co], k ascending.  bias may be NULL.
 *                              r must be a power of two <= 32.  channels_last != 0 writes out f32[b,r^3,cout]
 *                              (NDHWC, what cuDNN's tensor-core Conv3d kernels work in) instead of
 *                              f32[b,cout,r^3].  stats (or NULL): f64[b,blocks,cout,2], blocks =
 *                              bdm_sparse_conv3_stats_blocks(r): per-channel (sum, sum of squares) of the bias-less
 *                              output per block of rows -- the GroupNorm statistics, made by the producer. */
int bdm_avg_voxelize_compact(int b, int c, int n, int r, const float *feat, float *out,
                             const void *workspace, size_t workspace_bytes, bdm_stream_t stream);
/* the same, also leaving the bit pattern of max|out| in *amax_bits (a device word, zeroed by the call first): the
 * dynamic fp16 scale of bdm_conv3_tc05_fill_planes(amax_ready = 1) without another pass over the tensor */
int bdm_avg_voxelize_compact_amax(int b, int c, int n, int r, const float *feat, float *out,
                                  const void *workspace, size_t workspace_bytes, unsigned *amax_bits,
                                  bdm_stream_t stream);
int bdm_sparse_conv3_stats_blocks(int r);
int bdm_sparse_conv3_gather(int b, int cout, int n, int r, const float *taps, const float *bias,
                            float *out, int channels_last, double *stats, const void *workspace,
                            size_t workspace_bytes, bdm_stream_t stream);
/* replaces avg_voxelize_grad (src/voxelization/vox.cuh:7-8): grad_y f32[b,c,s] -> grad_x f32[b,c,n] */
int bdm_avg_voxelize_grad(int b, int c, int n, int s, const int *ind, const int *cnt,
                          const float *grad_y, float *grad_x, bdm_stream_t stream);

/* ---- trilinear_devoxelize ---------------------------------------------------------------------
 * replaces  void trilinear_devoxelize(int b,int c,int n,int r,int r2,int r3,bool is_training,
 *              const float*coords,const float*feat,int*inds,float*wgts,float*outs)
 *                                                            (src/interpolate/trilinear_devox.cuh:5-8)
 *   coords f32[b,3,n] in [0,r-1]   feat f32[b,c,r^3]   outs f32[b,c,n]
 *   inds i32[b,8,n], wgts f32[b,8,n]: written iff is_training (may be NULL otherwise).
 * workspace: bdm_trilinear_devoxelize_workspace_bytes(b,n,r) bytes, 16-byte aligned (x-slice binning
 * of the points for the shared-memory fast path); NULL selects the generic gather kernel.
 * planned: 0 = bin the points into `workspace` first; 1 = `workspace` already holds the result of
 * bdm_trilinear_devoxelize_plan for these coords (consecutive PVConv blocks of a stage share it). */
size_t bdm_trilinear_devoxelize_workspace_bytes(int b, int n, int r);
int bdm_trilinear_devoxelize_plan(int b, int n, int r, const float *coords, void *workspace,
                                  size_t workspace_bytes, bdm_stream_t stream);
int bdm_trilinear_devoxelize(int b, int c, int n, int r, int is_training, const float *coords,
                             const float *feat, int *inds, float *wgts, float *outs,
                             void *workspace, size_t workspace_bytes, int planned,
                             bdm_stream_t stream);
/* inference devoxelization from a channels-last grid feat f32[b,r^3,c] (same arithmetic, outs f32[b,c,n]);
 * optional epilogue of the PVConv block: outs = devox * gate[b,c] + residual[b,c,n] (NULL = skip) */
int bdm_trilinear_devoxelize_cl(int b, int c, int n, int r, const float *coords, const float *feat,
                                const float *gate, const float *residual, float *outs, bdm_stream_t stream);
/* the same from the UN-normalised grid: coef f32[b][c][2] = (A, B) with y = act(x*A + B) (act = Swish when swish != 0;
 * bdm_groupnorm_cl_sums produces coef) is applied to every corner value before the interpolation -- bit-identical to
 * devoxelizing bdm_groupnorm_act_cl's output, which is then never written (modules/pvconv.py:82-83, 95-97) */
int bdm_trilinear_devoxelize_cl_norm(int b, int c, int n, int r, const float *coords, const float *feat,
                                     const float *coef, int swish, const float *gate, const float *residual,
                                     float *outs, bdm_stream_t stream);
/* replaces trilinear_devoxelize_grad (trilinear_devox.cuh:9-11): grad_x f32[b,c,r3] is zeroed here */
int bdm_trilinear_devoxelize_grad(int b, int c, int n, int r3, const int *inds, const float *wgts,
                                  const float *grad_y, float *grad_x, bdm_stream_t stream);

/* ---- sampling -------------------------------------------------------------------------------
 * replaces gather_features / gather_features_grad / furthest_point_sampling
 *                                                                  (src/sampling/sampling.cuh:4-9)
 *   gather:  out[b,c,j] = features[b,c,indices[b,j]]        features f32[b,c,n], indices i32[b,m]
 *   fps:     indices i32[b,m]; the reference's `distances` scratch argument is gone (running
 *            distances live in registers); workspace only needed when n > BDM_FPS_REGISTER_MAX_N. */
int bdm_gather_features(int b, int c, int n, int m, const float *features, const int *indices,
                        float *out, bdm_stream_t stream);
int bdm_gather_features_grad(int b, int c, int n, int m, const float *grad_y, const int *indices,
                             float *grad_x, bdm_stream_t stream);
#define BDM_FPS_REGISTER_MAX_N 8192
size_t bdm_furthest_point_sampling_workspace_bytes(int b, int n);
int bdm_furthest_point_sampling(int b, int n, int m, const float *coords, int *indices,
                                void *workspace, size_t workspace_bytes, bdm_stream_t stream);

/* ---- ball query -------------------------------------------------------------------------------
 * replaces  void ball_query(int b,int n,int m,float r2,int u,const float*centers_coords,
 *              const float*points_coords,int*neighbors_indices)  (src/ball_query/ball_query.cuh:4-6)
 *   r2 = radius*radius computed in fp32 by the caller (ball_query.cpp:24). */
int bdm_ball_query(int b, int n, int m, float r2, int u, const float *centers_coords,
                   const float *points_coords, int *neighbors_indices, bdm_stream_t stream);

/* ---- grouping ---------------------------------------------------------------------------------
 * replaces grouping / grouping_grad                              (src/grouping/grouping.cuh:4-7)
 *   out[b,c,m,u] = features[b,c,indices[b,m,u]] */
int bdm_grouping(int b, int c, int n, int m, int u, const float *features, const int *indices,
                 float *out, bdm_stream_t stream);
/* bdm_grouping writing channels [channel_offset, channel_offset + c) of a tensor out f32[b,out_channels,m,u]
 * and, when centers f32[b,c,m] is not NULL, subtracting the centre: the reference's BallQuery.forward
 * (modules/ball_query.py:23-33) groups coordinates and features separately, subtracts the centres with a
 * broadcast op and concatenates the two; with this entry point both groupings fill the concatenated tensor
 * directly (same values: one fp32 subtraction per coordinate). */
int bdm_grouping_into(int b, int c, int n, int m, int u, const float *features, const int *indices,
                      const float *centers, float *out, int out_channels, int channel_offset,
                      bdm_stream_t stream);
int bdm_grouping_grad(int b, int c, int n, int m, int u, const float *grad_y, const int *indices,
                      float *grad_x, bdm_stream_t stream);

/* ---- three nearest neighbours + interpolation ---------------------------------------------------
 * replaces three_nearest_neighbors_interpolate / _grad
 *                                                    (src/interpolate/neighbor_interpolate.cuh:4-14)
 *   points f32[b,3,n], centers f32[b,3,m], centers_features f32[b,c,m]
 *   indices i32[b,3,n], weights f32[b,3,n], out f32[b,c,n]
 * The search and the interpolation are also exported separately so that a caller interpolating
 * several feature tensors over the same coordinates (modules/pointnet.py:107-108 does, twice per FP
 * stage) searches once. */
int bdm_three_nearest_neighbors_interpolate(int b, int c, int m, int n, const float *points_coords,
                                            const float *centers_coords,
                                            const float *centers_features, int *indices,
                                            float *weights, float *out, bdm_stream_t stream);
int bdm_three_nn_search(int b, int n, int m, const float *points_coords,
                        const float *centers_coords, float *weights, int *indices,
                        bdm_stream_t stream);
int bdm_three_nn_interpolate(int b, int c, int m, int n, const float *centers_features,
                             const int *indices, const float *weights, float *out,
                             bdm_stream_t stream);
int bdm_three_nearest_neighbors_interpolate_grad(int b, int c, int n, int m, const float *grad_y,
                                                 const int *indices, const float *weights,
                                                 float *grad_x, bdm_stream_t stream);

/* ---- projection conditioning --------------------------------------------------------------------
 * replaces the pytorch3d PointsRasterizer call + feature scatter inside
 * PointCloudProjectionModel.surface_projection (experiments/model/projection_model.py:127-157),
 * batched over the reference's per-sample Python loop (:205-212).
 *   points f32[b,n,3]; R f32[b,3,3] (row-vector convention X_view = X R + T); T f32[b,3] (already
 *   multiplied by scale_factor, :136-137); focal f32[b,2], principal f32[b,2] in NDC;
 *   feat f32[b,C,H,W]; radius in NDC (0.0075).
 *   zbuf u64[b,H,W] scratch (caller-owned, contents irrelevant);
 *   pix i32[b,n]: lowest pixel index (row*W+col) won by the point, or -1;   out f32[b,n,C]. */
int bdm_surface_projection(int b, int n, int C, int H, int W, float radius, const float *points,
                           const float *R, const float *T, const float *focal,
                           const float *principal, const float *feat,
                           unsigned long long *zbuf, int *pix, float *out, bdm_stream_t stream);
/* same, with the (step-invariant) feature map already in channel-last layout feat_hwc f32[b,H,W,C]:
 * the per-point gather then reads C contiguous floats instead of C strided sectors. */
int bdm_surface_projection_hwc(int b, int n, int C, int H, int W, float radius, const float *points,
                               const float *R, const float *T, const float *focal,
                               const float *principal, const float *feat_hwc,
                               unsigned long long *zbuf, int *pix, float *out, bdm_stream_t stream);

/* fused conditioning input: the projected features go straight into channels
 * [ch_off, ch_off+C) of the channel-first tensor out_cf f32[b,c_total,n] that the denoiser consumes,
 * skipping the [b,n,C] intermediate, the concat and the transpose of get_input_with_conditioning
 * (projection_model.py:179-231, point_cloud_model.py:65). */
int bdm_surface_projection_cf(int b, int n, int C, int H, int W, float radius, const float *points,
                              const float *R, const float *T, const float *focal,
                              const float *principal, const float *feat_hwc, unsigned long long *zbuf,
                              int *pix, float *out_cf, int c_total, int ch_off, bdm_stream_t stream);

/* ---- evaluation nearest neighbour (fp64) ----------------------------------------------------------
 * replaces pytorch3d knn (K=1) inside chamfer_distance (experiments/evaluation/evaluation_cd.py:125)
 * and compute_pc_to_pc_dist (experiments/evaluation/evaluation_f1.py:90-98).
 *   src f64[b,n,3], tgt f64[b,m,3];  expanded=0: direct (x-y)^2 form; expanded=1: the F-score
 *   expansion form -2ab+|a|^2+|b|^2 clamped at 1e-12.
 *   dist f64[b,n] (min squared distance), idx i32[b,n] (argmin, lowest index on ties; may be NULL). */
int bdm_nn_f64(int b, int n, int m, int expanded, const double *src, const double *tgt,
               double *dist, int *idx, bdm_stream_t stream);
/* The same search with the metrics' reductions fused in (evaluation_cd.py:125 mean of the minima; evaluation_f1.py:104-106
 * fraction of minima below thr): part_sum f64[b,blocks] / part_cnt i32[b,blocks] receive, per block of source
 * points, the sum of the minima and the number below thr; blocks = bdm_nn_f64_reduce_blocks(b, n). */
int bdm_nn_f64_reduce_blocks(int b, int n);
int bdm_nn_f64_reduce(int b, int n, int m, int expanded, double thr, const double *src, const double *tgt,
                      double *part_sum, int *part_cnt, bdm_stream_t stream);

/* ---- reverse-diffusion update of one sampling step ---------------------------
 * replaces the eager torch ops of diffusers' DDPMScheduler.step as driven by experiments/model/model.py:182-194
 * (mode 0; published algorithm, parity unpinned) and of GaussianDiffusion.p_sample,
 * experiments/pvd/__init__.py:136-224 (mode 1), elementwise over n floats:
 *   mode 0:  x0 = (x - c0*eps) * c1        mode 1:  x0 = c0*x - c1*eps
 *   out = c2*x0 + c3*x  (+ c4*noise when t > 0),   (c0..c4) = table[t][0..4],   t = *t_dev clamped to [0,rows-1]
 * table f32[rows][8] lives on the device and so does the timestep: a CUDA graph holding this kernel walks the
 * schedule by itself.  Products and sums are rounded one by one (bit-identical to the torch sequence).
 * out may alias x; all four arrays 16-byte aligned. */
int bdm_sampler_update(long long n, int mode, const float *x, const float *eps, const float *noise,
                       const float *table, int rows, const int *t_dev, float *out, bdm_stream_t stream);

/* ---- fused self-attention of the PVConv attention block ------------------------------------------
 * replaces, for inference, the two torch.matmul + softmax of Attention.forward (modules/pvconv.py:36-63):
 *   out[b,c,i] = sum_j softmax_j( q[b,:,i] . k[b,:,j] ) * v[b,c,j]       q,k,v,out f32[b,c,t], un-scaled logits
 * fp32-equivalent arithmetic (operands split into two fp16 halves after a per-tensor power-of-two scaling,
 * three tensor-core products per term, fp32 accumulation and softmax); the [t,t] logits are never written.
 * Runs on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory, operands by TMA bulk
 * copy); a pre-pass writes the fp16 operand planes into the workspace.
 * c must be 64 and t a multiple of 128 (BDM_ERR_BAD_SIZE otherwise: the caller keeps the torch route for
 * other shapes).  workspace: bdm_attention_workspace_bytes(b,c,t) bytes, 256-byte aligned. */
size_t bdm_attention_workspace_bytes(int b, int c, int t);
int bdm_attention(int b, int c, int t, const float *q, const float *k, const float *v, float *out,
                  void *workspace, size_t workspace_bytes, bdm_stream_t stream);
/* The same attention fed by one fused projection (the q, k, v 1x1 convolutions of modules/pvconv.py:40-50 as a single
 * GEMM over channels-last activations): qkv f32[b,t,ld] holds q | k | v of a token in columns [0,c) [c,2c) [2c,3c),
 * bias f32[3c] (or NULL: the convolutions' biases, added on the way in), out f32[b,t,c] token-major.
 * ld >= 3c, ld % 4 == 0; same size rules and workspace as bdm_attention. */
int bdm_attention_qkv(int b, int c, int t, const float *qkv, int ld, const float *bias, float *out,
                      void *workspace, size_t workspace_bytes, bdm_stream_t stream);

/* ---- dense side: fused [conv bias +] GroupNorm [+ Swish] [+ reduction] ----------
 * replaces the bias add of the preceding conv, the nn.GroupNorm(8, C) -> Swish pair that follows every
 * conv of the point-voxel blocks (modules/shared_mlp.py:25-31, modules/pvconv.py:75-88, :59-61) and,
 * optionally, the reduction that consumes the result: max over the innermost U neighbours
 * (modules/pointnet.py:86) or the per-channel sums of the squeeze-excite gate (modules/se.py:19).
 *   y = act(group_norm(x + conv_bias[c])), biased variance, eps inside the sqrt;
 *   act = v*sigmoid(v) when swish != 0, identity otherwise.
 *   x f32[b,c,s] (s = product of the trailing dims), conv_bias / gamma / beta f32[c] or NULL.
 *   max_over_u == 0: y f32[b,c,s]; tile_sums (or NULL) f32[b*c, bdm_groupnorm_tiles(b,c,s)] receives
 *                    per-tile sums of y (sum them for the channel total).
 *   max_over_u == U (power of two, 4..128, divides s): y f32[b,c,s/U] = max over each run of U values.
 *                    (evaluated as max(act(largest input), act(smallest input)) of the run: the affine map is
 *                    monotone and Swish unimodal, so one of the two extremes carries the maximum)
 * workspace: bdm_groupnorm_workspace_bytes(b,c,s) bytes, 16-byte aligned. */
size_t bdm_groupnorm_workspace_bytes(int b, int c, long long s);
int bdm_groupnorm_tiles(int b, int c, long long s);
int bdm_groupnorm_act(int b, int c, long long s, int groups, float eps, int swish, int max_over_u,
                      const float *x, const float *conv_bias, const float *gamma, const float *beta,
                      float *y, float *tile_sums, void *workspace, size_t workspace_bytes,
                      bdm_stream_t stream);
/* channels-last flavour: x, y f32[b,s,c]; tile_sums f32[b,tiles,c] (tiles = bdm_groupnorm_cl_tiles) or NULL.
 * c must be a power of two in [16,256] (bdm_groupnorm_cl_supported). */
/* kernels launched by this thread's last bdm_groupnorm_act / bdm_groupnorm_act_cl call (1 or 2) */
int bdm_groupnorm_last_launches(void);
/* squeeze-excite gate (modules/se.py:8-19) from the per-channel sums above:
 * gate[b,c] = sigmoid(w2 . act(w1 . (sum_t sums / count))), act = ReLU (use_relu) or Swish; sums is addressed
 * as sums[b*stride_b + t*stride_t + c*stride_c]. */
int bdm_se_gate(int b, int c, int hidden, int tiles, float count, long long stride_b, long long stride_t,
                long long stride_c, const float *sums, const float *w1, const float *w2, int use_relu,
                float *gate, bdm_stream_t stream);
int bdm_groupnorm_cl_supported(int c, int groups);
size_t bdm_groupnorm_cl_workspace_bytes(int b, int c, long long s);
int bdm_groupnorm_cl_tiles(int b, int c, long long s, int groups);
int bdm_groupnorm_act_cl(int b, int c, long long s, int groups, float eps, int swish, const float *x,
                         const float *conv_bias, const float *gamma, const float *beta, float *y,
                         float *tile_sums, void *workspace, size_t workspace_bytes, int precomputed_chunks,
                         bdm_stream_t stream);
/* precomputed_chunks > 0: workspace holds the producer's statistics f64[b,precomputed_chunks,c,2] (see
 * bdm_sparse_conv3_gather) and the statistics pass over x is skipped.
 * precomputed_chunks < 0: workspace holds GROUP-level producer statistics f64[b,-precomputed_chunks,groups,2] -- (sum, sum
 * of squares) per normalisation group over disjoint blocks of voxels, of the tensor as it is (conv_bias must be NULL):
 * what bdm_conv3_tc05 writes per unit.  The same convention holds for `chunks` of bdm_groupnorm_cl_sums and
 * bdm_groupnorm_swish_half_planar. */
/* Statistics-only half for a consumer that normalises on the fly (bdm_trilinear_devoxelize_cl_norm): x + producer
 * statistics partials f64[b,chunks,c,2] -> tile_sums f32[b,tiles,c] (tiles = bdm_groupnorm_cl_sums_tiles; sums of
 * y = act(group_norm(x + conv_bias)), what bdm_se_gate takes) and coef f32[b,c,2] = (A, B) with y = act(x*A + B).
 * One read of x, nothing else written. */
int bdm_groupnorm_cl_sums_tiles(int b, int c, long long s);
int bdm_groupnorm_cl_sums(int b, int c, long long s, int groups, float eps, int swish, const float *x,
                          const float *conv_bias, const float *gamma, const float *beta, const double *partials,
                          int chunks, float *tile_sums, float *coef, bdm_stream_t stream);

/* ---- dense side: the 3x3x3 convolution of the voxel branch on the 5th-generation tensor cores -------------------
 * replaces the second nn.Conv3d(c, c, 3, padding=1) of a PVConv block's voxel_layers (modules/pvconv.py:75-88), which
 * torch hands to cuDNN (TF32 under the default conv policy), together with the GroupNorm+Swish pass in front of it
 * (which now writes the convolution's fp16 operand directly) and the statistics pass of the GroupNorm behind it.
 * Precision: operands rounded to fp16 after a power-of-two scaling (11 significant bits, as TF32 keeps), exact
 * products, fp32 accumulation in tensor memory.
 *   activations xh: fp16 "chunk planes" [c_in/8][plane_rows][8] of the flat padded grid (position of voxel (x,y,z) of
 *       sample i: guard + i*sample_rows + (x*(r+1)+y)*(r+1)+z, see csrc/conv3_tc05.cu); the buffer must be zero-filled
 *       once by the caller (pad positions are never written), plane_rows = bdm_conv3_tc05_plane_rows(b, r).
 *   bdm_conv3_tc05_prepare   weight f32[c_out][c_in][3][3][3] -> `prepared` (bdm_conv3_tc05_weight_bytes, 256-byte
 *       aligned): scales + the per-stage shared-memory images of the weights.  gamma / beta f32[c_in] (or NULL) and
 *       group_elems (elements of one normalisation group) are those of the GroupNorm that produces xh: they bound
 *       the activations for the fp16 scaling.  Once per weight version.
 *   bdm_groupnorm_swish_half_planar   x f32[b][r^3][c] channels-last + its producer's statistics partials
 *       f64[b][chunks][c][2] (of the bias-less tensor) -> act(group_norm(x + conv_bias)) * act_scale as xh.
 *   bdm_conv3_tc05_fill_planes   the FIRST Conv3d's operand: per-occupied-voxel averages (bdm_avg_voxelize_compact)
 *       + the voxel plan -> xh (zeros at empty voxels), scaled by a power of two derived from max|average| on
 *       the device and recorded in `prepared` (prepare that convolution with gamma = beta = NULL, group_elems = 1);
 *       amax_ready != 0: bdm_avg_voxelize_compact_amax already wrote max|average| to word 4 of `prepared`.
 *       occ (or NULL): u32[b][bdm_conv3_tc05_occ_words(r)], one bit per non-zero row; passed on to bdm_conv3_tc05 it
 *       makes the convolution skip (loads and MMAs) the tap windows that hold only zero rows -- exact, since those
 *       products are zeros (NULL there = dense operand).
 *   bdm_conv3_tc05   out f32[b][r^3][c_out] = conv(xh) + bias; stats (or NULL): the result's GroupNorm(8) statistics.
 *       With workspace (bdm_conv3_tc05_workspace_bytes(b, r) bytes): f64[b][1][c_out][2] in the layout
 *       bdm_groupnorm_act_cl(precomputed_chunks = 1) takes (one more small kernel folds the units).  With workspace
 *       NULL: the per-unit group partials themselves, f64[b][bdm_conv3_tc05_units(c_in, c_out, r)][8][2], for
 *       precomputed_chunks = -bdm_conv3_tc05_units(c_in, c_out, r) (no extra kernel).
 * c_out in {32, 64, 128}, c_in = 32 or a multiple of 64, r a power of two (bdm_conv3_tc05_supported). */
int bdm_conv3_tc05_supported(int c_in, int c_out, int r);
long long bdm_conv3_tc05_plane_rows(int b, int r);
int bdm_conv3_tc05_units(int c_in, int c_out, int r);
size_t bdm_conv3_tc05_weight_bytes(int c_in, int c_out);
size_t bdm_conv3_tc05_workspace_bytes(int b, int r);
int bdm_conv3_tc05_prepare(int c_in, int c_out, const float *weight, const float *gamma, const float *beta,
                           long long group_elems, void *prepared, size_t prepared_bytes, bdm_stream_t stream);
int bdm_groupnorm_swish_half_planar(int b, int c, int r, int groups, float eps, int swish, const float *x,
                                    const float *conv_bias, const float *gamma, const float *beta,
                                    const double *partials, int chunks, const void *prepared, void *xh,
                                    long long plane_rows, bdm_stream_t stream);
int bdm_conv3_tc05_fill_planes(int b, int c, int n, int r, const float *compact, const void *plan_workspace,
                               size_t plan_workspace_bytes, void *prepared, void *xh, long long plane_rows,
                               int amax_ready, unsigned *occ, bdm_stream_t stream);
int bdm_conv3_tc05_occ_words(int r);
int bdm_conv3_tc05(int b, int c_in, int c_out, int r, const void *xh, long long plane_rows, const void *prepared,
                   const float *bias, float *out, double *stats, void *workspace, size_t workspace_bytes,
                   const unsigned *occ, bdm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* BDM_B200_H */
