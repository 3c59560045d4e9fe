/*
 * cmr_b200.h - C ABI of libcmr_b200.so, the sm_100a implementation of CMR-Agent's
 * per-iteration geometric hot path (environment step + PointNet++ front-end).
 *
 * The reference (y2w-oc/CMR-Agent) is pure Python; its "FFI" for this path is the set of
 * module-level functions in environment/environment.py and models/pointnet_util.py.  Each entry
 * point below names the reference lines it replaces (paths relative to the reference root).
 * The Python drop-ins in cmr_agent_b200/{environment,pointnet_util}.py bind these through ctypes
 * and keep the reference's signatures; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless marked "host"; tensors are dense, row-major,
 *     in exactly the layout the reference's tensors have (shapes in brackets);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous and never
 *     synchronises the host;
 *   - return value: 0 = ok; CMR_E* (< 0) = bad argument, nothing was launched;
 *     > 0 = the cudaError_t of a failed launch.  cmr_error_string() explains any of them;
 *   - there is no CPU fallback: without a CUDA device every call returns a cudaError_t.
 */
#ifndef CMR_B200_H_
#define CMR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CMR_API __attribute__((visibility("default")))
#else
#define CMR_API
#endif

#define CMR_ABI_VERSION 5

#define CMR_OK 0
#define CMR_EINVAL (-1)      /* null pointer, non-positive size */
#define CMR_EALIGN (-2)      /* pointer not aligned as documented */
#define CMR_ERANGE (-3)      /* size outside what the kernels support */
#define CMR_EUNSUPPORTED (-4)

#define CMR_REWARD_SHIPPED 0  /* environment.py:272-290 as shipped: pose is ignored */
#define CMR_REWARD_INTENDED 1 /* applies the disentangled transform of the commented line :273 */

CMR_API int cmr_abi_version(void);
CMR_API const char *cmr_error_string(int code);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
CMR_API unsigned long long cmr_launch_count(void);

/* ------------------------------------------------------------------ environment ---- */

/* Bytes of scratch cmr_episode_prepare/cmr_observe need for (B, N, C, H*W). */
CMR_API size_t cmr_workspace_bytes(int B, int N, int C, int P);

/* pc.mean(dim=2) - environment.py:46,91,274.  pc [B,3,N] f32 -> mean [B,3] f32.
 * Deterministic, accumulated in fp64 (correctly rounded in practice).  NOTE torch's own fp32 mean
 * is not correctly rounded; a host that needs the reference's bits on a given device passes
 * torch's mean to the calls below instead (the Python drop-in does). */
CMR_API int cmr_cloud_mean(const float *pc, int B, int N, float *mean, void *stream);

/* Once per episode batch: the boolean-mask compaction of environment.py:48-49, hoisted out of the
 * iteration loop.  overlap [B,N] u8 (torch.bool), feat [B,C,N] f32 channel-major.
 * Writes into `workspace` (cmr_workspace_bytes): per-128-point prefix counts, the number M[b] of
 * predicted-overlap points and their features transposed to point-major rows [B, M, C].
 * C must be a multiple of 4 and <= 256. */
CMR_API int cmr_episode_prepare(const uint8_t *overlap, const float *feat, int B, int N, int C, void *workspace,
                        void *stream);
/* The two halves of cmr_episode_prepare, for hosts with a second stream: cmr_project needs only the scan (prefix counts
 * + cleared bucket counters); the compaction - which reads all of feat - is needed from the first cmr_tile_scatter on
 * and may run beside the first projection.  cmr_episode_compact must be ordered after cmr_episode_scan. */
CMR_API int cmr_episode_scan(const uint8_t *overlap, int B, int N, int C, void *workspace, void *stream);
CMR_API int cmr_episode_compact(const uint8_t *overlap, const float *feat, int B, int N, int C, void *workspace,
                                void *stream);

/* observation_from_a_pose - environment.py:25-126.
 *   pc [B,3,N] f32, overlap [B,N] u8, img_feat [B,C,H,W] f32, K [B,3,3] f32, pose [B,4,4] f32,
 *   mean [B,3] f32 (the cloud mean, see cmr_cloud_mean), workspace from cmr_episode_prepare.
 *   obs2d [B,2C,H,W] f32 = cat(img_feat, scatter-mean of the predicted-overlap points' features at
 *   their round-half-even pixel) ; obs3d [B,5,N] f32 = cat(pc, overlap, in_frustum).
 *   pix_out (optional, may be NULL) [B,N] i32: v*W+u of every point, H*W when out of frustum -
 *   the integer by-product of :67-72, exported for parity checks.
 *   mvis_out (optional) [B] i32: predicted-overlap points that landed inside the frustum.
 * Sums run in point order per pixel (deterministic, no floating-point atomics).  The workspace's pixel-id
 * scratch and bucket buffers are rewritten by every call: do not run two observes on one workspace
 * concurrently. */
CMR_API int cmr_observe(const float *pc, const uint8_t *overlap, const float *img_feat, const float *K,
                const float *pose, const float *mean, void *workspace, int B, int N, int C, int H,
                int W, float *obs2d, float *obs3d, int32_t *pix_out, int32_t *mvis_out, void *stream);

/* The two halves of cmr_observe, for callers that need only one observation or time them apart:
 *   cmr_project      environment.py:88-124 (obs3d) + :54-72 for the predicted-overlap points (pixel ids
 *                    into the workspace).  When img_feat and obs2d are given it also carries the image
 *                    half of obs2d (obs2d[:, 0:C] = img_feat, :83) as tiled TMA traffic, if their layout
 *                    allows (16-byte aligned, H*W % 4 == 0, H*W >= 128); *image_copied (host, optional)
 *                    reports whether it did.
 *                    It also APPENDS the visible predicted-overlap points to the workspace's per-bucket
 *                    buffers, which the next cmr_tile_scatter consumes and clears.  By default the bucket
 *                    counters are cleared first (one memset), so the call may be repeated freely; pass
 *                    CMR_PROJECT_PAIRED in `flags` when every cmr_project on this workspace is followed by
 *                    exactly one cmr_tile_scatter - what cmr_observe does - to skip that memset.
 *   cmr_tile_scatter environment.py:74-86: the projected half obs2d[:, C:2C] from what the last
 *                    cmr_project left (once per cmr_project); with copy_image != 0 it also copies the
 *                    image half itself. */
#define CMR_PROJECT_PAIRED 1
CMR_API int cmr_project(const float *pc, const uint8_t *overlap, const float *K, const float *pose, const float *mean,
                        void *workspace, int B, int N, int C, int H, int W, float *obs3d, int32_t *pix_out,
                        int32_t *mvis_out, const float *img_feat, float *obs2d, int *image_copied, int flags,
                        void *stream);
CMR_API int cmr_tile_scatter(const float *img_feat, const float *K, void *workspace, int B, int N, int C, int H,
                             int W, int copy_image, float *obs2d, void *stream);

/* to_disentangled - environment.py:15-21.  poses [B,4,4] in place: t <- (t - m) + R m. */
CMR_API int cmr_to_disentangled(float *poses, const float *mean, int B, void *stream);

/* step - environment.py:179-207, in place on pose [B,4,4].
 *   action_r [B,1] (3-DoF) or [B,3] (6-DoF) i64, action_t [B,2] or [B,3] i64;
 *   rot_tab [3,nbins,3,3] f32: per-axis rotation matrices Rx,Ry,Rz(r_steps[i]) exactly as
 *   environment.py:235-260 builds them; t_tab [nbins] f32 = float32(t_steps).
 *   R <- ((Rx @ Ry) @ Rz) @ R with every 3x3 product an FMA chain; t <- t + move_t.
 * Out-of-range actions leave that pose untouched and raise the sticky flag of cmr_take_fault. */
CMR_API int cmr_step(float *pose, const int64_t *action_r, const int64_t *action_t, const float *rot_tab,
             const float *t_tab, int nbins, int dof6, int B, void *stream);

/* reward - environment.py:263-302.
 *   target = pc_in_cam_space [B,3,N] f32, mask = pc_mask [B,N] u8, prev (may be NULL) [B] f32.
 *   dist [B] f32 = mean over masked points of |target - moved|^2 (NaN when the mask is empty);
 *   reward [B] f32 = +0.5 / -0.5 / 0 against prev (zeros when prev is NULL).
 *   scratch: >= cmr_reward_scratch_bytes(B) bytes. */
CMR_API size_t cmr_reward_scratch_bytes(int B);
CMR_API int cmr_reward(const float *target, const float *pc, const uint8_t *mask, const float *mean, const float *pose,
               const float *prev, int mode, int B, int N, void *scratch, float *reward, float *dist,
               void *stream);

/* expert - environment.py:143-176 on the device (no D2H / scipy / H2D round trip per step).
 *   delta_R in fp32, then fp64: scipy's Rotation.from_matrix -> as_euler('xyz') algorithm, the reference's
 *   ">3 rad" fix-ups and a first-minimum argmin over r_steps / t_steps [nbins] f64 (device).
 *   action_r [B,1] (3-DoF) or [B,3], action_t [B,2] or [B,3], i64. */
CMR_API int cmr_expert(const float *pose_source, const float *pose_target, const double *r_steps, const double *t_steps,
                       int nbins, int dof6, int B, int64_t *action_r, int64_t *action_t, void *stream);

/* The shipped reward (environment.py:272-290) ignores its pose: its distance is a constant of the episode batch.
 * A host that has kept the distance of an earlier cmr_reward call on the same batch (dist_cached [B]) gets the
 * comparison with `prev` (:293-298) and fresh copies from this call instead of re-reading 25 bytes per point. */
CMR_API int cmr_reward_compare(const float *dist_cached, const float *prev, int B, float *reward, float *dist,
                               void *stream);

/* One agent iteration as ONE host call: the loop body of Train_Agent.py:229-248 / Test_Agent.py:154-170 around
 * the agent's forward pass -  step (environment.py:179-207), reward of the new pose (:263-302), observation of the
 * new pose (:25-126).  `a` holds everything that is constant over the iterations of an episode batch and is filled
 * once per batch; the per-iteration arguments choose the parts:
 *   action_r/action_t NULL -> no step;  reward/dist NULL -> no reward;  obs2d/obs3d NULL -> no observation.
 * With reward_mode == CMR_REWARD_SHIPPED and dist_cached != NULL the reward is cmr_reward_compare. */
typedef struct cmr_iteration_args {
    const float *pc;          /* [B,3,N] */
    const uint8_t *overlap;   /* [B,N] */
    const float *img_feat;    /* [B,C,H,W] */
    const float *K;           /* [B,3,3] */
    const float *mean;        /* [B,3] */
    void *workspace;          /* cmr_episode_prepare */
    const float *rot_tab;     /* step tables, see cmr_step */
    const float *t_tab;
    const float *target;      /* reward: pc_in_cam_space [B,3,N] */
    const uint8_t *mask;      /* reward: pc_mask != 0, [B,N] */
    void *reward_scratch;     /* cmr_reward_scratch_bytes(B), zeroed once */
    const float *dist_cached; /* [B] or NULL */
    int B, N, C, H, W, nbins, dof6, reward_mode;
} cmr_iteration_args;
CMR_API int cmr_iteration(const cmr_iteration_args *a, float *pose, const int64_t *action_r, const int64_t *action_t,
                          const float *prev, float *reward, float *dist, float *obs2d, float *obs3d, void *stream);

/* ------------------------------------------------------------------ rollout session ---- */

/* A whole registration rollout - init (environment.py:129-140), to_disentangled of the target (:15-21), then
 * `iters` times  observation_from_a_pose -> step -> reward  (Test_Agent.py:154-170 / Train_Agent.py:229-248 with the
 * actions given) - driven from HOST buffers by one call, pipelined: the session owns `depth` slots of device memory,
 * a copy stream and a compute stream, and uploads rollout k+1 while the kernels of rollout k run.
 *   cmr_session_create   allocates everything (device buffers, pinned result staging, streams, events) for batches
 *                        of the configured shape.  features_resident != 0: `feat` and `img_feat` of cmr_rollout_inputs
 *                        are DEVICE pointers (where CMR-Agent's feature network leaves them) and are not copied.
 *   cmr_session_submit   enqueues the uploads and the kernels of one rollout and returns immediately; host buffers
 *                        must stay valid (and should be pinned) until the matching cmr_session_wait.
 *                        pc_mask is int64 as the dataset delivers it; actions are [iters, B, 1|3] / [iters, B, 2|3] int64.
 *   cmr_session_wait     blocks until rollout `ticket` is complete and copies its results to the caller:
 *                        rewards / dists [iters, B], final poses [B,4,4], disentangled target poses [B,4,4]
 *                        (any of them may be NULL).
 *   cmr_session_stats    average host->device rate of the uploads so far (GB/s, CUDA events on the copy stream)
 *                        and the bytes one rollout uploads. */
typedef struct cmr_session cmr_session;
typedef struct cmr_session_config {
    int B, N, C, H, W, iters, dof6, reward_mode, depth, features_resident, nbins;
    const float *rot_tab; /* host: step tables, see cmr_step ([3, nbins+1, 3, 3]) */
    const float *t_tab;   /* host: [nbins] */
} cmr_session_config;
typedef struct cmr_rollout_inputs {
    const float *pc;            /* host [B,3,N] */
    const uint8_t *overlap;     /* host [B,N] */
    const float *feat;          /* host (or device, features_resident) [B,C,N] */
    const float *img_feat;      /* host (or device, features_resident) [B,C,H,W] */
    const float *K;             /* host [B,3,3] */
    const float *P;             /* host [B,4,4]: data['P'] */
    const float *pc_in_cam;     /* host [B,3,N] */
    const int64_t *pc_mask;     /* host [B,N] */
    const int64_t *action_r;    /* host [iters,B,1|3] */
    const int64_t *action_t;    /* host [iters,B,2|3] */
} cmr_rollout_inputs;
CMR_API int cmr_session_create(const cmr_session_config *cfg, cmr_session **out);
CMR_API void cmr_session_destroy(cmr_session *s);
CMR_API int cmr_session_submit(cmr_session *s, const cmr_rollout_inputs *in, long long *ticket);
CMR_API int cmr_session_wait(cmr_session *s, long long ticket, float *rewards, float *dists, float *poses,
                             float *target_poses);
CMR_API int cmr_session_stats(const cmr_session *s, double *h2d_gbs, double *bytes_per_rollout);
/* the observation of the LAST iteration of rollout `ticket` (still in its slot), copied device-to-device into the
 * caller's obs2d [B,2C,H,W] / obs3d [B,5,N] on `stream` (ordered after the rollout) */
CMR_API int cmr_session_last_observation(cmr_session *s, long long ticket, float *obs2d, float *obs3d, void *stream);

/* ------------------------------------------------------------------ pointnet_util ---- */

/* square_distance - pointnet_util.py:19-33. src [B,S,3], dst [B,N,3] (any strides, in floats) ->
 * out [B,S,N] f32, (dx*dx + dy*dy) + dz*dz unfused. */
CMR_API int cmr_square_distance(const float *src, const int64_t src_stride[3], const float *dst,
                        const int64_t dst_stride[3], int B, int S, int N, float *out, void *stream);

/* index_points - pointnet_util.py:36-47.  points [B,N,C] rows of `row_bytes` bytes, idx [B,S] i64
 * (flatten [B,S,K] to [B,S*K]) -> out [B,S,row_bytes].  Out-of-range indices write zeros and raise
 * the sticky flag of cmr_take_fault. */
CMR_API int cmr_index_points(const void *points, const int64_t *idx, int B, int N, int S, int row_bytes, void *out,
                     void *stream);
/* backward of index_points for f32: grad_points [B,N,C] += grad_out rows (grad_points pre-zeroed). */
CMR_API int cmr_index_points_backward(const float *grad_out, const int64_t *idx, int B, int N, int S, int C,
                              float *grad_points, void *stream);

/* farthest_point_sample - pointnet_util.py:50-70.  xyz [B,N,3] f32, start [B] i64 (the draw of
 * :62, made by the host on the CPU generator) -> out [B,npoint] i64.  Lowest index wins ties. */
CMR_API int cmr_farthest_point_sample(const float *xyz, const int64_t *start, int B, int N, int npoint, int64_t *out,
                              void *stream);

/* The same sampling with cell pruning on the uniform grid of cmr_knn_grid (workspace of the same size): per round
 * only the cells that the newest centroid can still affect are updated, one SM carries a whole cloud.  Indices
 * identical to cmr_farthest_point_sample.  N up to about 41900 (the running distances live in shared memory;
 * CMR_ERANGE beyond). */
CMR_API int cmr_farthest_point_sample_grid(const float *xyz, const int64_t *start, int B, int N, int npoint, void *workspace,
                                           int64_t *out, void *stream);

/* kNN = square_distance(...).argsort()[:, :, :k] - pointnet_util.py:115-116,233-234,
 * models/PointNN.py:215-216, with the stable order (distance, index).  out [B,S,k] i64. k <= 128. */
CMR_API int cmr_knn(const float *query, const float *ref, int B, int S, int N, int k, int64_t *out, void *stream);

/* The same kNN, result identical bit for bit, with a uniform grid over the reference cloud built on the fly
 * (counting sort by cell, then one warp per query walks rings of cells until no unexplored point can enter the
 * list): two orders of magnitude fewer distance evaluations on LiDAR-scale clouds.
 * workspace: the size cmr_knn_grid_workspace_bytes reports for (B, N), 256-byte aligned. */
CMR_API size_t cmr_knn_grid_workspace_bytes(int B, int N);
CMR_API int cmr_knn_grid(const float *query, const float *ref, int B, int S, int N, int k, void *workspace, int64_t *out,
                         void *stream);

/* query_ball_point - pointnet_util.py:73-93.  radius2 = float32(radius**2). out [B,S,nsample] i64. */
CMR_API int cmr_query_ball_point(const float *query, const float *ref, float radius2, int nsample, int B, int S, int N,
                         int64_t *out, void *stream);

/* The same ball query on the uniform grid of cmr_knn_grid (workspace of the same size): only the cells the ball can
 * touch are read.  Result identical to cmr_query_ball_point.  radius = the reference's python float as f32,
 * radius2 = float32(radius**2) as above.  nsample <= 128. */
CMR_API int cmr_query_ball_point_grid(const float *query, const float *ref, float radius2, float radius, int nsample, int B,
                                      int S, int N, void *workspace, int64_t *out, void *stream);

/* grouping tail of sample_and_group - pointnet_util.py:120-129 fused: out [B,S,K,3+D] f32 =
 * cat(xyz[idx] - new_xyz, points[idx]); points may be NULL (D = 0). */
CMR_API int cmr_group_points(const float *xyz, const float *points, const float *new_xyz, const int64_t *idx, int B,
                     int N, int S, int K, int D, float *out, void *stream);

/* Sticky device-side fault flag (out-of-range index / action).  Reads and clears it; this call
 * synchronises `stream`.  0 = none. */
CMR_API int cmr_take_fault(void *stream);

/* ------------------------------------------------------------------ cost volume ---- */

/* The pose-sampling cost volume of models/IterModel.py:272-351: the same projection + ordered scatter as the
 * observation, for K candidate poses per cloud, without the disentangling (X = R p + t) and for all N columns
 * at once (the reference projects before it masks).
 *   cmr_cost_volume_prepare  once per batch of clouds: compacts the rows of the masked points.
 *                            mask [B,N] u8 (IterModel uses pc_overlap_pred[0] for every cloud), feat [B,C,N] f32.
 *   cmr_cost_volume_warp     pc [B,3,N], Kmat [B,3,3], poses [B*K,4,4] f32 (rows 0..2 = the [3,4] of :276-279)
 *                            -> out [B*K, C, H*W] f32: channel c < mean_channels = scatter-mean of feat[c] over the
 *                            masked points that land in the pixel (:341), channel c >= mean_channels = their SUM
 *                            (:343: pass the in-camera scores as an extra channel).  mean_channels is a multiple of
 *                            64 or >= C.  Sums run in point order.  B*K <= 65535, H*W <= 12288.
 *                            The reference's shape (C = 68: 64 features + score + padding, mean_channels = 64, H*W a
 *                            multiple of 32 up to 8192, N <= 65535) takes a dedicated pair of kernels (one counting
 *                            sort per pose, csrc/cost_volume_kernels.cuh); every other shape takes the observation's
 *                            bucket kernels.  Same bits either way (CMR_B200_CV=buckets forces the second path). */
CMR_API size_t cmr_cost_volume_workspace_bytes(int B, int K, int N, int C, int P);
CMR_API int cmr_cost_volume_prepare(const uint8_t *mask, const float *feat, int B, int K, int N, int C,
                                    void *workspace, void *stream);
CMR_API int cmr_cost_volume_warp(const float *pc, const uint8_t *mask, const float *Kmat, const float *poses,
                                 void *workspace, int B, int K, int N, int C, int H, int W, int mean_channels,
                                 float *out, void *stream);

/* ------------------------------------------------------------------ bilinear sampling ---- */

/* Image features sampled bilinearly at the points' projections ("bilinearly sample image features onto visible points",
 * BASELINE.json north_star).  An EXTRA operator: the reference has no point-side gather - its
 * observation is the reverse scatter of environment/environment.py:67-83, cmr_observe above - so this one is
 * specified by F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=True) evaluated at the pixel
 * coordinates of environment.py:54-59, times the frustum mask of :61-65 (oracle/sample_oracle.py).
 *   cmr_sample_prepare         once per batch of images: img_feat [B,C,P] f32 -> workspace ([B,P,C], a pixel = one row)
 *   cmr_sample_image_features  pc [B,3,N], Kmat [B,3,3], pose [B,4,4] (16-byte aligned), mean [B,3] (the clouds' means,
 *                              cmr_episode_prepare) -> out [B,C,N] f32 (zero where the point is outside the frustum),
 *                              in_cam [B,N] u8 (may be NULL).  C even.  The four weighted neighbours are summed left to
 *                              right, every operation rounded: bit-identical to the CPU restatement. */
CMR_API size_t cmr_sample_workspace_bytes(int B, int C, int P);
CMR_API int cmr_sample_prepare(const float *img_feat, int B, int C, int P, void *workspace, void *stream);
CMR_API int cmr_sample_image_features(const float *pc, const float *Kmat, const float *pose, const float *mean,
                                      const void *workspace, int B, int N, int C, int H, int W, float *out,
                                      uint8_t *in_cam, void *stream);

/* ------------------------------------------------------------------ agent: heads ---- */

/* The actor-critic heads of the agent - models/CMRAgent.py:70-86 (policy_r, policy_t, value: Linear - LeakyReLU -
 * Linear - LeakyReLU - Linear each) as applied at :106-113.  One call = layer l of ALL heads: a GROUPED linear layer
 *     out[b, n] = act(bias[n] + sum_k in[b, in_off_g + k] * W[w_off_g + (n - n0_g) * K_g + k]),  n in [n0_g, n1_g)
 * desc (HOST): `groups` x {in_off, K, n0, n1, w_off} int64; the groups tile [0, N) in order; K <= 256.
 * in [B, in_stride], out [B, out_stride] f32 on the device; activate != 0 applies LeakyReLU(negative_slope).
 * fp32, explicit fused multiply-adds in a fixed order: deterministic, a row's result does not depend on B. */
CMR_API int cmr_grouped_linear(const float *in, int in_stride, const float *W, const float *bias, const int64_t *desc,
                               int groups, int B, int N, float negative_slope, int activate, float *out, int out_stride,
                               void *stream);

/* Epilogue of one convolution layer of the agent's 2-D head - models/CMRAgent.py:34-61 (state_2d_embed: Conv2d 3x3,
 * [BatchNorm2d], LeakyReLU, [AvgPool2d]), EVAL MODE: the caller runs the convolution WITHOUT its bias and folds bias
 * and BatchNorm2d into per-channel scale / shift (scale = g / sqrt(var + eps), shift = (bias - mean) * scale + beta;
 * 1 and bias without a BatchNorm).  One pass instead of up to four elementwise launches:
 *     y = pool(LeakyReLU(x * scale[c] + shift[c]))
 * x [B,C,H,W] f32, 16-byte aligned; channels_last = 0: NCHW contiguous, 1: torch's channels_last ([B][H][W][C] in
 * memory, C % 4 == 0) - cuDNN's tensor-core convolutions are NHWC kernels and convert every layer's input and output when
 * handed NCHW (half of their time); y has the layout of x.  pool 0: none, y [B,C,H,W] (may be x itself; NCHW:
 * H*W % 4 == 0);  1: AvgPool2d(2,2), y [B,C,H/2,W/2] (H, W even; NCHW: W % 4 == 0);  2: AvgPool2d((H,W)), y [B,C,1,1]. */
CMR_API int cmr_conv_epilogue(const float *x, const float *scale, const float *shift, float negative_slope, int pool,
                              int channels_last, int B, int C, int H, int W, float *y, void *stream);

/* The deterministic action - models/CMRAgent.py:118-124 (`action_from_logits(..., deterministic=True)`:
 * `argmax(Categorical(logits=x).probs, dim=-1)` for the rotation and the translation logits), Test_Agent.py:166.
 * One launch instead of torch's 22.  r_logits [B,degree_r,steps], t_logits [B,degree_t,steps] f32, innermost two
 * dimensions dense, *_batch_stride in elements (the heads' logits are column blocks of one wider matrix);
 * action_r [B,degree_r], action_t [B,degree_t] int64; probs_r / probs_t: NULL, or [B,degree,steps] f32 receiving
 * the probabilities (tests).  9 <= steps <= 16 (the reference: 11): within that range the kernel reproduces torch's
 * logsumexp and softmax operation by operation, INCLUDING the order of their sums, so that the probabilities are
 * torch's bit for bit and ties between rounded probabilities resolve to the same (first) index; other step counts
 * return CMR_EUNSUPPORTED and the caller keeps the reference's own function. */
CMR_API int cmr_deterministic_action(const float *r_logits, int degree_r, int64_t r_batch_stride, const float *t_logits,
                                     int degree_t, int64_t t_batch_stride, int B, int steps, int64_t *action_r,
                                     int64_t *action_t, float *probs_r, float *probs_t, void *stream);

/* The observation handed to the agent's 2-D head in the layout its convolutions run in - models/CMRAgent.py:89
 * (`self.state_2d_embed(state_2d)`: the first Conv2d reads obs2d [B,2C,H/4,W/4] as environment.py:126 returns it, NCHW):
 * x [B,C,H,W] f32 contiguous -> y the same values as torch's channels_last ([B][H][W][C] in memory), a pure copy
 * (what `x.contiguous(memory_format=torch.channels_last)` returns, 2.3x faster at the reference's shape).  x != y. */
CMR_API int cmr_to_channels_last(const float *x, int B, int C, int H, int W, float *y, void *stream);

/* ------------------------------------------------------------------ agent: 3-D tower ---- */

/* The 3-D state tower of the agent - models/CMRAgent.py:25-29 (state_3d_embed: four ConvBNReLURes1D blocks,
 * models/PointNN.py:260-282) and its forward loop :92-101 (block, global max over the points, repeat + cat),
 * EVAL MODE: the caller folds every BatchNorm1d into the 1x1 convolution before it (W' = W g / sqrt(var + eps),
 * b' = (b - mean) g / sqrt(var + eps) + beta) and hands over plain fp32 matrices.  tcgen05 tensor cores, fp32
 * accumulation in TMEM, three bf16 passes over split operands per product (1e-5 on the output's scale).
 *   cmr_tower_pack     once per set of weights and block: folded weights -> device blob (cmr_tower_blob_bytes(kind)).
 *                      CMR_TOWER_FIRST: W1 [5,5] b1 [5] W2 [64,5] b2 [64] Ws [64,5] bs [64]   (net.0/1, net.3/4, shortcut)
 *                      CMR_TOWER_MID:   W1 [128,128] b1 [128] W2 [64,128] b2 [64] Ws [64,128] bs [64]
 *                      CMR_TOWER_LAST:  W1 [128,128] b1 [128] W2 [128,128] b2 [128], Ws = bs = NULL (identity shortcut)
 *                      all row-major [out, in] f32 on the device; blob 128-byte aligned.
 *   cmr_tower_forward  obs3d [B,5,N] f32 (observation_3d of cmr_observe) -> embed [B,128] f32 = embed_3d of
 *                      CMRAgent.py:101.  workspace: cmr_tower_workspace_bytes(B, N), 1024-byte aligned. */
enum { CMR_TOWER_FIRST = 0, CMR_TOWER_MID = 1, CMR_TOWER_LAST = 2 };
CMR_API size_t cmr_tower_blob_bytes(int kind);
CMR_API int cmr_tower_pack(int kind, const float *W1, const float *b1, const float *W2, const float *b2, const float *Ws,
                           const float *bs, void *blob, void *stream);
CMR_API size_t cmr_tower_workspace_bytes(int B, int N);
CMR_API int cmr_tower_forward(const float *obs3d, const void *blob1, const void *blob2, const void *blob3,
                              const void *blob4, void *workspace, int B, int N, float *embed, void *stream);

/* ------------------------------------------------------------------ dataset side ---- */

/* FarthestSampler.sample - dataset/KittiDataset.py:107-126 (= dataset/NuScenesDataset.py:25-44), float64 like
 * the numpy original.  pts [B,3,M] f64 channel-major (numpy's [3, M] per sample), start [B] i64 (init_idx of
 * :118, drawn by the host) -> out_idx [B,k] i64 and, if out_pts is not NULL, out_pts [B,3,k] f64.
 * np.argmax semantics: the first maximum wins.  M <= 16384, k <= M. */
CMR_API int cmr_fps_f64(const double *pts, const int64_t *start, int B, int M, int k, int64_t *out_idx,
                        double *out_pts, void *stream);

/* cKDTree(ref.T).query(query.T, k=1)[1] - dataset/KittiDataset.py:365-366: the nearest reference point (node) of
 * every query point, exact, float64.  query [B,3,N], ref [B,3,S] f64 channel-major -> out [B,N] i64.
 * Lowest index on exact ties (scipy leaves ties unspecified). */
CMR_API int cmr_nearest_f64(const double *query, const double *ref, int B, int N, int S, int64_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CMR_B200_H_ */
