/*
 * sage_ba.h -- C ABI of the B200-native dense bundle-adjustment backend for SAGE-SLAM.
 *
 * This is the drop-in boundary for the reference's `df_cuda` static library
 * (/root/reference/system/sources/cuda/CMakeLists.txt:37-46): every `df::*_calculate` free
 * function the factor classes (core/gtsam/*_factor.cpp) and the camera tracker
 * (core/system/camera_tracker.cpp) call has one entry point below, with plain pointers and
 * sizes only -- no torch / ATen / Eigen types.  INTEGRATION.md shows the `df::` shim a
 * maintainer links in place of df_cuda.
 *
 * Conventions
 *   - All arithmetic is fp32 (the reference hard-codes <float>, photometric_factor_kernels.cpp:1111).
 *     J^T J | J^T r are reduced on the tensor cores with a 3xTF32 split and bounded accumulation chains: against sums of
 *     the same fp32 rows in fp64 every block agrees to about 1e-5 relative (Jacobi-scaled), independent of the sample count
 *     and of how a factor is split over CTAs; the reference (materialised J, cuBLAS sgemm) sits at about 1e-6.
 *   - Rotations are row-major float[9], translations float[3]; poses are keyframe->world
 *     (pose_wk); T10 = T1^-1 T0 (gtsam/photometric_factor.cpp:280-281).
 *   - AtA is row-major [D,D], Atb is [D]; variable order inside a factor is the reference's:
 *       photometric / reprojection: [pose0(v,w) pose1(v,w) code0(C) scale0]          D = 13+C
 *       geometric:                  [pose0 pose1 code0(C) code1(C) scale0 scale1]    D = 14+2C
 *       tracker:                    [rel. pose (v,w)] (+ scale0)                     D = 6 / 7
 *   - Every function returns 0 on success, non-zero on failure; sage_ba_last_error() gives
 *     the message.  (The reference prints and exit()s on CUDA errors,
 *     photometric_factor_kernels.cpp:23-31; the df:: shim maps non-zero to the same.)
 *   - Calls on one context are serialised on that context's stream; use one context per
 *     host thread (the reference is called from up to 4 threads, deepfactors.cpp:1497-1505).
 *   - Pointers are HOST pointers unless the parameter is documented as device memory.
 */
#ifndef SAGE_BA_H_
#define SAGE_BA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAGE_BA_MAX_LEVELS 8
#define SAGE_BA_MAX_CODE 32

enum { SAGE_BA_HOST = 0, SAGE_BA_DEVICE = 1 };

/* df::PinholeCamera<float> (common/pinhole_camera.h:124-130): width/height stored as float. */
typedef struct sage_ba_camera
{
  float fx, fy, u0, v0, width, height;
} sage_ba_camera;

typedef struct sage_ba_context sage_ba_context;
typedef struct sage_ba_keyframe sage_ba_keyframe;
typedef struct sage_ba_problem sage_ba_problem;

/* ------------------------------------------------------------------------------------------
 * Context: device, stream, workspaces, cuBLAS/cuSOLVER handles.
 * stream: a cudaStream_t to run on, or NULL to create a private non-blocking stream.
 * ---------------------------------------------------------------------------------------- */
int sage_ba_create(sage_ba_context **ctx, int device, void *stream);
void sage_ba_destroy(sage_ba_context *ctx);
const char *sage_ba_last_error(const sage_ba_context *ctx);
const char *sage_ba_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches counter) */
long sage_ba_launch_count(const sage_ba_context *ctx);
int sage_ba_synchronize(sage_ba_context *ctx);
/* the cudaStream_t every call on this context is enqueued on (the caller's, or the private one created by sage_ba_create) */
void *sage_ba_stream(const sage_ba_context *ctx);
/* Process-wide choice of the geometric lineariser (the kernel behind df::geometric_jac_error_calculate,
 * cuda/geometric_factor_kernels.cpp:472-720): 1 = tcgen05.mma kind::tf32 with the accumulators in tensor memory (code sizes 16
 * and 32), 0 = mma.sync.  on < 0 only queries.  Returns the previous setting; the environment variable SAGE_BA_GEO_TC gives the
 * initial one.  Both produce the same J^T J to fp32 round-off (3xTF32 split either way).  A problem keeps the choice that was in
 * force when it was created (slice counts and partial buffers depend on it); single-factor calls read it per call. */
int sage_ba_set_geometric_tcgen05(int on);

/* ------------------------------------------------------------------------------------------
 * Keyframe: the immutable per-frame tensors of df::Frame<float> (core/mapping/frame.h:16-125),
 * handed over in the REFERENCE layouts and re-laid-out once on the device (channel-last
 * feature+gradient pyramid, pixel-major depth basis).  `memory` says where the pointers live.
 * ---------------------------------------------------------------------------------------- */
typedef struct sage_ba_keyframe_desc
{
  int memory;                          /* SAGE_BA_HOST or SAGE_BA_DEVICE for every pointer below */
  int height, width, levels;           /* level-0 size, pyramid levels L                          */
  int feat_channels, code_size;        /* F (16 or 32), C (8, 16 or 32)                           */
  sage_ba_camera camera;               /* level-0 camera; the pyramid is derived like CameraPyramid */
  const float *feat_map;               /* [F, H, W] feature-net output, OR NULL.  When given, the Gaussian pyramid with
                                          gradients is built on the device (Mapper::GenerateGaussianPyramidWithGrad,
                                          mapper.cpp:1385-1426) and the two pyramid pointers below are ignored.     */
  const float *feat_map_pyramid;       /* [F, SP]      Frame::feat_map_pyramid                    */
  const float *feat_map_grad_pyramid;  /* [2, F, SP]   Frame::feat_map_grad_pyramid (0:dx 1:dy)   */
  const float *dpt_map_bias;           /* [H*W]        Frame::dpt_map_bias (may be NULL for a tracked frame) */
  const float *dpt_jac_code;           /* [H*W, C] addressed with the two strides below (elements) */
  long jac_stride_row, jac_stride_col; /* reference view: (1, H*W) (code_depth_network.cpp:38-39)  */
  const float *video_mask;             /* [H, W] float 0/1   *Frame::video_mask_ptr               */
  const int64_t *sampled_locations_1d; /* [N] int64    Frame::sampled_locations_1d (may be NULL)  */
  const float *sampled_locations_homo; /* [N, 3]       Frame::sampled_locations_homo              */
  int num_samples;                     /* N                                                       */
  int borrow_depth;                    /* != 0 (DEVICE memory, dpt_jac_code already pixel-major: strides (C, 1)): dpt_map_bias,
                                          dpt_jac_code and video_mask are used IN PLACE, nothing is copied or allocated; the caller
                                          keeps them alive and unchanged while the handle is used.  For per-call views such as KF1
                                          of the reference's geometric operators (geometric_factor.cpp:340-342).               */
} sage_ba_keyframe_desc;

int sage_ba_keyframe_create(sage_ba_context *ctx, const sage_ba_keyframe_desc *desc, sage_ba_keyframe **kf);
void sage_ba_keyframe_destroy(sage_ba_context *ctx, sage_ba_keyframe *kf);
/* Replace the keyframe's dpt_map_bias ([H*W], HOST or DEVICE per `memory`) in place.  Used by the df:: shim, which receives
 * KF1's state-dependent depth map instead of (bias, code) from the reference's geometric operators. */
int sage_ba_keyframe_set_bias(sage_ba_context *ctx, sage_ba_keyframe *kf, const float *dpt_map_bias, int memory);
/* the CameraPyramid<float> derived for this keyframe (common/camera_pyramid.h:18-32) */
int sage_ba_keyframe_cameras(const sage_ba_keyframe *kf, sage_ba_camera *cams /* [levels] */, int *level_offsets);

/* ------------------------------------------------------------------------------------------
 * Single-factor entry points == the reference operator API.  Synchronous: results are in the
 * HOST output buffers on return (the reference's callers copy them to the host immediately,
 * gtsam/photometric_factor.cpp:304-306).  n_inliers may be NULL.
 * ---------------------------------------------------------------------------------------- */

/* df::photometric_jac_error_calculate<CS,FS>  cuda/photometric_factor_kernels.cpp:1061-1164 */
int sage_ba_photometric_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1,
                                  const float *R10, const float *t10, const float *R0, const float *t0,
                                  const float *R1, const float *t1, const float *code0, float scale0, float eps,
                                  const float *weights /* [L] */, float *AtA, float *Atb, float *error,
                                  float *n_inliers);

/* df::photometric_error_calculate<FS>  cuda/photometric_factor_kernels.cpp:990-1059 */
int sage_ba_photometric_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1,
                              const float *R10, const float *t10, const float *code0, float scale0, float eps,
                              const float *weights, float *error, float *n_inliers);

/* df::tracker_photo_jac_error_calculate<FS> (:1166-1245) and ..._with_scale (:1247-1325).
 * with_scale != 0 -> 7x7 system with the scale column (scale0 used), else 6x6.
 * sampled_dpts_0 [N], sampled_locations_homo_0 [N,3], sampled_features_0 [L,N,F] are DEVICE
 * pointers (they are device tensors in the tracker, camera_tracker.cpp:1086-1123). */
int sage_ba_tracker_photo_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *frame1, const float *R,
                                    const float *t, const float *sampled_dpts_0, const float *sampled_locations_homo_0,
                                    const float *sampled_features_0, int num_samples, int with_scale, float scale0,
                                    float eps, const float *weights, float *AtA, float *Atb, float *error,
                                    float *n_inliers);

/* df::tracker_photo_error_calculate<FS>  (:1327-1384) */
int sage_ba_tracker_photo_error(sage_ba_context *ctx, const sage_ba_keyframe *frame1, const float *R, const float *t,
                                const float *sampled_dpts_0, const float *sampled_locations_homo_0,
                                const float *sampled_features_0, int num_samples, float eps, const float *weights,
                                float *error, float *n_inliers);

/* The tracker's one-time pre-sampling of the keyframe features at its own sample points
 * (camera_tracker.cpp:1104-1123: grid_sample, bilinear, zeros padding, align_corners=false),
 * producing the DEVICE tensors the two calls above take.  out_* are device buffers owned by the
 * caller: dpts [N], homo [N,3], feats [L,N,F].  dpts = Keyframe::dpt_map at the sample points
 * = scale0 * (bias + jac . code0). */
int sage_ba_tracker_presample(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const float *code0, float scale0,
                              float *out_dpts, float *out_homo, float *out_feats);

/* df::geometric_jac_error_calculate<CS>  cuda/geometric_factor_kernels.cpp:882-950.
 * The per-call preparation the reference caller does (depth map of KF1 from code1, its spatial
 * gradient, the pixel-major basis copy; gtsam/geometric_factor.cpp:317-320,340-342) happens
 * inside, from code1/scale1. */
int sage_ba_geometric_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1,
                                const float *R10, const float *t10, const float *R0, const float *t0,
                                const float *R1, const float *t1, const float *code0, const float *code1,
                                float scale0, float scale1, float eps, float loss_param, float weight, float *AtA,
                                float *Atb, float *error, float *n_inliers);

/* df::geometric_error_calculate<CS>  (:837-880) */
int sage_ba_geometric_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1,
                            const float *R10, const float *t10, const float *code0, const float *code1, float scale0,
                            float scale1, float eps, float loss_param, float weight, float *error,
                            float *n_inliers);

/* df::reprojection_jac_error_calculate<CS>  cuda/reprojection_factor_kernels.cpp:467-538.
 * matched_locations_1d_0 [M] int32, matched_locations_homo_0 [M,3], matched_locations_2d_1 [M,2]
 * are HOST arrays (M <= 4096). */
int sage_ba_reprojection_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const float *R10,
                                   const float *t10, const float *R0, const float *t0, const float *R1,
                                   const float *t1, const float *code0, float scale0, const int32_t *loc1d,
                                   const float *homo, const float *match2d, int num_matches, float eps,
                                   float loss_param, float weight, float *AtA, float *Atb, float *error,
                                   float *n_inliers);

/* df::reprojection_error_calculate<CS>  (:421-465) */
int sage_ba_reprojection_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const float *R10, const float *t10,
                               const float *code0, float scale0, const int32_t *loc1d, const float *homo,
                               const float *match2d, int num_matches, float eps, float loss_param, float weight,
                               float *error, float *n_inliers);

/* df::tracker_reproj_jac_error_calculate (:540-600) / df::tracker_reproj_error_calculate (:602-642).
 * dpts [M], homo [M,3], match2d [M,2] are HOST arrays. */
int sage_ba_tracker_reproj_jac_error(sage_ba_context *ctx, const sage_ba_camera *camera, const float *R,
                                     const float *t, const float *dpts, const float *homo, const float *match2d,
                                     int num_matches, float eps, float loss_param, float weight, float *AtA,
                                     float *Atb, float *error, float *n_inliers);
int sage_ba_tracker_reproj_error(sage_ba_context *ctx, const sage_ba_camera *camera, const float *R, const float *t,
                                 const float *dpts, const float *homo, const float *match2d, int num_matches,
                                 float eps, float loss_param, float weight, float *error, float *n_inliers);

/* df::tracker_match_geom_jac_error_calculate (cuda/match_geometry_factor_kernels.cpp:1385-1421), ..._with_scale
 * (:1423-1459) and tracker_match_geom_error_calculate (:1361-1383): 3-D point-to-point term between matched keypoints,
 * Fair loss per axis, normalised by the number of matches.  All arrays are HOST arrays; sampled_dpts_0 is the (scaled)
 * depth exactly as the reference passes it.  with_scale != 0 -> 7x7 system (scale_0 column). */
int sage_ba_tracker_match_geom_jac_error(sage_ba_context *ctx, const float *R, const float *t, const float *sampled_dpts_0,
                                         const float *matched_dpts_1, const float *sampled_locations_homo_0,
                                         const float *matched_locations_homo_1, int num_matches, int with_scale, float scale0,
                                         float loss_param, float weight, float *AtA, float *Atb, float *error);
int sage_ba_tracker_match_geom_error(sage_ba_context *ctx, const float *R, const float *t, const float *sampled_dpts_0,
                                     const float *matched_dpts_1, const float *sampled_locations_homo_0,
                                     const float *matched_locations_homo_1, int num_matches, float loss_param, float weight,
                                     float *error);

/* df::match_geometry_jac_error_calculate<CS> (cuda/match_geometry_factor_kernels.cpp:1675-1823) and
 * df::match_geometry_error_calculate<CS> (:1567-1673): 3-D point-to-point term between matched keypoints of two
 * keyframes, depths from each keyframe's bias/basis/code/scale.  Variable order of the (14+2C)-square system:
 * [pose0 6 | pose1 6 | code0 C | code1 C | scale0 | scale1].  loss_type mirrors the reference's robust_loss_type
 * string.  Location / homogeneous arrays are HOST arrays of num_matches entries. */
enum
{
  SAGE_BA_LOSS_FAIR = 0,     /* "fair"     */
  SAGE_BA_LOSS_L2 = 1,       /* "L2"       */
  SAGE_BA_LOSS_HUBER = 2,    /* "huber"    */
  SAGE_BA_LOSS_UNBIASED = 3  /* "unbiased" */
};
int sage_ba_match_geometry_jac_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1,
                                     const float *R10, const float *t10, const float *R0, const float *t0, const float *R1,
                                     const float *t1, const float *code0, const float *code1, float scale0, float scale1,
                                     const int32_t *sampled_locations_1d_0, const int32_t *matched_locations_1d_1,
                                     const float *sampled_locations_homo_0, const float *matched_locations_homo_1,
                                     int num_matches, float loss_param, float weight, int loss_type, float *AtA, float *Atb,
                                     float *error);
int sage_ba_match_geometry_error(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *kf1,
                                 const float *R10, const float *t10, const float *code0, const float *code1, float scale0,
                                 float scale1, const int32_t *sampled_locations_1d_0, const int32_t *matched_locations_1d_1,
                                 const float *sampled_locations_homo_0, const float *matched_locations_homo_1,
                                 int num_matches, float loss_param, float weight, int loss_type, float *error);

/* df::loop_mg_jac_error_calculate (:1511-1565) and df::loop_mg_error_calculate (:1479-1509): the loop-closure
 * pose/scale graph's factor (core/deepfactors.cpp:409-573).  14x14 system, order [pose0 6 | pose1 6 | scale0 | scale1];
 * the depth arrays are the UNSCALED depths of the matched keypoints (HOST arrays). */
int sage_ba_loop_mg_jac_error(sage_ba_context *ctx, const float *R10, const float *t10, const float *R0, const float *t0,
                              const float *R1, const float *t1, const float *sampled_unscaled_dpts_0,
                              const float *matched_unscaled_dpts_1, const float *sampled_locations_homo_0,
                              const float *matched_locations_homo_1, int num_matches, float scale0, float scale1,
                              float loss_param, float weight, float *AtA, float *Atb, float *error);
int sage_ba_loop_mg_error(sage_ba_context *ctx, const float *R10, const float *t10, const float *sampled_unscaled_dpts_0,
                          const float *matched_unscaled_dpts_1, const float *sampled_locations_homo_0,
                          const float *matched_locations_homo_1, int num_matches, float scale0, float scale1,
                          float loss_param, float weight, float *error);

/* ------------------------------------------------------------------------------------------
 * Dense descriptor cycle-matching: the torch expression chain in the
 * constructors of ReprojectionFactor (core/gtsam/reprojection_factor.cpp:57-92) and
 * MatchGeometryFactor (core/gtsam/match_geometry_factor.cpp:62-97) and in
 * CameraTracker::FeatureMatchingGeo (core/system/camera_tracker.cpp:798-834):
 *   m_k = argmax_p -sum_c (desc0[c, kp_k] - desc1[c, p])^2,   c_k = argmax_p -sum_c (desc1[c, m_k] - desc0[c, p])^2,
 *   keypoint k is an inlier when |pixel(kp_k) - pixel(c_k)|^2 <= cyc_consis_thresh^2.
 * feat_desc_{0,1}: [channels, H, W] (Frame::feat_desc, channel-major), HOST or DEVICE per `memory`;
 * channels in {8, 16, 32, 64}.  keypoint_locations_1d: HOST [K] int64 = valid_locations_1d[keypoint_indexes]
 * (the reference's mt19937 shuffle stays with the caller).  Outputs (HOST; any may be NULL except num_inliers):
 *   raw_matched_locations_1d_1 [K], cyc_matched_locations_1d_0 [K],
 *   inlier_positions [<= K]: positions 0..K-1 of the inlier keypoints in ascending order (torch::nonzero order), so
 *   matched_keypoint_indexes = keypoint_indexes[inlier_positions], matched_locations_1d_1 = raw[inlier_positions].
 * kernel_ms: optional, device time of the six launches (CUDA events on the context stream).
 * Responses are evaluated in fp32 exactly as written above (unfused, channel order), ties keep the lowest pixel index.
 * ---------------------------------------------------------------------------------------- */
int sage_ba_cycle_match(sage_ba_context *ctx, int memory, const float *feat_desc_0, const float *feat_desc_1, int channels, int height,
                        int width, const int64_t *keypoint_locations_1d, int num_keypoints, float cyc_consis_thresh,
                        int32_t *raw_matched_locations_1d_1, int32_t *cyc_matched_locations_1d_0, int32_t *inlier_positions,
                        int *num_inliers, float *kernel_ms);

/* ------------------------------------------------------------------------------------------
 * CameraTracker::TrackNewFrame LM loop (core/system/camera_tracker.cpp:1034-1310, loop
 * :1156-1279): damped Gauss-Newton on the 6-DoF relative pose T_ck of `frame1` w.r.t. `kf0`,
 * photometric (+ optional reprojection) terms, same damping / acceptance / convergence rules.
 * ---------------------------------------------------------------------------------------- */
typedef struct sage_ba_tracker_config
{                                    /* TrackerConfig, configs/slam_run.flags:17-23 */
  int max_num_iters;                 /* 40   */
  float init_damp, min_damp, max_damp; /* 1e-4, 1e-6, 1e-2 */
  float damp_dec_factor, damp_inc_factor; /* 10, 100 */
  float jac_update_err_inc_threshold; /* 1e-2 */
  float min_grad_thresh, min_param_inc_thresh;
  float dpt_eps;
  float photo_weights[SAGE_BA_MAX_LEVELS];
  int use_photo, use_reproj;
  float reproj_loss_param, reproj_weight; /* loss param and inlier_multiplier*factor weight */
  int use_match_geom;                     /* TrackFrame (7-DoF) only */
  float match_geom_loss_param, match_geom_weight;
} sage_ba_tracker_config;

typedef struct sage_ba_tracker_report
{
  int iterations, jacobian_evals, error_evals;
  float final_error, final_damp;
} sage_ba_tracker_report;

/* R, t: in = initial guess of T10 (frame1 <- kf0), out = estimate.  Match arrays (HOST) may be
 * NULL when use_reproj == 0. */
int sage_ba_track_new_frame(sage_ba_context *ctx, const sage_ba_keyframe *kf0, const sage_ba_keyframe *frame1,
                            const float *code0, float scale0, const sage_ba_tracker_config *cfg, float *R, float *t,
                            const float *match_dpts, const float *match_homo, const float *match2d, int num_matches,
                            sage_ba_tracker_report *report);

/* CameraTracker::TrackFrame (core/system/camera_tracker.cpp:1312-1672, loop :1479-1630): 7-DoF LM on the relative pose AND
 * the depth scale of `frame0` (the frame being tracked, sampled at its own points) against the reference keyframe `kf1`;
 * photometric term with the scale column + optional match-geometry term.  R, t, scale: in = initial guess, out = estimate.
 * match arrays (HOST): unscaled keypoint depths of frame0 [M], their rays [M,3], matched depths / rays in kf1.
 * Returns 0 on success; 2 when the photometric term reports no overlap and no match-geometry term is enabled (:1500-1504). */
int sage_ba_track_frame(sage_ba_context *ctx, const sage_ba_keyframe *frame0, const sage_ba_keyframe *kf1, const float *code0,
                        const sage_ba_tracker_config *cfg, float *R, float *t, float *scale, const float *match_unscaled_dpts_0,
                        const float *match_homo_0, const float *match_dpts_1, const float *match_homo_1, int num_matches,
                        sage_ba_tracker_report *report);

/* Host helpers of the tracker loops, exposed for callers that keep their own LM loop (and pinned against the reference's Eigen
 * code, tests/golden/host_pins.npz): x = (AtA + damp diag(AtA)).colPivHouseholderQr().solve(Atb), n = 6 | 7, float
 * (core/system/camera_tracker.cpp:1182-1183); se3_exp<float> (core/mapping/mapping_utils.h:316-346), R [9] row-major, t [3]. */
int sage_ba_tracker_solve(const float *AtA, const float *Atb, int n, float damp, float *x);
int sage_ba_se3_exp(const float *omega, const float *v, float *R, float *t);

/* The LM loop behind sage_ba_track_new_frame (dof = 6; core/system/camera_tracker.cpp:1156-1279) and sage_ba_track_frame
 * (dof = 7; :1479-1630) over a caller-supplied cost: jac fills AtA [dof*dof] row-major, Atb [dof] and the error at (R, t, scale),
 * err returns the error.  Host code only (no context, no GPU): for callers with their own factors, and the hook through which the
 * loop itself is pinned against the reference's loops (tests/golden/loop_pins.npz, tests/test_loop_pins.py).  Only the LM fields
 * of cfg are read.  R, t (and scale, dof = 7): in = initial guess, out = estimate.  Returns 0. */
typedef void (*sage_ba_lm_jac_callback)(void *user, const float *R, const float *t, float scale, float *AtA, float *Atb, float *error);
typedef float (*sage_ba_lm_err_callback)(void *user, const float *R, const float *t, float scale);
int sage_ba_tracker_lm_callbacks(int dof, const sage_ba_tracker_config *cfg, float *R, float *t, float *scale, sage_ba_lm_jac_callback jac,
                                 sage_ba_lm_err_callback err, void *user, sage_ba_tracker_report *report);

/* ------------------------------------------------------------------------------------------
 * Batched local bundle adjustment (new; the reference delegates this to GTSAM ISAM2,
 * core/mapping/mapper.cpp:544).  A problem owns K keyframes' states (pose_wk, code, scale) on
 * the device and a factor list; one launch linearises every factor of a kind.
 * Global variable order: [pose_0 .. pose_{K-1} (6 each) | (code_k (C), scale_k) for k = 0..K-1].
 * ---------------------------------------------------------------------------------------- */
int sage_ba_problem_create(sage_ba_context *ctx, int num_keyframes, sage_ba_keyframe *const *kfs,
                           sage_ba_problem **problem);
void sage_ba_problem_destroy(sage_ba_problem *problem);

int sage_ba_problem_add_photometric(sage_ba_problem *p, int kf0, int kf1, const float *weights /* [L] */);
int sage_ba_problem_add_geometric(sage_ba_problem *p, int kf0, int kf1, float loss_param, float weight);
int sage_ba_problem_add_reprojection(sage_ba_problem *p, int kf0, int kf1, const int32_t *loc1d, const float *homo,
                                     const float *match2d, int num_matches, float loss_param, float weight);
/* CodeFactor (gtsam/code_factor.cpp:42-104): AtA = w I, Atb = w (init - code), err = w mean((init-code)^2) */
int sage_ba_problem_add_code_prior(sage_ba_problem *p, int kf, const float *init_code, float weight);
/* ScaleFactor (gtsam/scale_factor.cpp:115-130) */
int sage_ba_problem_add_scale_prior(sage_ba_problem *p, int kf, float init_scale, float weight);
/* hold a keyframe's pose (and optionally scale) fixed: the gauge anchor (mapper.cpp:190-192) */
int sage_ba_problem_fix(sage_ba_problem *p, int kf, int fix_pose, int fix_scale);
/* linear solver (stands where ISAM2's multifrontal Cholesky stands, core/mapping/mapper.cpp:544):
 * 0 (default) hand-written block-sparse Cholesky over keyframes: eliminating a keyframe's [pose | code | scale] block column is
 *   the Schur complement of that keyframe onto the keyframes it shares factors with; one CTA per block column, columns run as
 *   soon as the columns they depend on are done (nested-dissection order over the keyframe index line); fp64; 2 launches;
 * 1 cross-check: the dense system with the code+scale block ordered first and ONE cuSOLVER potrf (the Schur complement onto the
 *   6K x 6K pose block forms in the trailing sub-matrix) -- O(dim^2) memory, library kernels;
 * 2 as 0 with the natural (temporal) elimination order. */
int sage_ba_problem_set_solver(sage_ba_problem *p, int solver);
/* Multi-GPU sharding, one process per GPU.  The distinct ordered pairs (kf0 -> kf1) of the factor list, sorted by (kf0, kf1), are
 * cut into `world` equal runs and every factor of a pair belongs to the pair's rank (sage_ba_shard_plan): a keyframe's pairs are
 * consecutive in that order, so its maps are read by one GPU (two at a cut), and the ranks' loads differ by at most one pair.
 * Every rank adds EVERY factor (the list defines the normal equations); a rank needs the device data only of the keyframes its
 * own pairs touch (host of an owned pair: depth + samples + features; target: features + mask (+ depth for geometric factors))
 * and may pass NULL for the others in sage_ba_problem_create.  Call set_shard before the first linearisation. */
int sage_ba_problem_set_shard(sage_ba_problem *p, int rank, int world);
/* owner rank of each of `num_pairs` ordered pairs (any order, duplicates allowed) under the rule above */
int sage_ba_shard_plan(int num_pairs, const int *pair_kf0, const int *pair_kf1, int world, int *owner);
/* sage_ba_problem_lm_step re-uses the linearisation after a rejected step (the state did not move); always != 0 forces the full
 * iteration every time (what BASELINE's "LM iteration" metric counts). */
int sage_ba_problem_set_relinearize_always(sage_ba_problem *p, int always);
/* CTAs per factor are normally sized so that a launch ends on a full wave of THIS rank's factors (fastest; a factor's outputs
 * then agree across GPU counts to fp32 round-off).  on != 0 makes the decomposition depend on the problem only: with the
 * fixed-order assembly and the flag-ordered solver the whole LM trajectory is then bit-identical for every number of GPUs, at
 * ~5-10 % of the factor kernels' speed.  Call before the first linearisation. */
int sage_ba_problem_set_deterministic(sage_ba_problem *p, int on);

int sage_ba_problem_set_state(sage_ba_problem *p, const float *poses /* [K,12] R row-major then t */,
                              const float *codes /* [K,C] */, const float *scales /* [K] */, float eps);
int sage_ba_problem_get_state(sage_ba_problem *p, float *poses, float *codes, float *scales);

/* Mapper::UpdateMap (core/mapping/mapper.cpp:1141-1180): hand the accepted estimate back to the map.  poses [K,12],
 * codes [K,C], scales [K] are HOST; dpt_maps [K, H*W] (HOST or DEVICE per `memory`) receives, per keyframe,
 * UpdateDepth(...) = scale * (dpt_map_bias + dpt_jac_code . code)  (core/mapping/mapping_utils.h:216-222).  Any may be NULL. */
int sage_ba_problem_update_map(sage_ba_problem *p, float *poses, float *codes, float *scales, float *dpt_maps, int memory);
int sage_ba_problem_dim(const sage_ba_problem *p);         /* K * (7 + C)                       */
int sage_ba_problem_num_factors(const sage_ba_problem *p); /* all kinds, priors included        */
long sage_ba_problem_num_residuals(const sage_ba_problem *p); /* scalar residual rows per linearisation */
/* packed per-factor output buffer (DEVICE, fp32): every factor's [AtA | Atb | error | inliers], laid out as `world` equal
 * segments, segment r holding the factors rank r owns in order of addition (world = 1: simply the order of addition).  This is
 * the buffer the ranks exchange once per LM iteration (in-place all-gather of the segments). */
int sage_ba_problem_factor_buffer(sage_ba_problem *p, float **device_ptr, size_t *count);
/* per factor (order of addition): offset in the factor buffer, offset in the cost buffer, owner rank; any may be NULL */
int sage_ba_problem_factor_offsets(sage_ba_problem *p, int *offsets, int *cost_offsets, int *owners);
/* symbolic factorisation of solver 0 / 2: blocks of the factor (diagonal + sub-diagonal incl. fill), fill blocks, and the
 * longest dependency chain of block columns (the critical path of the elimination) */
int sage_ba_problem_solver_info(sage_ba_problem *p, int *num_blocks, int *fill_blocks, int *depth);
/* tuning aid (environment SAGE_BA_SOLVER_TRACE=1 at solve time): per block column, in elimination order, the %globaltimer
 * stamps (ns) [start, column loaded, updates done, diagonal factored, panel solved, published, ns waiting on flags, #dependencies];
 * out [K][8], positions_to_keyframes [K] (may be NULL) */
int sage_ba_problem_solver_trace(sage_ba_problem *p, long long *out, int *positions_to_keyframes);
/* packed per-factor [error | inliers] buffer of the last cost evaluation (DEVICE, fp32) */
int sage_ba_problem_cost_buffer(sage_ba_problem *p, float **device_ptr, size_t *count);

/* stage 1: linearise this shard's factors at the current state into the factor buffer (async) */
int sage_ba_problem_linearize(sage_ba_problem *p);
/* stage 2: assemble H (fp64, dense) and g from the (reduced) factor buffer, add priors (async).
 * H/g out may be NULL; otherwise HOST buffers [dim*dim] / [dim] filled synchronously. */
int sage_ba_problem_assemble(sage_ba_problem *p, double *H, double *g, double *cost);
/* stage 3: solve (H + damp diag(H)) d = g by Schur complement of the code+scale block onto the
 * pose block (cuSOLVER potrf/potrs), write the candidate state x (+) d; delta may be NULL. */
int sage_ba_problem_solve(sage_ba_problem *p, double damp, double *delta /* HOST [dim] or NULL */);
/* stage 4: evaluate this shard's factor errors at the candidate (which=1) or current (which=0)
 * state into the cost buffer (async) */
int sage_ba_problem_evaluate(sage_ba_problem *p, int which);
/* stage 5: sum the (reduced) cost buffer + priors -> total cost (synchronous) */
int sage_ba_problem_cost(sage_ba_problem *p, int which, double *cost);
/* stage 6: make the candidate the current state */
int sage_ba_problem_accept(sage_ba_problem *p);

/* CUDA-event timing of the launches inside linearize / evaluate / assemble / solve, per kind (ms summed and
 * launch counts since the last reset).  Used by bench.py for the roofline line; off by default.  While it is on, the three
 * factor kinds run one after the other on the context's stream (they normally overlap on forked streams), so that a kind's
 * time is its own. */
enum
{
  SAGE_BA_PROF_PHOTO_JAC = 0,
  SAGE_BA_PROF_GEO_JAC = 1,
  SAGE_BA_PROF_REPROJ_JAC = 2,
  SAGE_BA_PROF_PHOTO_ERR = 3,
  SAGE_BA_PROF_GEO_ERR = 4,
  SAGE_BA_PROF_REPROJ_ERR = 5,
  SAGE_BA_PROF_DEPTH_PREP = 6,
  SAGE_BA_PROF_ASSEMBLE = 7,
  SAGE_BA_PROF_SOLVE = 8,
  SAGE_BA_PROF_COMM = 9,
  SAGE_BA_PROF_KINDS = 10
};
int sage_ba_problem_profile(sage_ba_problem *p, int enable);
int sage_ba_problem_profile_read(sage_ba_problem *p, double *ms /* [KINDS] */, long *counts /* [KINDS] */, int reset);
/* factors of each kind this process's shard owns (valid after the first linearize / buffer query) */
int sage_ba_problem_shard_counts(const sage_ba_problem *p, int *n_photo, int *n_geo, int *n_reproj);

/* Collectives.  Preferred: a NCCL communicator owned by the library -- the exchange is then issued from C++ on the context's
 * stream inside linearize/evaluate's callers (sage_ba_problem_lm_step / _lm / _exchange), one ncclAllGather of the owned segment
 * per linearisation and one (a few hundred bytes) per trial step.  libnccl is bound at run time (the copy already loaded in the
 * process is used when there is one).  Rank 0 draws an id, the host program distributes the 128 bytes by any means. */
typedef struct sage_ba_comm sage_ba_comm;
int sage_ba_nccl_unique_id(char id[128]);
int sage_ba_comm_create(sage_ba_context *ctx, const char id[128], int rank, int world, sage_ba_comm **comm);
int sage_ba_comm_wrap(void *nccl_comm /* ncclComm_t of the caller */, int rank, int world, sage_ba_comm **comm);
void sage_ba_comm_destroy(sage_ba_comm *comm);
int sage_ba_problem_set_comm(sage_ba_problem *p, sage_ba_comm *comm);
/* make the factor buffer (which = 0) or the cost buffer (which = 1) complete on every rank (no-op for world = 1) */
int sage_ba_problem_exchange(sage_ba_problem *p, int which);
/* Alternative for hosts without NCCL: a callback that sum-all-reduces a device buffer (other ranks' segments are zero). */
typedef int (*sage_ba_allreduce_fn)(void *device_buffer, size_t count, void *user); /* fp32 sum, on ctx stream */
int sage_ba_problem_set_allreduce(sage_ba_problem *p, sage_ba_allreduce_fn fn, void *user);

typedef struct sage_ba_lm_options
{
  int max_iters;
  double init_damp, min_damp, max_damp, damp_dec_factor, damp_inc_factor;
  double min_rel_decrease; /* stop when an accepted step lowers the cost by less than this fraction */
  int max_trials;          /* damping increases per iteration before giving up */
} sage_ba_lm_options;

typedef struct sage_ba_lm_report
{
  int iterations, linearizations, evaluations, accepted;
  double initial_cost, final_cost, final_damp;
} sage_ba_lm_report;

/* One LM iteration exactly as BASELINE defines it (linearise every factor at the current estimate -> assemble -> Schur solve
 * with *damp -> evaluate the candidate -> accept / reject), enqueued back to back with ONE host synchronisation that returns
 * both costs.  *damp is updated by the tracker's rule (divide by damp_dec_factor on acceptance, multiply by damp_inc_factor
 * otherwise, clamped; camera_tracker.cpp:1218-1245).  The reference's hand-written LM syncs >= 6 times per iteration. */
int sage_ba_problem_lm_step(sage_ba_problem *p, double *damp, double min_damp, double max_damp, double damp_dec_factor,
                            double damp_inc_factor, double *cost, double *candidate_cost, int *accepted);

/* Full LM loop (linearize -> [allreduce] -> assemble -> solve -> evaluate -> [allreduce] ->
 * accept/reject), acceptance rule as the tracker's (camera_tracker.cpp:1218-1245). */
int sage_ba_problem_lm(sage_ba_problem *p, const sage_ba_lm_options *opt, sage_ba_lm_report *report);

#ifdef __cplusplus
}
#endif
#endif /* SAGE_BA_H_ */
