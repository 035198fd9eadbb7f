/* ppsfm_b200.h — C-ABI of libppsfm_b200.so, the B200 (sm_100a) implementation of the
 * privacy-preserving-SfM hot path (line-lifted absolute pose under RANSAC, line-reprojection
 * bundle adjustment).
 *
 * The reference (colmap/privacy_preserving_sfm) has no FFI layer; the seam is its C++ API.
 * Each entry point below names the reference interface it replaces (file:line, relative to the
 * upstream repository).  The header-only C++ adaptor in privacy_preserving_sfm_b200/cpp/
 * re-creates the reference signatures on top of this ABI.
 *
 * Conventions
 *   - all pointers are HOST memory, caller-allocated, unless the name says `_resident`;
 *   - calls are blocking; one context per host thread (a context owns one CUDA device, its
 *     streams, scratch buffers and the PRNG that mirrors the reference's thread_local mt19937);
 *   - return value: PPSFM_OK (0); PPSFM_NO_SOLUTION (1) where the reference returns `false`;
 *     negative = contract violation / CUDA failure (where the reference CHECK-aborts) —
 *     ppsfm_last_error() describes it.  There is NO CPU fallback: without a CUDA device
 *     ppsfm_ctx_create fails with PPSFM_ERR_CUDA.
 *   - lines  : n x 3 doubles row-major (a,b,c), normalised camera coordinates, ||(a,b)|| = 1
 *              (FeatureLine::Line(), src/feature/types.h:98-138)
 *     aligned: n bytes (FeatureLine::IsAligned()); may be NULL (= all false)
 *     points : n x 3 doubles row-major (std::vector<Eigen::Vector3d> memory)
 *     model  : 12 doubles = Eigen::Matrix3x4d column-major (R col0, R col1, R col2, t)
 *     qvec   : (w, x, y, z)  (src/base/pose.cc:41-44)
 */
#ifndef PPSFM_B200_H_
#define PPSFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPSFM_OK 0
#define PPSFM_NO_SOLUTION 1
#define PPSFM_ERR_INVALID (-1)
#define PPSFM_ERR_CUDA (-2)
#define PPSFM_ERR_NCCL (-3)

typedef struct ppsfm_ctx ppsfm_ctx;
typedef struct ppsfm_corr ppsfm_corr; /* correspondence set resident in HBM */

/* RANSACOptions, src/optim/ransac.h:47-76 (same fields, same defaults via ppsfm_ransac_options_default) */
typedef struct {
  double max_error;
  double min_inlier_ratio;
  double confidence;
  double dyn_num_trials_multiplier;
  uint64_t min_num_trials;
  uint64_t max_num_trials;
} ppsfm_ransac_options;

/* RANSAC<P6LEstimator>::Report, src/optim/ransac.h:82-99 (+ provenance of the winner) */
typedef struct {
  int32_t success;
  uint64_t num_trials;
  uint64_t num_inliers;  /* support.num_inliers */
  double residual_sum;   /* support.residual_sum (summed in index order, as the reference) */
  double model[12];
  int64_t best_trial;     /* trial index that produced the best model, -1 if none */
  int32_t best_model_idx; /* index inside that trial's solver output */
  uint64_t num_models_scored;
} ppsfm_ransac_report;

/* Device-side timings of the last RANSAC call (CUDA events on the context's stream). */
typedef struct {
  double solve_ms;   /* batched P6L / re3q3 kernel(s) */
  double score_ms;   /* batched scoring kernel(s) (the dominant kernel) */
  double exact_ms;   /* exact-order support + mask kernels */
  double total_ms;   /* first launch to last kernel end */
  uint64_t score_pairs;    /* (model, correspondence) pairs evaluated by the scoring kernel */
  uint64_t score_launches; /* scoring-kernel launches */
  uint64_t kernel_launches; /* all kernels launched by the call */
  double comm_ms;    /* sharded call: all-reduce of the counts + pruning-bound update */
} ppsfm_ransac_timing;

/* ---- context -------------------------------------------------------------------------- */
int ppsfm_ctx_create(int device, ppsfm_ctx** out);
void ppsfm_ctx_destroy(ppsfm_ctx* ctx);
const char* ppsfm_last_error(const ppsfm_ctx* ctx);
const char* ppsfm_version(void);

/* SetPRNGSeed, src/util/random.cc:38-50 (the context's generator starts at seed 0 like
 * kDefaultPRNGSeed, src/util/random.h:46). */
void ppsfm_set_prng_seed(ppsfm_ctx* ctx, uint32_t seed);
/* Host-only self-test (no context, no GPU): the event-driven replay of the reference's trial loop
 * (src/optim/ransac.h:213-249) used by ppsfm_ransac_p6l against the literal model-by-model loop
 * on `rounds` random waves.  Returns the number of waves on which they differ (0 = pass). */
int ppsfm_selftest_replay(uint32_t seed, int rounds);
/* next raw mt19937 output without advancing (state fingerprint for parity tests) */
uint32_t ppsfm_prng_peek(const ppsfm_ctx* ctx);

void ppsfm_ransac_options_default(ppsfm_ransac_options* opt);

/* RANSAC<P6LEstimator>::ComputeNumTrials, src/optim/ransac.h:158-176 */
uint64_t ppsfm_compute_num_trials(uint64_t num_inliers, uint64_t num_samples, double confidence,
                                  double num_trials_multiplier);

/* RandomSampler::Sample x num_trials on a freshly Initialize()d sampler,
 * src/optim/random_sampler.cc:40-62 + src/util/random.h:120-128; advances the context PRNG. */
int ppsfm_sample_table(ppsfm_ctx* ctx, size_t n, size_t num_trials, uint32_t* table_out);

/* ---- A5 + A8: batched scoring ------------------------------------------------------------
 * P6LEstimator::Residuals -> ComputeSquaredLineReprojectionError (src/estimators/utils.cc:40-89)
 * for K models at once, fused with InlierSupportMeasurer::Evaluate
 * (src/optim/support_measurement.cc:36-60).  residuals_out (K x n, may be NULL),
 * num_inliers_out[K], residual_sum_out[K] (index-order sums, bit-identical to the reference). */
int ppsfm_line_residuals(ppsfm_ctx* ctx, const double* lines, const double* points, size_t n,
                         const double* models, size_t num_models, double max_residual,
                         double* residuals_out, uint64_t* num_inliers_out,
                         double* residual_sum_out);

/* Test hook: the same inlier counts through the count-only scoring kernel of the RANSAC loop
 * (division-free filter + reference fallback), for tests that aim at the filter's edge cases. */
int ppsfm_score_models(ppsfm_ctx* ctx, const double* lines, const double* points, size_t n,
                       const double* models, size_t num_models, double max_residual,
                       uint32_t* counts_out);

/* ---- A3 + A4: batched minimal solver -------------------------------------------------------
 * P6LEstimator::Estimate (src/estimators/absolute_pose.cc:79-162) incl. re3q3
 * (lib/re3q3/re3q3/re3q3.h:16-200) for H samples of 6 correspondences, one thread per
 * hypothesis.  models_out: H x 8 x 12, num_models_out: H. */
int ppsfm_p6l_solve_batch(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                          const double* points, size_t n, const uint32_t* sample_idx,
                          size_t num_samples, double* models_out, int32_t* num_models_out);

/* ---- A6 + A7: RANSAC<P6LEstimator, InlierSupportMeasurer, RandomSampler>::Estimate ---------
 * src/optim/ransac.h:144-278.  Hypotheses are generated and scored on the GPU in waves; the
 * sequential best-so-far / adaptive-abort logic is replayed exactly on the host.
 * inlier_mask: n bytes or NULL. Returns PPSFM_OK also when report->success == 0. */
int ppsfm_ransac_p6l(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                     const double* points, size_t n, const ppsfm_ransac_options* options,
                     ppsfm_ransac_report* report, uint8_t* inlier_mask);

/* ---- A9: EstimateAbsolutePoseFromLines, src/estimators/pose.cc:52-94 ------------------------
 * Returns PPSFM_OK (true) or PPSFM_NO_SOLUTION (false). report may be NULL. */
int ppsfm_estimate_absolute_pose_from_lines(ppsfm_ctx* ctx, const double* lines,
                                            const uint8_t* aligned, const double* points,
                                            size_t n, const ppsfm_ransac_options* options,
                                            double* qvec, double* tvec, uint64_t* num_inliers,
                                            uint8_t* inlier_mask, ppsfm_ransac_report* report);

/* ---- resident variants (inputs already in HBM; used for the kernel-side metric) ------------ */
int ppsfm_corr_upload(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                      const double* points, size_t n, ppsfm_corr** out);
void ppsfm_corr_free(ppsfm_ctx* ctx, ppsfm_corr* corr);
int ppsfm_ransac_p6l_resident(ppsfm_ctx* ctx, const ppsfm_corr* corr,
                              const ppsfm_ransac_options* options, ppsfm_ransac_report* report,
                              uint8_t* inlier_mask);
void ppsfm_get_ransac_timing(const ppsfm_ctx* ctx, ppsfm_ransac_timing* out);

/* ---- ONE Estimate call sharded over the GPUs of a communicator (RANSAC row) ----
 * Collective: after ppsfm_comm_init every rank calls with the SAME correspondence set, options
 * and generator state (ppsfm_set_prng_seed).  The trials of RANSAC<>::Estimate
 * (src/optim/ransac.h:178-278) are generated on every rank, the models of each wave are scored
 * 1/world per GPU, one NCCL all-reduce of the 32-bit inlier counts per wave gives every rank
 * every count (the exact-pruning bound is raised to the best count over all ranks), and every
 * rank replays the reference loop (:213-249) over the same counts: report, mask and generator
 * state are those of the single-GPU call, on every rank.  With world == 1 (or without a
 * communicator) the calls are the plain ones. */
int ppsfm_ransac_p6l_sharded(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned,
                             const double* points, size_t n, const ppsfm_ransac_options* options,
                             ppsfm_ransac_report* report, uint8_t* inlier_mask);
int ppsfm_ransac_p6l_resident_sharded(ppsfm_ctx* ctx, const ppsfm_corr* corr,
                                      const ppsfm_ransac_options* options,
                                      ppsfm_ransac_report* report, uint8_t* inlier_mask);
/* host-only shard accounting: of a wave's num_models compact models, how many rank `rank` scores
 * (blocks of 512 models dealt round-robin to the ranks) */
uint64_t ppsfm_ransac_shard_models(uint64_t num_models, int rank, int world);

/* =============================================================================================
 * Line-reprojection bundle adjustment
 * ============================================================================================= */
#define PPSFM_BA_MAX_TRACE 128

typedef struct ppsfm_ba ppsfm_ba; /* a BA problem resident in HBM */

/* The problem BundleAdjuster::SetUp builds from a Reconstruction + BundleAdjustmentConfig
 * (src/optim/bundle_adjustment.cc:326-542), flattened to arrays.  qvecs / tvecs / points are
 * updated IN PLACE like Image::Qvec/Tvec and Point3D::XYZ (:357-359).
 *   pose_flags[i]: bit0 = constant pose (config.HasConstantPose or !refine_extrinsics -> the
 *                  BundleAdjustmentConstantPoseLineCostFunction branch :383-398, and images added
 *                  by AddPointToProblem :450-487); bit1..3 = tvec[0..2] constant
 *                  (config.ConstantTvec -> SubsetParameterization :425-432)
 *   point_const[p]: ParameterizePoints :530-542 (partial tracks, ConstantPoints)
 *   camera_model:  COLMAP ids 0 SIMPLE_PINHOLE, 1 PINHOLE, 2 SIMPLE_RADIAL, 3 RADIAL, 4 OPENCV,
 *                  5 OPENCV_FISHEYE, 6 FULL_OPENCV, 7 FOV, 8 SIMPLE_RADIAL_FISHEYE,
 *                  9 RADIAL_FISHEYE, 10 THIN_PRISM_FISHEYE (src/base/camera_models.h:189-248)
 *   camera_params / camera_const: intrinsics are constant unless an options.refine_* flag is
 *                  set and camera_const[c] == 0 (ParameterizeCameras :490-528); only then is
 *                  camera_params written (Camera::Params() updated in place). */
typedef struct {
  int32_t num_images;
  double* qvecs;               /* num_images x 4 (w,x,y,z) */
  double* tvecs;               /* num_images x 3 */
  const uint8_t* pose_flags;   /* may be NULL (= all variable) */
  const int32_t* image_camera; /* index into cameras */
  int32_t num_cameras;
  const int32_t* camera_model;
  double* camera_params;       /* num_cameras x 12, zero padded */
  int32_t num_points;
  double* points;              /* num_points x 3 */
  const uint8_t* point_const;  /* may be NULL */
  int64_t num_obs;
  const int32_t* obs_image;
  const int32_t* obs_point;
  const double* obs_line;      /* num_obs x 3, ||(a,b)|| = 1 */
  const uint8_t* camera_const; /* config.IsConstantCamera per camera; may be NULL (= none) */
} ppsfm_ba_problem;

/* BundleAdjustmentOptions (src/optim/bundle_adjustment.h:49-100) + the ceres::Solver::Options
 * fields the reference sets (src/controllers/incremental_mapper.cc:196-243). */
typedef struct {
  int32_t loss_type; /* LossFunctionType: 0 TRIVIAL, 1 SOFT_L1, 2 CAUCHY */
  double loss_scale;
  int32_t max_num_iterations;
  double function_tolerance;
  double gradient_tolerance;
  double parameter_tolerance;
  int32_t max_num_consecutive_invalid_steps;
  double initial_trust_region_radius;
  double max_trust_region_radius;
  double min_trust_region_radius;
  double min_relative_decrease;
  double min_lm_diagonal;
  double max_lm_diagonal;
  int32_t jacobi_scaling;
  int32_t num_threads; /* ignored (kept for layout parity with the CPU oracle) */
  /* BundleAdjustmentOptions::refine_focal_length / refine_principal_point / refine_extra_params
   * (src/optim/bundle_adjustment.h:57-63); all 0 = constant cameras, what the reference's mapper
   * runs with (src/controllers/incremental_mapper.h:81-83) */
  int32_t refine_focal_length;
  int32_t refine_principal_point;
  int32_t refine_extra_params;
} ppsfm_ba_options;

/* The fields of ceres::Solver::Summary the reference reads / prints
 * (src/optim/bundle_adjustment.cc:544-598, src/sfm/incremental_mapper.cc:860-861). */
typedef struct {
  double initial_cost;
  double final_cost;
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t termination_type; /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE */
  int64_t num_residuals;
  int64_t num_residuals_reduced;
  int32_t num_effective_parameters_reduced;
  double total_time_s;
  double jacobian_time_s;       /* device time of the Jacobian-build kernel (CUDA events) */
  double linear_solver_time_s;  /* host wall time of reduced system + Cholesky + back-subst. */
  double final_gradient_max_norm;
  int32_t trace_len;
  double trace_cost[PPSFM_BA_MAX_TRACE];
  double trace_radius[PPSFM_BA_MAX_TRACE];
  int32_t trace_accepted[PPSFM_BA_MAX_TRACE];
  int32_t jacobian_launches;
  int64_t kernel_launches;
  /* device time (CUDA events) of the phases of the linear solve, summed over iterations */
  double schur_time_s;     /* reduced camera system assembly */
  double cholesky_time_s;  /* dense factorisation + triangular solves */
  double backsub_time_s;   /* back-substitution, model cost, candidate state + cost */
} ppsfm_ba_summary;

void ppsfm_ba_options_default(ppsfm_ba_options* opt);

/* BundleAdjuster::Solve (src/optim/bundle_adjustment.cc:260-320).  Returns PPSFM_OK (true) or
 * PPSFM_NO_SOLUTION (zero residuals, :269-271). */
int ppsfm_ba_solve(ppsfm_ctx* ctx, const ppsfm_ba_problem* problem,
                   const ppsfm_ba_options* options, ppsfm_ba_summary* summary);

/* Resident variant: upload once, run (repeatedly, after ppsfm_ba_reset), download. */
int ppsfm_ba_create(ppsfm_ctx* ctx, const ppsfm_ba_problem* problem,
                    const ppsfm_ba_options* options, ppsfm_ba** out);
int ppsfm_ba_run(ppsfm_ba* ba, ppsfm_ba_summary* summary);
int ppsfm_ba_reset(ppsfm_ba* ba);
int ppsfm_ba_download(ppsfm_ba* ba, const ppsfm_ba_problem* problem);
void ppsfm_ba_free(ppsfm_ba* ba);

/* RefineAbsolutePoseFromLines (src/estimators/pose.cc:96-213, options pose.h:84-108) with constant
 * intrinsics (the defaults).  Returns PPSFM_OK when summary.IsSolutionUsable(), else
 * PPSFM_NO_SOLUTION.  The _ex variant takes AbsolutePoseRefinementOptions::refine_focal_length /
 * refine_extra_params (pose.cc:149-183) and updates camera_params in place. */
int ppsfm_refine_absolute_pose_from_lines(ppsfm_ctx* ctx, const uint8_t* inlier_mask,
                                          const double* lines, const double* points, size_t n,
                                          int camera_model, const double* camera_params,
                                          double gradient_tolerance, int max_num_iterations,
                                          double loss_function_scale, double* qvec, double* tvec,
                                          ppsfm_ba_summary* summary);
int ppsfm_refine_absolute_pose_from_lines_ex(ppsfm_ctx* ctx, const uint8_t* inlier_mask,
                                             const double* lines, const double* points, size_t n,
                                             int camera_model, double* camera_params,
                                             int refine_focal_length, int refine_extra_params,
                                             double gradient_tolerance, int max_num_iterations,
                                             double loss_function_scale, double* qvec,
                                             double* tvec, ppsfm_ba_summary* summary);

/* Test hooks (parity tests only).
 * ppsfm_ba_linearize: per observation residual[2], tangent Jacobians jac_cam[2x6] (rotation via
 * ceres::QuaternionParameterization | translation) and jac_point[2x3], loss-corrected, unscaled —
 * the counterpart of BundleAdjustmentLineCostFunction + AutoDiff (src/base/cost_functions.h:46-106).
 * ppsfm_dense_cholesky_solve: the reduced-camera-system solver on a host matrix. */
int ppsfm_ba_linearize(ppsfm_ctx* ctx, const ppsfm_ba_problem* problem,
                       const ppsfm_ba_options* options, double* residuals, double* jac_cam,
                       double* jac_point, double* cost);
int ppsfm_dense_cholesky_solve(ppsfm_ctx* ctx, const double* A, int n, const double* b, double* x);

/* =============================================================================================
 * Multi-GPU (one process / context per GPU).  The only data-path collective of the path is the
 * all-reduce of the reduced camera system in bundle adjustment: points are dealt
 * round-robin to ranks (p % world == rank, with all their observations), cameras are replicated,
 * every LM iteration all-reduces U, g_c, the bordered reduced matrix and four scalars over
 * NCCL / NVLink; the dense solve is replicated.  After ppsfm_comm_init every ppsfm_ba_* call on
 * that context is collective: all ranks pass the SAME problem and get the same result back.
 * The 128-byte id comes from rank 0 (ppsfm_comm_get_unique_id) and is distributed by the host
 * application (torch.distributed / MPI / a file).
 * ============================================================================================= */
int ppsfm_comm_get_unique_id(ppsfm_ctx* ctx, char* id128);
int ppsfm_comm_init(ppsfm_ctx* ctx, int world_size, int rank, const char* id128);
void ppsfm_comm_destroy(ppsfm_ctx* ctx);
int ppsfm_comm_rank(const ppsfm_ctx* ctx);
int ppsfm_comm_world_size(const ppsfm_ctx* ctx);
/* in-place sum all-reduce of a host buffer (staged through HBM); test / bench helper */
int ppsfm_comm_allreduce_sum_host(ppsfm_ctx* ctx, double* host, size_t count);
/* host-only shard accounting: out[4] = {kept observations on `rank`, owned points with
 * observations, camera blocks, kept observations in total} */
int ppsfm_ba_shard_stats(const ppsfm_ba_problem* problem, int rank, int world, int64_t* out);

/* =============================================================================================
 * Four-view initialisation from lifted lines with known gravity (host code, no GPU needed):
 * replaces init::initialize_reconstruction (src/init/initializer.h:103-108, initializer.cc:57-215)
 * and, underneath, FourView2dEstimator / PlanarOffsetEstimator (src/init/sfm2d.h:48-100,
 * src/init/initializer.h:62-101) run by ransac_lib::LocallyOptimizedMSAC
 * (lib/RansacLib/RansacLib/ransac.h:127-271).
 *   lines    [4][n][3]  (a, b, c) per image in normalised camera coordinates, same track order
 *   aligned  [4][n]     FeatureLine::IsAligned (the aligned / unaligned split must agree)
 *   gravity  [4][3]     gravity direction per image
 *   poses_out[4][12]    row-major 3x4 [R | t]
 * Returns PPSFM_OK, PPSFM_NO_SOLUTION (the reference's `false`) or PPSFM_ERR_INVALID where the
 * reference CHECK-aborts (initializer.cc:83, 99-104).
 * ============================================================================================= */
typedef struct ppsfm_init_options {   /* init::InitOptions, src/init/initializer.h:49-58 */
  double min_tri_angle;
  double min_num_inliers;
  double max_error;
} ppsfm_init_options;
typedef struct ppsfm_init_report {
  int32_t num_aligned, num_unaligned;
  int32_t inliers_2d, inliers_3d;
  uint32_t iterations_2d, iterations_3d;
  double mean_tri_angle_deg;
} ppsfm_init_report;
void ppsfm_init_options_default(ppsfm_init_options* options);
int ppsfm_initialize_reconstruction(const double* lines, const uint8_t* aligned, size_t n,
                                    const double* gravity, const ppsfm_init_options* options,
                                    double* poses_out, double* inlier_ratio,
                                    ppsfm_init_report* report);
/* The same run with the data-parallel part of both LO-MSAC loops on the GPU:
 * every candidate model of FourView2dEstimator (src/init/sfm2d.cc:302-444) and
 * PlanarOffsetEstimator (src/init/initializer.cc:219-333) is scored by one kernel launch per
 * minimal sample — all tracks triangulated and evaluated, MSAC sum in track order — with scores
 * bit-identical to the host's, so the result equals ppsfm_initialize_reconstruction's.
 * gpu_launches (optional): kernels launched. */
int ppsfm_initialize_reconstruction_gpu(ppsfm_ctx* ctx, const double* lines,
                                        const uint8_t* aligned, size_t n, const double* gravity,
                                        const ppsfm_init_options* options, double* poses_out,
                                        double* inlier_ratio, ppsfm_init_report* report,
                                        int64_t* gpu_launches);
/* Test hook: the generic and the fixed-size least-squares routine of the initialisation on one
 * m x n system ((3, 2) or (4, 3)); both must return the same bits. */
int ppsfm_init_test_qr(const double* A, int m, int n, const double* b, double* x_generic,
                       double* x_fixed);

/* =============================================================================================
 * Observation / point filters that follow every bundle adjustment, on a
 * track-major view of the reconstruction: the observations of point p are
 * [track_start[p], track_start[p+1]) in track order.  Replaces Reconstruction::FilterPoints3D
 * (src/base/reconstruction.cc:425-440 = FilterPoints3DWithLargeReprojectionError :650-719 followed
 * by FilterPoints3DWithSmallTriangulationAngle :594-648) and
 * Reconstruction::FilterObservationsWithNegativeDepth (:442-460).  Outputs are delete masks the
 * caller applies to its Reconstruction (DeleteObservation / DeletePoint3D) and the value the
 * reference returns (num_filtered); point_error [P] is in/out (Point3D::SetError for survivors:
 * error sum / REMAINING track length, :706-713).  The negative-depth filter follows
 * DeleteObservation's cascade (:255-275): a point whose track is down to <= 3 elements when one
 * of its observations is deleted goes away as a whole (point_deleted [P], may be NULL; all its
 * observations are flagged) and its later observations are not counted.
 * ============================================================================================= */
typedef struct ppsfm_filter_problem {
  int32_t num_images;
  const double* qvecs;           /* [num_images][4] (w, x, y, z) */
  const double* tvecs;           /* [num_images][3] */
  const int32_t* image_camera;   /* [num_images] */
  int32_t num_cameras;
  const int32_t* camera_model;   /* COLMAP model id (0..10) */
  const double* camera_params;   /* [num_cameras][12] */
  const int32_t* camera_width;   /* Camera::Width() */
  const int32_t* camera_height;
  int32_t num_points;
  const double* points;          /* [num_points][3] */
  const int64_t* track_start;    /* [num_points + 1] */
  int64_t num_obs;
  const int32_t* obs_image;      /* [num_obs] */
  const double* obs_line;        /* [num_obs][3], |(a, b)| = 1 */
  const uint8_t* obs_aligned;    /* [num_obs] FeatureLine::IsAligned */
} ppsfm_filter_problem;
int ppsfm_filter_points3d(ppsfm_ctx* ctx, const ppsfm_filter_problem* problem,
                          double max_reproj_error, double min_tri_angle_deg, uint8_t* obs_deleted,
                          uint8_t* point_deleted, double* point_error, size_t* num_filtered);
int ppsfm_filter_observations_with_negative_depth(ppsfm_ctx* ctx,
                                                  const ppsfm_filter_problem* problem,
                                                  uint8_t* obs_deleted, uint8_t* point_deleted,
                                                  size_t* num_filtered);

/* =============================================================================================
 * Batched robust line triangulation: EstimateTriangulation
 * (src/estimators/triangulation.h:143-147, triangulation.cc:118-149) for every track of a
 * track-major problem (ppsfm_filter_problem; `points` is not read, num_points = number of
 * tracks).  LORANSAC over the 3-combinations of a track in lexicographic order
 * (src/optim/loransac.h:91-234, combination_sampler.cc:41-70), multi-view point from lines
 * (src/base/triangulation.cc:41-57), squared angular or line-reprojection residuals.
 * exhaustive_threshold: tracks up to this length get min_num_trials = C(n, 3)
 * (src/sfm/incremental_triangulator.cc:527-531).  Outputs: xyz [T][3], success [T],
 * inlier_mask [O], num_trials [T] (may be NULL).
 * ============================================================================================= */
typedef struct ppsfm_triangulation_options {
  double min_tri_angle;          /* radians */
  int32_t residual_type;         /* 0 ANGULAR_ERROR (max_error in radians), 1 REPROJECTION_ERROR (pixels) */
  double max_error;              /* RANSACOptions */
  double min_inlier_ratio;
  double confidence;
  double dyn_num_trials_multiplier;
  uint64_t min_num_trials;
  uint64_t max_num_trials;
  int32_t exhaustive_threshold;
} ppsfm_triangulation_options;
void ppsfm_triangulation_options_default(ppsfm_triangulation_options* options);
int ppsfm_estimate_triangulation_batch(ppsfm_ctx* ctx, const ppsfm_filter_problem* tracks,
                                       const ppsfm_triangulation_options* options, double* xyz,
                                       uint8_t* success, uint8_t* inlier_mask,
                                       uint32_t* num_trials);

/* CameraModel::ImageToWorldThreshold (src/base/camera_models.h:533-543): pixel threshold / mean
 * focal length of the model (host arithmetic).  PPSFM_ERR_INVALID for an unsupported model id. */
int ppsfm_image_to_world_threshold(int camera_model, const double* camera_params, double threshold,
                                   double* out);
/* RotationMatrixToQuaternion (src/base/pose.cc:41-44 = Eigen::Quaterniond(R)); R column-major
 * 3x3, qvec (w, x, y, z).  Host arithmetic. */
void ppsfm_rotation_matrix_to_quaternion(const double* R, double* qvec);

/* ---- measurement helpers (bench.py only; not part of the reference surface) ------------------
 * FP64 issue rate in 1e12 thread-instructions/s: fused (DFMA) and unfused (DMUL/DADD mix). */
int ppsfm_bench_fp64_peak(ppsfm_ctx* ctx, double* dfma_tips, double* dmuladd_tips);
/* Packed float FMA (FFMA2) rate in 1e12 FMAs/s: the ceiling of the score filter's float stage. */
int ppsfm_bench_fp32_peak(ppsfm_ctx* ctx, double* ffma_tips);
/* Write-only / read-only HBM bandwidth (GB/s) over a 2 GiB buffer. */
int ppsfm_bench_hbm_rw_peak(ppsfm_ctx* ctx, double* write_gbs, double* read_gbs);
/* Evict L2 by writing `bytes` of scratch HBM (blocking). */
int ppsfm_bench_l2_flush(ppsfm_ctx* ctx, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* PPSFM_B200_H_ */
