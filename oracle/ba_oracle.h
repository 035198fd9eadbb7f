/* ba_oracle.h — CPU oracle of the line-reprojection bundle adjustment.
 *
 * TEST INFRASTRUCTURE ONLY (see ppsfm_oracle.h).
 *
 * PARITY UNPINNED at the Ceres boundary: the reference hands the problem to ceres::Solve
 * (src/optim/bundle_adjustment.cc:306, src/estimators/pose.cc:201) and Ceres is neither in
 * /root/reference nor installed; no reference test constructs a BundleAdjuster or the line cost
 * functors.  PINNED against the reference's own sources: the two cost functors and the camera
 * models — src/base/cost_functions.h and src/base/camera_models.{h,cc} compiled from
 * /root/reference against stand-ins (oracle/build_ref.sh -> oracle/_ref/libref_cost.so) give
 * bit-identical residuals and Jacobian blocks (tests/test_ref_cost.py).  This oracle restates
 *   - the residual functors of src/base/cost_functions.h:46-191 evaluated on forward-mode dual
 *     numbers (what ceres::AutoDiffCostFunction does),
 *   - the camera models of src/base/camera_models.h (all 11),
 *   - problem assembly / gauge rules of src/optim/bundle_adjustment.cc:326-542,
 *   - Ceres' documented trust-region Levenberg-Marquardt with Jacobi scaling, loss-function
 *     corrector and Schur elimination of the points,
 * and is cross-checked against scipy.optimize.least_squares in tests/.
 *
 * The struct layouts are identical to include/ppsfm_b200.h (ppsfm_ba_problem / _options /
 * _summary) so the same numpy buffers can be handed to both.
 */
#ifndef PPSFM_BA_ORACLE_H_
#define PPSFM_BA_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_BA_MAX_TRACE 128

typedef struct {
  int32_t num_images;
  double* qvecs;               /* num_images x 4 (w,x,y,z), in/out */
  double* tvecs;               /* num_images x 3, in/out */
  const uint8_t* pose_flags;   /* bit0: pose constant; bit1..3: tvec[0..2] constant */
  const int32_t* image_camera; /* index into cameras */
  int32_t num_cameras;
  const int32_t* camera_model; /* COLMAP model ids (camera_models.h:189-248) */
  double* camera_params;       /* num_cameras x 12, zero padded; in/out when intrinsics are refined */
  int32_t num_points;
  double* points;              /* num_points x 3, in/out */
  const uint8_t* point_const;  /* may be NULL */
  int64_t num_obs;
  const int32_t* obs_image;
  const int32_t* obs_point;
  const double* obs_line;      /* num_obs x 3 */
  const uint8_t* camera_const; /* config.IsConstantCamera per camera (bundle_adjustment.cc:497);
                                  may be NULL (= none) */
} orc_ba_problem;

typedef struct {
  int32_t loss_type; /* 0 TRIVIAL, 1 SOFT_L1, 2 CAUCHY (bundle_adjustment.h:51) */
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
  int32_t num_threads;
  /* BundleAdjustmentOptions::refine_* (bundle_adjustment.h:57-63) -> ParameterizeCameras
   * (bundle_adjustment.cc:490-528): which groups of Camera::Params() are variable */
  int32_t refine_focal_length;
  int32_t refine_principal_point;
  int32_t refine_extra_params;
} orc_ba_options;

typedef struct {
  double initial_cost;
  double final_cost;
  int32_t num_successful_steps;
  int32_t num_unsuccessful_steps;
  int32_t termination_type; /* 0 CONVERGENCE, 1 NO_CONVERGENCE (iteration cap), 2 FAILURE */
  int64_t num_residuals;         /* 2 x observations handed in */
  int64_t num_residuals_reduced; /* after dropping all-constant residual blocks */
  int32_t num_effective_parameters_reduced;
  double total_time_s;
  double jacobian_time_s;
  double linear_solver_time_s;
  double final_gradient_max_norm;
  int32_t trace_len;
  double trace_cost[ORC_BA_MAX_TRACE];     /* cost after each iteration (iteration 0 = initial) */
  double trace_radius[ORC_BA_MAX_TRACE];
  int32_t trace_accepted[ORC_BA_MAX_TRACE];
} orc_ba_summary;

void orc_ba_options_default(orc_ba_options* opt);
/* BundleAdjuster::Solve on the SoA problem; updates qvecs / tvecs / points in place.
 * Returns 1 (true) or 0 (no residuals, bundle_adjustment.cc:269-271). */
int orc_ba_solve(const orc_ba_problem* problem, const orc_ba_options* options,
                 orc_ba_summary* summary);

/* One residual block (cost_functions.h:62-100) with its Jacobians as ceres::AutoDiff would return
 * them: residual[2], jac_q[2x4], jac_t[2x3], jac_X[2x3] (row-major). */
void orc_line_cost(int camera_model, const double* camera_params, const double* line,
                   const double* qvec, const double* tvec, const double* point, double* residual,
                   double* jac_q, double* jac_t, double* jac_X);
/* The block (2; 4, 3, 3, kNumParams) of intrinsics refinement: additionally jac_camera[2x12]
 * (row-major, columns >= kNumParams zero). */
void orc_line_cost_intr(int camera_model, const double* camera_params, const double* line,
                        const double* qvec, const double* tvec, const double* point,
                        double* residual, double* jac_q, double* jac_t, double* jac_X,
                        double* jac_camera);
/* Same block in the tangent space used by the solver: jac_cam[2x6] = [rotation(3) via
 * QuaternionParameterization | translation(3)], jac_X[2x3]. */
void orc_line_cost_tangent(int camera_model, const double* camera_params, const double* line,
                           const double* qvec, const double* tvec, const double* point,
                           double* residual, double* jac_cam, double* jac_X);
/* 0.5 * sum rho(|r|^2) of the problem at its current state. */
double orc_ba_cost(const orc_ba_problem* problem, const orc_ba_options* options);
/* ceres::QuaternionParameterization::Plus */
void orc_quaternion_plus(const double* q, const double* delta, double* q_plus);

/* RefineAbsolutePoseFromLines (src/estimators/pose.cc:96-213) with constant intrinsics:
 * Cauchy loss, 6-dof LM over the inlier correspondences. Returns IsSolutionUsable(). */
int orc_refine_absolute_pose(const double* lines, const double* points, const uint8_t* inlier_mask,
                             size_t n, int camera_model, const double* camera_params,
                             double gradient_tolerance, int max_num_iterations, double loss_scale,
                             double* qvec, double* tvec, orc_ba_summary* summary);

#ifdef __cplusplus
}
#endif
#endif
