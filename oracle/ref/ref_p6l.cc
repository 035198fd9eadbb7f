// ref_p6l.cc — TEST INFRASTRUCTURE.  C entry points around the REFERENCE's own sources of the
// absolute-pose RANSAC path, compiled from where they lie under /root/reference (never copied):
//   src/estimators/absolute_pose.cc   P6LEstimator::Estimate / Residuals        
//   lib/re3q3/re3q3/re3q3.h           re3q3                                     (A4)
//   src/estimators/utils.cc           ComputeSquaredLineReprojectionError       (A5)
//   src/optim/ransac.h                RANSAC<Estimator, SupportMeasurer, Sampler>::Estimate (A6)
//   src/optim/random_sampler.cc, src/util/random.{h,cc}   sampler + PRNG        (A7)
//   src/optim/support_measurement.cc  Inlier / MEstimator support               (A8)
//   src/estimators/pose.cc            EstimateAbsolutePoseFromLines             (A9)
//   src/base/pose.cc                  RotationMatrixToQuaternion, NormalizeQuaternion
//                                     RefineAbsolutePoseFromLines: the problem it builds (A10)
// Eigen and glog are absent in this image: the sources compile against the stand-ins under
// oracle/ref/shim/ (minieigen.h: the Eigen calls are the restatements of oracle/eigen_restated.h,
// shared with the oracle; glog/logging.h: CHECK* = print + abort).  So what tests/test_ref_p6l.py
// pins is the reference's SOURCE TEXT against the oracle's restatement of it, bit for bit; the
// inside of Eigen (determinant, PartialPivLU, EigenSolver, product association) stays unpinned.
// Built by oracle/build_ref.sh into oracle/_ref/libref_p6l.so.  Signatures mirror the orc_*
// functions of oracle/ppsfm_oracle.h.
#include <cstdint>
#include <cstring>
#include <random>
#include <vector>

#include "base/camera.h"
#include "base/pose.h"
#include "estimators/absolute_pose.h"
#include "estimators/pose.h"
#include "estimators/utils.h"
#include "feature/types.h"
#include "optim/random_sampler.h"
#include "optim/ransac.h"
#include "optim/support_measurement.h"
#include "util/random.h"

// defined (non-inline) by lib/re3q3/re3q3/re3q3.h inside the absolute_pose.cc translation unit
int re3q3(Eigen::Matrix<double, 3, 10> coeffs, Eigen::Matrix<double, 3, 8>* solutions,
          bool try_random_var_change);

#define REF_API extern "C" __attribute__((visibility("default")))

namespace {

struct Options {  // == orc_ransac_options
  double max_error, min_inlier_ratio, confidence, dyn_num_trials_multiplier;
  uint64_t min_num_trials, max_num_trials;
};
struct Report {  // == orc_ransac_report
  int32_t success;
  uint64_t num_trials, num_inliers;
  double residual_sum;
  double model[12];
  int64_t best_trial;
  int32_t best_model_idx;
  uint64_t num_models_scored;
};

void Gather(const double* lines, const uint8_t* aligned, const double* points, size_t n,
            colmap::FeatureLines* X, std::vector<Eigen::Vector3d>* Y) {
  X->clear();
  Y->clear();
  for (size_t i = 0; i < n; ++i) {
    X->emplace_back(Eigen::Vector3d(lines[3 * i], lines[3 * i + 1], lines[3 * i + 2]),
                    aligned != nullptr && aligned[i] != 0);
    Y->emplace_back(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
  }
}

}  // namespace


REF_API void ref_set_prng_seed(uint32_t seed) { colmap::SetPRNGSeed(seed); }

// the next 32-bit output of the generator, without consuming it (== orc_prng_peek)
REF_API uint32_t ref_prng_peek(void) {
  if (colmap::PRNG == nullptr) colmap::SetPRNGSeed();
  std::mt19937 copy = *colmap::PRNG;
  return static_cast<uint32_t>(copy());
}

REF_API void ref_line_residuals(const double* lines, const double* points, size_t n, const double* model,
                        double* residuals_out) {
  std::vector<Eigen::Vector3d> l, p;
  for (size_t i = 0; i < n; ++i) {
    l.emplace_back(lines[3 * i], lines[3 * i + 1], lines[3 * i + 2]);
    p.emplace_back(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
  }
  Eigen::Matrix3x4d m;
  std::memcpy(m.data(), model, sizeof(double) * 12);  // column-major 3x4, like Eigen's storage
  std::vector<double> r;
  colmap::ComputeSquaredLineReprojectionError(l, p, m, &r);
  std::memcpy(residuals_out, r.data(), sizeof(double) * n);
}

REF_API void ref_inlier_support(const double* residuals, size_t n, double max_residual,
                        uint64_t* num_inliers, double* residual_sum) {
  colmap::InlierSupportMeasurer m;
  const auto s = m.Evaluate(std::vector<double>(residuals, residuals + n), max_residual);
  *num_inliers = s.num_inliers;
  *residual_sum = s.residual_sum;
}

// 1 if support (n1, s1) is better than (n2, s2): InlierSupportMeasurer::Compare
REF_API int ref_inlier_support_compare(uint64_t n1, double s1, uint64_t n2, double s2) {
  colmap::InlierSupportMeasurer m;
  colmap::InlierSupportMeasurer::Support a, b;
  a.num_inliers = n1;
  a.residual_sum = s1;
  b.num_inliers = n2;
  b.residual_sum = s2;
  return m.Compare(a, b) ? 1 : 0;
}

REF_API void ref_mestimator_support(const double* residuals, size_t n, double max_residual,
                            uint64_t* num_inliers, double* score) {
  colmap::MEstimatorSupportMeasurer m;
  const auto s = m.Evaluate(std::vector<double>(residuals, residuals + n), max_residual);
  *num_inliers = s.num_inliers;
  *score = s.score;
}

REF_API uint64_t ref_compute_num_trials(uint64_t num_inliers, uint64_t num_samples, double confidence,
                                double num_trials_multiplier) {
  return colmap::RANSAC<colmap::P6LEstimator>::ComputeNumTrials(num_inliers, num_samples,
                                                                 confidence, num_trials_multiplier);
}

REF_API void ref_sample_table(size_t n, size_t num_trials, uint32_t* table_out) {
  colmap::RandomSampler sampler(colmap::P6LEstimator::kMinNumSamples);
  sampler.Initialize(n);
  for (size_t t = 0; t < num_trials; ++t) {
    const std::vector<size_t> idx = sampler.Sample();
    for (int i = 0; i < 6; ++i) table_out[6 * t + i] = static_cast<uint32_t>(idx[i]);
  }
}

// coeffs: 3 x 10 row-major; solutions: 8 x 3 (solution-major) like orc_re3q3
REF_API int ref_re3q3(const double* coeffs, double* solutions) {
  Eigen::Matrix<double, 3, 10> c;
  for (int k = 0; k < 3; ++k)
    for (int j = 0; j < 10; ++j) c(k, j) = coeffs[10 * k + j];
  Eigen::Matrix<double, 3, 8> s;
  for (int k = 0; k < 8; ++k)
    for (int i = 0; i < 3; ++i) s(i, k) = 0.0;
  const int n = re3q3(c, &s, true);
  for (int k = 0; k < 8; ++k)
    for (int i = 0; i < 3; ++i) solutions[3 * k + i] = s(i, k);
  return n;
}

// lines6 / points6: 6 x 3 row-major; models_out: up to 8 column-major 3x4 matrices
REF_API int ref_p6l_estimate(const double* lines6, const uint8_t* aligned6, const double* points6,
                     double* models_out) {
  colmap::FeatureLines X;
  std::vector<Eigen::Vector3d> Y;
  Gather(lines6, aligned6, points6, 6, &X, &Y);
  const std::vector<colmap::P6LEstimator::M_t> models = colmap::P6LEstimator::Estimate(X, Y);
  for (size_t k = 0; k < models.size(); ++k)
    std::memcpy(models_out + 12 * k, models[k].data(), sizeof(double) * 12);
  return static_cast<int>(models.size());
}

// colmap::RANSAC<P6LEstimator>(options).Estimate(X, Y) — the serial loop the GPU path replaces.
// best_trial / best_model_idx / num_models_scored are not observable from outside the loop: -1 / 0.
REF_API void ref_ransac_p6l(const double* lines, const uint8_t* aligned, const double* points, size_t n,
                    const Options* options, Report* report, uint8_t* inlier_mask) {
  colmap::RANSACOptions o;
  o.max_error = options->max_error;
  o.min_inlier_ratio = options->min_inlier_ratio;
  o.confidence = options->confidence;
  o.dyn_num_trials_multiplier = options->dyn_num_trials_multiplier;
  o.min_num_trials = options->min_num_trials;
  o.max_num_trials = options->max_num_trials;
  colmap::FeatureLines X;
  std::vector<Eigen::Vector3d> Y;
  Gather(lines, aligned, points, n, &X, &Y);
  colmap::RANSAC<colmap::P6LEstimator> ransac(o);
  const auto r = ransac.Estimate(X, Y);
  std::memset(report, 0, sizeof(*report));
  report->success = r.success ? 1 : 0;
  report->num_trials = r.num_trials;
  report->num_inliers = r.support.num_inliers;
  report->residual_sum = r.support.residual_sum;
  std::memcpy(report->model, r.model.data(), sizeof(double) * 12);
  report->best_trial = -1;
  report->best_model_idx = -1;
  if (inlier_mask != nullptr) {
    std::memset(inlier_mask, 0, n);
    for (size_t i = 0; i < r.inlier_mask.size(); ++i) inlier_mask[i] = r.inlier_mask[i] ? 1 : 0;
  }
}


// colmap::EstimateAbsolutePoseFromLines (src/estimators/pose.cc:52-94): the RANSAC call, the
// rejection of mostly-aligned inlier sets, the quaternion conversion and the NaN check.
// Returns the function's bool; inlier_mask is what the function leaves in *inlier_mask.
REF_API int ref_estimate_absolute_pose_from_lines(const double* lines, const uint8_t* aligned,
                                                  const double* points, size_t n,
                                                  const Options* options, double* qvec,
                                                  double* tvec, uint64_t* num_inliers,
                                                  uint8_t* inlier_mask) {
  colmap::RANSACOptions o;
  o.max_error = options->max_error;
  o.min_inlier_ratio = options->min_inlier_ratio;
  o.confidence = options->confidence;
  o.dyn_num_trials_multiplier = options->dyn_num_trials_multiplier;
  o.min_num_trials = options->min_num_trials;
  o.max_num_trials = options->max_num_trials;
  colmap::FeatureLines X;
  std::vector<Eigen::Vector3d> Y;
  Gather(lines, aligned, points, n, &X, &Y);
  Eigen::Vector4d q(0, 0, 0, 0);
  Eigen::Vector3d t(0, 0, 0);
  size_t ninl = 0;
  std::vector<char> mask;
  const bool ok = colmap::EstimateAbsolutePoseFromLines(o, X, Y, &q, &t, &ninl, &mask);
  for (int k = 0; k < 4; ++k) qvec[k] = q(k);
  for (int k = 0; k < 3; ++k) tvec[k] = t(k);
  *num_inliers = ninl;
  std::memset(inlier_mask, 0, n);
  for (size_t i = 0; i < mask.size(); ++i) inlier_mask[i] = mask[i] ? 1 : 0;
  return ok ? 1 : 0;
}

// colmap::RefineAbsolutePoseFromLines (src/estimators/pose.cc:96-213) against the RECORDING
// ceres::Problem of the stand-in: what the reference hands to Ceres for a pose refinement.
// out[0] residual blocks, [1] all blocks are (2; 4, 3, 3, k) on the same qvec / tvec / camera
// pointers, [2] loss kind (2 = Cauchy) and [3] its scale, [4] constant 3-blocks (the points),
// [5] the residual blocks' points are exactly the inlier points in order, [6] qvec has a 4 -> 3
// parameterisation, [7] tvec is neither constant nor parameterised, [8] variable mask of the
// camera parameters, [9] linear solver type (1 = DENSE_QR), [10] gradient_tolerance,
// [11] max_num_iterations, [12] num_threads, [13..16] qvec after the call (normalised in place).
REF_API int ref_refine_absolute_pose_setup(const double* lines, const double* points,
                                           const uint8_t* inlier_mask, size_t n, int camera_model,
                                           const double* camera_params, int refine_focal_length,
                                           int refine_extra_params, double gradient_tolerance,
                                           int max_num_iterations, double loss_scale,
                                           const double* qvec_in, const double* tvec_in,
                                           double* out) {
  colmap::AbsolutePoseRefinementOptions o;
  o.gradient_tolerance = gradient_tolerance;
  o.max_num_iterations = max_num_iterations;
  o.loss_function_scale = loss_scale;
  o.refine_focal_length = refine_focal_length != 0;
  o.refine_extra_params = refine_extra_params != 0;
  o.print_summary = false;
  colmap::Camera camera;
  camera.SetModelId(camera_model);
  camera.SetWidth(1000);
  camera.SetHeight(1000);
  camera.SetParams(std::vector<double>(camera_params, camera_params + camera.NumParams()));
  std::vector<Eigen::Vector3d> l, p;
  std::vector<char> mask(inlier_mask, inlier_mask + n);
  for (size_t i = 0; i < n; ++i) {
    l.emplace_back(lines[3 * i], lines[3 * i + 1], lines[3 * i + 2]);
    p.emplace_back(points[3 * i], points[3 * i + 1], points[3 * i + 2]);
  }
  Eigen::Vector4d q(qvec_in[0], qvec_in[1], qvec_in[2], qvec_in[3]);
  Eigen::Vector3d t(tvec_in[0], tvec_in[1], tvec_in[2]);
  const bool usable = colmap::RefineAbsolutePoseFromLines(o, mask, l, p, &q, &t, &camera);

  const ceres::SolveRecord& r = ceres::LastSolveRecord();
  out[0] = r.num_residual_blocks;
  bool uniform = true, points_in_order = true;
  size_t next = 0;
  for (int b = 0; b < r.num_residual_blocks; ++b) {
    const auto& sz = r.block_sizes[b];
    uniform = uniform && sz.size() == 4 && sz[0] == 4 && sz[1] == 3 && sz[2] == 3 &&
              sz[3] == static_cast<int>(camera.NumParams()) && r.blocks[b][0] == q.data() &&
              r.blocks[b][1] == t.data() && r.blocks[b][3] == camera.ParamsData() &&
              r.loss_kind[b] == r.loss_kind[0] && r.loss_scale[b] == r.loss_scale[0];
    while (next < n && !mask[next]) ++next;
    // block3_values: tvec (3) then the point (3) (then the camera if it has three parameters)
    points_in_order = points_in_order && next < n && r.block3_values[b].size() >= 6 &&
                      r.block3_values[b][3] == p[next](0) && r.block3_values[b][4] == p[next](1) &&
                      r.block3_values[b][5] == p[next](2);
    ++next;
  }
  out[1] = uniform ? 1 : 0;
  out[2] = r.num_residual_blocks ? r.loss_kind[0] : -1;
  out[3] = r.num_residual_blocks ? r.loss_scale[0] : 0.0;
  int const_points = 0;
  bool camera_const = false, tvec_touched = false, quat = false;
  for (const double* c : r.constant_blocks) {
    if (c == camera.ParamsData()) camera_const = true;
    else if (c == t.data() || c == q.data()) tvec_touched = true;
    else ++const_points;
  }
  unsigned variable = camera_const ? 0u : (1u << camera.NumParams()) - 1;
  for (const auto& pr : r.parameterizations) {
    if (pr.block == q.data()) quat = pr.global_size == 4 && pr.local_size == 3;
    else if (pr.block == t.data()) tvec_touched = true;
    else if (pr.block == camera.ParamsData())
      for (int k : pr.constant) variable &= ~(1u << k);
  }
  out[4] = const_points;
  out[5] = points_in_order ? 1 : 0;
  out[6] = quat ? 1 : 0;
  out[7] = tvec_touched ? 0 : 1;
  out[8] = variable;
  out[9] = static_cast<int>(r.options.linear_solver_type);
  out[10] = r.options.gradient_tolerance;
  out[11] = r.options.max_num_iterations;
  out[12] = r.options.num_threads;
  for (int k = 0; k < 4; ++k) out[13 + k] = q(k);
  return usable ? 1 : 0;
}
