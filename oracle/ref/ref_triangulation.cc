// ref_triangulation.cc — TEST INFRASTRUCTURE.  The REFERENCE's own robust line triangulation
// (SURVEY §8 row f1), compiled from where it lies under /root/reference:
//   src/estimators/triangulation.cc    EstimateTriangulation, TriangulationEstimator::Estimate /
//                                      Residuals (cheirality, triangulation angle)
//   src/base/triangulation.cc          TriangulateMultiViewPoint, CalculateTriangulationAngle
//   src/base/projection.cc             CalculateNormalizedLineAngularError,
//                                      CalculateSquaredLineReprojectionError, HasPointPositiveDepth
//   src/base/camera.cc, camera_models.cc   Camera::WorldToImage
//   src/optim/loransac.h, ransac.h     LORANSAC<>::Estimate (local optimisation, abort rule)
//   src/optim/combination_sampler.cc, src/util/math.{h,cc}, src/optim/support_measurement.cc
// against the stand-ins of oracle/ref/shim/.  The one piece that is NOT the reference's text is
// the n x 4 JacobiSVD inside TriangulateMultiViewPoint: the stand-in provides the null vector by
// eigen_restated::NullVectorNx4, the iteration the oracle uses.  This translation unit plays the
// caller (IncrementalTriangulator::Create, src/sfm/incremental_triangulator.cc:468-561): it builds
// PointData / PoseData per track — projection matrix and centre exactly as the oracle's MakePose
// builds them — and applies the exhaustive-sampling rule (:527-531).
// Built by oracle/build_ref.sh into oracle/_ref/libref_tri.so.
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "base/camera.h"
#include "estimators/triangulation.h"
#include "optim/combination_sampler.h"
#include "optim/loransac.h"
#include "util/math.h"

namespace {

struct Problem {  // same layout as ppsfm_filter_problem
  int32_t num_images; const double* qvecs; const double* tvecs; const int32_t* image_camera;
  int32_t num_cameras; const int32_t* camera_model; const double* camera_params;
  const int32_t* camera_width; const int32_t* camera_height;
  int32_t num_points; const double* points; const int64_t* track_start;
  int64_t num_obs; const int32_t* obs_image; const double* obs_line; const uint8_t* obs_aligned;
};
struct Options {  // same layout as ppsfm_triangulation_options
  double min_tri_angle; int32_t residual_type; double max_error, min_inlier_ratio, confidence,
      dyn_num_trials_multiplier; uint64_t min_num_trials, max_num_trials; int32_t exhaustive_threshold;
};

// [R | t] and the projection centre -R^T t of a (w, x, y, z) quaternion + translation: the
// caller's Image::ProjectionMatrix() / ProjectionCenter(), written like the oracle's MakePose
void MakePose(const double* qv, const double* t, Eigen::Matrix3x4d* P, Eigen::Vector3d* c) {
  const double n = std::sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2] + qv[3] * qv[3]);
  const Eigen::Matrix3d R = Eigen::Quaterniond(qv[0] / n, qv[1] / n, qv[2] / n, qv[3] / n).toRotationMatrix();
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) (*P)(r, k) = R(r, k);
    (*P)(r, 3) = t[r];
  }
  for (int k = 0; k < 3; ++k) (*c)(k) = -(R(0, k) * t[0] + R(1, k) * t[1] + R(2, k) * t[2]);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int ref_estimate_triangulation_batch(
    const Problem* pbp, const Options* optp, double* xyz, uint8_t* success, uint8_t* inlier_mask,
    uint32_t* num_trials) {
  const Problem& pb = *pbp;
  const Options& opt = *optp;
  std::vector<colmap::Camera> cameras(pb.num_cameras);
  for (int c = 0; c < pb.num_cameras; ++c) {
    cameras[c].SetCameraId(c + 1);
    cameras[c].SetModelId(pb.camera_model[c]);
    cameras[c].SetWidth(pb.camera_width[c]);
    cameras[c].SetHeight(pb.camera_height[c]);
    cameras[c].SetParams(std::vector<double>(pb.camera_params + 12 * c,
                                             pb.camera_params + 12 * c + cameras[c].NumParams()));
  }
  typedef colmap::TriangulationEstimator TE;
  for (int t = 0; t < pb.num_points; ++t) {
    const int64_t k0 = pb.track_start[t], k1 = pb.track_start[t + 1];
    const size_t n = static_cast<size_t>(k1 - k0);
    success[t] = 0;
    if (num_trials) num_trials[t] = 0;
    for (int64_t k = k0; k < k1; ++k) inlier_mask[k] = 0;
    if (n < 2) continue;  // (EstimateTriangulation CHECK-aborts below two views)
    std::vector<TE::PointData> point_data;
    std::vector<TE::PoseData> pose_data;
    for (int64_t k = k0; k < k1; ++k) {
      const int img = pb.obs_image[k];
      Eigen::Matrix3x4d P;
      Eigen::Vector3d c;
      MakePose(pb.qvecs + 4 * (size_t)img, pb.tvecs + 3 * (size_t)img, &P, &c);
      const double* l = pb.obs_line + 3 * (size_t)k;
      point_data.emplace_back(Eigen::Vector3d(l[0], l[1], l[2]));
      pose_data.emplace_back(P, c, &cameras[pb.image_camera[img]]);
    }
    colmap::EstimateTriangulationOptions o;
    o.min_tri_angle = opt.min_tri_angle;
    o.residual_type = opt.residual_type == 0 ? TE::ResidualType::ANGULAR_ERROR
                                             : TE::ResidualType::REPROJECTION_ERROR;
    o.ransac_options.max_error = opt.max_error;
    o.ransac_options.min_inlier_ratio = opt.min_inlier_ratio;
    o.ransac_options.confidence = opt.confidence;
    o.ransac_options.dyn_num_trials_multiplier = opt.dyn_num_trials_multiplier;
    o.ransac_options.min_num_trials = opt.min_num_trials;
    o.ransac_options.max_num_trials = opt.max_num_trials;
    // incremental_triangulator.cc:527-531: exhaustive sampling for short tracks
    if (static_cast<int>(n) <= opt.exhaustive_threshold)
      o.ransac_options.min_num_trials = colmap::NChooseK(n, 3);

    std::vector<char> mask;
    Eigen::Vector3d X(0, 0, 0);
    const bool ok = colmap::EstimateTriangulation(o, point_data, pose_data, &mask, &X);

    // the trial count is not part of EstimateTriangulation's interface: run the same LORANSAC
    // again the way triangulation.cc:133-141 sets it up, and require the same outcome
    if (n >= 3) {
      colmap::LORANSAC<TE, TE, colmap::InlierSupportMeasurer, colmap::CombinationSampler> ransac(
          o.ransac_options);
      ransac.estimator.SetMinTriAngle(o.min_tri_angle);
      ransac.estimator.SetResidualType(o.residual_type);
      ransac.local_estimator.SetMinTriAngle(o.min_tri_angle);
      ransac.local_estimator.SetResidualType(o.residual_type);
      const auto report = ransac.Estimate(point_data, pose_data);
      if (report.success != ok) std::abort();
      if (ok)
        for (int c = 0; c < 3; ++c)
          if (!(report.model(c) == X(c))) std::abort();
      if (num_trials) num_trials[t] = static_cast<uint32_t>(report.num_trials);
    }
    if (!ok) continue;
    success[t] = 1;
    for (int c = 0; c < 3; ++c) xyz[3 * (size_t)t + c] = X(c);
    for (size_t i = 0; i < n; ++i) inlier_mask[k0 + i] = mask[i] ? 1 : 0;
  }
  return 0;
}
