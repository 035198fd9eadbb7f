// estimators_triangulation.cc — the drop-in for the reference's robust line triangulation.
//
// Compiled inside the reference tree in place of the body of
//     bool EstimateTriangulation(...)              src/estimators/triangulation.cc:117-149
// (declared in src/estimators/triangulation.h:143-147, included here, so the compiler checks the
// signature and the PointData / PoseData / options types against the reference's own), with
// -DPPSFM_WITH_EIGEN, linking -lppsfm_b200.  Callers: IncrementalTriangulator::Create / Continue
// (src/sfm/incremental_triangulator.cc:468-561).  They set min_num_trials = NChooseK(n, 3) for
// short tracks themselves (:527-531), so exhaustive_threshold stays 0 here.  One call is one
// track; a triangulator that wants the GPU's throughput collects the tracks of an image and
// calls ppsfm::EstimateTriangulationBatch once.
#include "estimators/triangulation.h"  // the reference's declarations

#ifndef PPSFM_WITH_EIGEN
#define PPSFM_WITH_EIGEN
#endif
#include "ppsfm_adaptor.h"

namespace colmap {

bool EstimateTriangulation(const EstimateTriangulationOptions& options,
                           const std::vector<TriangulationEstimator::PointData>& point_data,
                           const std::vector<TriangulationEstimator::PoseData>& pose_data,
                           std::vector<char>* inlier_mask, Eigen::Vector3d* xyz) {
  typedef ppsfm::TriangulationEstimator PE;
  ppsfm::EstimateTriangulationOptions o;
  o.min_tri_angle = options.min_tri_angle;
  o.residual_type = options.residual_type == TriangulationEstimator::ResidualType::ANGULAR_ERROR
                        ? PE::ResidualType::ANGULAR_ERROR
                        : PE::ResidualType::REPROJECTION_ERROR;
  o.ransac_options.max_error = options.ransac_options.max_error;
  o.ransac_options.min_inlier_ratio = options.ransac_options.min_inlier_ratio;
  o.ransac_options.confidence = options.ransac_options.confidence;
  o.ransac_options.dyn_num_trials_multiplier = options.ransac_options.dyn_num_trials_multiplier;
  o.ransac_options.min_num_trials = options.ransac_options.min_num_trials;
  o.ransac_options.max_num_trials = options.ransac_options.max_num_trials;
  o.exhaustive_threshold = 0;
  std::vector<PE::PointData> points(point_data.size());
  std::vector<PE::PoseData<Camera>> poses(pose_data.size());
  for (size_t i = 0; i < point_data.size(); ++i) points[i] = PE::PointData(point_data[i].line);
  for (size_t i = 0; i < pose_data.size(); ++i)
    poses[i] = PE::PoseData<Camera>(pose_data[i].proj_matrix, pose_data[i].proj_center,
                                    pose_data[i].camera);
  return ppsfm::EstimateTriangulation<Camera>(o, points, poses, inlier_mask, xyz);
}

}  // namespace colmap
