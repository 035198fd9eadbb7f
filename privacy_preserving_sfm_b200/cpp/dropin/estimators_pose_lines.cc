// estimators_pose_lines.cc — the drop-in for the reference's line-pose entry points.
//
// A maintainer of colmap/privacy_preserving_sfm compiles THIS file inside the reference tree in
// place of the bodies of
//     bool EstimateAbsolutePoseFromLines(...)      src/estimators/pose.cc:52-94
//     bool RefineAbsolutePoseFromLines(...)        src/estimators/pose.cc:96-213
// (declared in src/estimators/pose.h:110-122, which this file includes, so the compiler checks the
// signatures against the reference's own declarations) with -DPPSFM_WITH_EIGEN and links
// -lppsfm_b200.  The callers — IncrementalMapper::RegisterNextImage,
// src/sfm/incremental_mapper.cc:719-735 — stay untouched: same types (colmap::RANSACOptions,
// colmap::FeatureLines, colmap::AbsolutePoseRefinementOptions, colmap::Camera), same return
// values, same in-place outputs.  There is no CPU fallback: without the CUDA library the link
// fails, without a device the first call aborts (PPSFM_CHECK, like glog CHECK).
//
// tests/test_dropin.py compiles this file against the reference's headers where /root/reference
// exists (CPU), and runs it on the GPU against the reference's own RANSAC loop.
#include "estimators/pose.h"  // the reference's declarations

#ifndef PPSFM_WITH_EIGEN
#define PPSFM_WITH_EIGEN
#endif
#include "ppsfm_adaptor.h"

namespace colmap {

bool EstimateAbsolutePoseFromLines(const RANSACOptions& options, const FeatureLines& lines2D,
                                   const std::vector<Eigen::Vector3d>& points3D,
                                   Eigen::Vector4d* qvec, Eigen::Vector3d* tvec,
                                   size_t* num_inliers, std::vector<char>* inlier_mask) {
  ppsfm::RANSACOptions o;  // same fields, same defaults (src/optim/ransac.h:47-76)
  o.max_error = options.max_error;
  o.min_inlier_ratio = options.min_inlier_ratio;
  o.confidence = options.confidence;
  o.dyn_num_trials_multiplier = options.dyn_num_trials_multiplier;
  o.min_num_trials = options.min_num_trials;
  o.max_num_trials = options.max_num_trials;
  ppsfm::FeatureLines lines(lines2D.size());
  for (size_t i = 0; i < lines2D.size(); ++i)
    lines[i] = ppsfm::FeatureLine(lines2D[i].Line(), lines2D[i].IsAligned());
  return ppsfm::EstimateAbsolutePoseFromLines(o, lines, points3D, qvec, tvec, num_inliers,
                                              inlier_mask);
}

bool RefineAbsolutePoseFromLines(const AbsolutePoseRefinementOptions& options,
                                 const std::vector<char>& inlier_mask,
                                 const std::vector<Eigen::Vector3d>& lines2D,
                                 const std::vector<Eigen::Vector3d>& points3D,
                                 Eigen::Vector4d* qvec, Eigen::Vector3d* tvec, Camera* camera) {
  ppsfm::AbsolutePoseRefinementOptions o;  // src/estimators/pose.h:84-108
  o.gradient_tolerance = options.gradient_tolerance;
  o.max_num_iterations = options.max_num_iterations;
  o.loss_function_scale = options.loss_function_scale;
  o.refine_focal_length = options.refine_focal_length;
  o.refine_extra_params = options.refine_extra_params;
  o.print_summary = options.print_summary;
  return ppsfm::RefineAbsolutePoseFromLines(o, inlier_mask, lines2D, points3D, qvec, tvec, camera);
}

}  // namespace colmap
