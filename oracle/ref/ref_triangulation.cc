// ref_triangulation.cc — TEST INFRASTRUCTURE.  The REFERENCE's own robust-estimation loop of the
// line triangulation, compiled from where it lies under /root/reference:
//   src/optim/loransac.h               LORANSAC<>::Estimate (local optimisation, abort rule)
//   src/optim/ransac.h                 constructor cap, ComputeNumTrials
//   src/optim/combination_sampler.cc   CombinationSampler
//   src/util/math.{h,cc}               NChooseK, NextCombination
//   src/optim/support_measurement.cc   InlierSupportMeasurer
// driving the ORACLE's per-track estimator (oracle/triangulation_oracle.cc: Track::Estimate,
// Track::Residual — included below, not linked) exactly the way EstimateTriangulation
// (src/estimators/triangulation.cc:117-149) and its caller (src/sfm/incremental_triangulator.cc:
// 518-533) drive TriangulationEstimator.  The estimator itself (TriangulateMultiViewPoint's
// JacobiSVD, projection.cc) needs too much of Eigen to compile here; what this pins, bit for bit,
// is the control flow the oracle restates inline: sampling order, the support comparison, when
// the local optimisation runs and wins, the dynamic trial bound, the final mask.  Same approach
// as ref_init.cc (row A17).  Built by oracle/build_ref.sh into oracle/_ref/libref_tri.so.
#include "../triangulation_oracle.cc"  // Problem, Options, Track, MakePose (anonymous namespace)

#include <array>

#include "optim/combination_sampler.h"
#include "optim/loransac.h"
#include "util/math.h"

namespace {

struct TrackEstimator {
  typedef size_t X_t;  // observation index inside the track
  typedef size_t Y_t;
  typedef std::array<double, 3> M_t;
  static constexpr int kMinNumSamples = 3;  // TriangulationEstimator::kMinNumSamples
  const Track* track = nullptr;

  std::vector<M_t> Estimate(const std::vector<X_t>& X, const std::vector<Y_t>&) const {
    double P[3];
    if (!track->Estimate(X, P)) return std::vector<M_t>();
    return std::vector<M_t>{M_t{{P[0], P[1], P[2]}}};
  }
  void Residuals(const std::vector<X_t>& X, const std::vector<Y_t>&, const M_t& xyz,
                 std::vector<double>* residuals) const {
    residuals->resize(X.size());
    for (size_t i = 0; i < X.size(); ++i) (*residuals)[i] = track->Residual(X[i], xyz.data());
  }
};

}  // namespace

extern "C" int ref_estimate_triangulation_batch(const Problem* pbp, const Options* optp,
                                                double* xyz, uint8_t* success,
                                                uint8_t* inlier_mask, uint32_t* num_trials) {
  const Problem& pb = *pbp;
  const Options& opt = *optp;
  for (int t = 0; t < pb.num_points; ++t) {
    const int64_t k0 = pb.track_start[t], k1 = pb.track_start[t + 1];
    const size_t n = static_cast<size_t>(k1 - k0);
    success[t] = 0;
    if (num_trials) num_trials[t] = 0;
    for (int64_t k = k0; k < k1; ++k) inlier_mask[k] = 0;
    if (n < 3) continue;  // triangulation.cc:129-130
    Track tr;
    tr.pb = &pb;
    tr.opt = &opt;
    for (int64_t k = k0; k < k1; ++k) {
      const int img = pb.obs_image[k];
      tr.pose.push_back(MakePose(pb.qvecs + 4 * (size_t)img, pb.tvecs + 3 * (size_t)img));
      tr.line.push_back(pb.obs_line + 3 * (size_t)k);
      tr.cam.push_back(pb.image_camera[img]);
    }
    colmap::RANSACOptions ro;
    ro.max_error = opt.max_error;
    ro.min_inlier_ratio = opt.min_inlier_ratio;
    ro.confidence = opt.confidence;
    ro.dyn_num_trials_multiplier = opt.dyn_num_trials_multiplier;
    ro.min_num_trials = opt.min_num_trials;
    ro.max_num_trials = opt.max_num_trials;
    // incremental_triangulator.cc:527-531: exhaustive sampling for short tracks
    if (static_cast<int>(n) <= opt.exhaustive_threshold) ro.min_num_trials = colmap::NChooseK(n, 3);
    colmap::LORANSAC<TrackEstimator, TrackEstimator, colmap::InlierSupportMeasurer,
                     colmap::CombinationSampler>
        ransac(ro);
    ransac.estimator.track = &tr;
    ransac.local_estimator.track = &tr;
    std::vector<size_t> X(n);
    std::iota(X.begin(), X.end(), 0);
    const auto report = ransac.Estimate(X, X);
    if (num_trials) num_trials[t] = static_cast<uint32_t>(report.num_trials);
    if (!report.success) continue;
    success[t] = 1;
    for (int c = 0; c < 3; ++c) xyz[3 * (size_t)t + c] = report.model[c];
    for (size_t i = 0; i < n; ++i) inlier_mask[k0 + i] = report.inlier_mask[i] ? 1 : 0;
  }
  return 0;
}
