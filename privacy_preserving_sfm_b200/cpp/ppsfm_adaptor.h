// ppsfm_adaptor.h — header-only C++ adaptor that re-creates the reference's estimator /
// bundle-adjustment API surface on top of the C-ABI of libppsfm_b200.so (include/ppsfm_b200.h).
//
// It is what a maintainer of colmap/privacy_preserving_sfm drops in so that
// IncrementalMapper (src/sfm/incremental_mapper.cc:719-735, 857-858, 929-930) compiles unchanged:
//
//   reference declaration                                   this header
//   ------------------------------------------------------  -----------------------------------
//   RANSACOptions                src/optim/ransac.h:47-76    ppsfm::RANSACOptions
//   FeatureLine / FeatureLines   src/feature/types.h:98-149  ppsfm::FeatureLine / FeatureLines
//   P6LEstimator                 estimators/absolute_pose.h  ppsfm::P6LEstimator
//   RANSAC<P6LEstimator>::Report src/optim/ransac.h:82-99    ppsfm::RANSAC_P6L::Report
//   EstimateAbsolutePoseFromLines   estimators/pose.h:110    ppsfm::EstimateAbsolutePoseFromLines
//   RefineAbsolutePoseFromLines     estimators/pose.h:117    ppsfm::RefineAbsolutePoseFromLines
//   BundleAdjustmentOptions/Config  optim/bundle_adjustment.h ppsfm::BundleAdjustment{Options,Config}
//   BundleAdjuster::Solve/Summary   optim/bundle_adjustment.h ppsfm::BundleAdjuster<Reconstruction>
//   EstimateTriangulation           estimators/triangulation.h:143  ppsfm::EstimateTriangulation
//                                   (+ EstimateTriangulationBatch: all tracks of an image at once)
//   Reconstruction::FilterPoints3D, FilterObservationsWithNegativeDepth
//                                   base/reconstruction.cc:425-460  ppsfm::FilterPoints3D, ...
//   Camera::ImageToWorldThreshold   base/camera_models.h:533-543    ppsfm::ImageToWorldThreshold
//
// Vector types: with -DPPSFM_WITH_EIGEN the Eigen types of the reference are used
// (Eigen::Vector3d, Eigen::Vector4d, Eigen::Matrix3x4d); otherwise std::array stand-ins with the
// same memory layout (.data() -> contiguous doubles), so the header builds without Eigen.
//
// Error convention: bool for expected failures; contract violations abort like
// glog CHECK (PPSFM_CHECK); there are no exceptions and no CPU fallback.
#ifndef PPSFM_ADAPTOR_H_
#define PPSFM_ADAPTOR_H_

#include <array>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "ppsfm_b200.h"

#ifdef PPSFM_WITH_EIGEN
#include <Eigen/Core>
#endif

#define PPSFM_CHECK(cond)                                                              \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      std::fprintf(stderr, "Check failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__);   \
      std::abort();                                                                    \
    }                                                                                  \
  } while (0)

namespace ppsfm {

#ifdef PPSFM_WITH_EIGEN
typedef Eigen::Vector3d Vector3d;
typedef Eigen::Vector4d Vector4d;
typedef Eigen::Matrix<double, 3, 4> Matrix3x4d;
#else
typedef std::array<double, 3> Vector3d;
typedef std::array<double, 4> Vector4d;
typedef std::array<double, 12> Matrix3x4d;  // column-major 3x4, as Eigen::Matrix3x4d in memory
#endif

typedef uint32_t image_t;
typedef uint32_t camera_t;
typedef uint64_t point3D_t;
const point3D_t kInvalidPoint3DId = std::numeric_limits<point3D_t>::max();

// One context per host thread (the reference's PRNG is thread_local too, util/random.cc:36).
inline ppsfm_ctx* ThreadContext(int device = 0) {
  thread_local ppsfm_ctx* ctx = nullptr;
  if (ctx == nullptr) {
    const int rc = ppsfm_ctx_create(device, &ctx);
    if (rc != PPSFM_OK) {
      std::fprintf(stderr, "ppsfm_ctx_create failed (rc=%d): no usable CUDA device\n", rc);
      std::abort();
    }
  }
  return ctx;
}

// SetPRNGSeed (src/util/random.h:52)
inline void SetPRNGSeed(unsigned seed = 0) { ppsfm_set_prng_seed(ThreadContext(), seed); }

// ---- src/feature/types.h:98-138 ----------------------------------------------------------------
struct FeatureLine {
  FeatureLine() = default;
  explicit FeatureLine(const Vector3d& line) : line_(line) {}
  FeatureLine(const Vector3d& line, bool is_aligned) : line_(line), is_aligned_(is_aligned) {}
  FeatureLine(const Vector3d& line, bool is_aligned, point3D_t id)
      : line_(line), is_aligned_(is_aligned), point3D_id_(id) {}
  bool IsAligned() const { return is_aligned_; }
  void SetAligned(bool a) { is_aligned_ = a; }
  const Vector3d& Line() const { return line_; }
  void SetLine(const Vector3d& l) { line_ = l; }
  point3D_t Point3DId() const { return point3D_id_; }
  bool HasPoint3D() const { return point3D_id_ != kInvalidPoint3DId; }
  void SetPoint3DId(point3D_t id) { point3D_id_ = id; }

 private:
  Vector3d line_{};
  bool is_aligned_ = false;
  point3D_t point3D_id_ = kInvalidPoint3DId;
};
typedef std::vector<FeatureLine> FeatureLines;

// ---- src/optim/ransac.h:47-76 --------------------------------------------------------------------
struct RANSACOptions {
  double max_error = 0.0;
  double min_inlier_ratio = 0.1;
  double confidence = 0.99;
  double dyn_num_trials_multiplier = 3.0;
  size_t min_num_trials = 0;
  size_t max_num_trials = std::numeric_limits<size_t>::max();
  void Check() const {
    PPSFM_CHECK(max_error > 0);
    PPSFM_CHECK(min_inlier_ratio >= 0 && min_inlier_ratio <= 1);
    PPSFM_CHECK(confidence >= 0 && confidence <= 1);
    PPSFM_CHECK(min_num_trials <= max_num_trials);
  }
  ppsfm_ransac_options ToC() const {
    return ppsfm_ransac_options{max_error, min_inlier_ratio, confidence,
                                dyn_num_trials_multiplier, min_num_trials, max_num_trials};
  }
};

namespace internal {
inline void Flatten(const FeatureLines& lines, std::vector<double>* l, std::vector<uint8_t>* a) {
  l->resize(3 * lines.size());
  a->resize(lines.size());
  for (size_t i = 0; i < lines.size(); ++i) {
    const double* p = lines[i].Line().data();
    (*l)[3 * i] = p[0]; (*l)[3 * i + 1] = p[1]; (*l)[3 * i + 2] = p[2];
    (*a)[i] = lines[i].IsAligned() ? 1 : 0;
  }
}
template <class V3>
inline void Flatten(const std::vector<V3>& v, std::vector<double>* out) {
  out->resize(3 * v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    const double* p = v[i].data();
    (*out)[3 * i] = p[0]; (*out)[3 * i + 1] = p[1]; (*out)[3 * i + 2] = p[2];
  }
}
inline void CheckRc(int rc) {
  if (rc < 0) {
    std::fprintf(stderr, "libppsfm_b200 failed (rc=%d): %s\n", rc,
                 ppsfm_last_error(ThreadContext()));
    std::abort();
  }
}
}  // namespace internal

// ---- src/estimators/absolute_pose.h:48-75 --------------------------------------------------------
class P6LEstimator {
 public:
  typedef FeatureLine X_t;
  typedef Vector3d Y_t;
  typedef Matrix3x4d M_t;
  static const int kMinNumSamples = 6;

  static std::vector<M_t> Estimate(const std::vector<X_t>& lines2D,
                                   const std::vector<Y_t>& points3D) {
    PPSFM_CHECK(lines2D.size() == 6 && points3D.size() == 6);
    std::vector<double> l, p;
    std::vector<uint8_t> a;
    internal::Flatten(lines2D, &l, &a);
    internal::Flatten(points3D, &p);
    const uint32_t idx[6] = {0, 1, 2, 3, 4, 5};
    double models[96];
    int32_t n = 0;
    internal::CheckRc(ppsfm_p6l_solve_batch(ThreadContext(), l.data(), a.data(), p.data(), 6, idx,
                                            1, models, &n));
    std::vector<M_t> out(n);
    for (int m = 0; m < n; ++m)
      for (int j = 0; j < 12; ++j) out[m].data()[j] = models[12 * m + j];
    return out;
  }

  static void Residuals(const std::vector<X_t>& lines2D, const std::vector<Y_t>& points3D,
                        const M_t& proj_matrix, std::vector<double>* residuals) {
    PPSFM_CHECK(lines2D.size() == points3D.size());
    std::vector<double> l, p;
    std::vector<uint8_t> a;
    internal::Flatten(lines2D, &l, &a);
    internal::Flatten(points3D, &p);
    residuals->resize(lines2D.size());
    uint64_t cnt;
    double sum;
    internal::CheckRc(ppsfm_line_residuals(ThreadContext(), l.data(), p.data(), lines2D.size(),
                                           proj_matrix.data(), 1, 1.0, residuals->data(), &cnt,
                                           &sum));
  }
};

// ---- RANSAC<P6LEstimator, InlierSupportMeasurer, RandomSampler> (src/optim/ransac.h) --------------
class RANSAC_P6L {
 public:
  struct Support {
    size_t num_inliers = 0;
    double residual_sum = std::numeric_limits<double>::max();
  };
  struct Report {
    bool success = false;
    size_t num_trials = 0;
    Support support;
    std::vector<char> inlier_mask;
    Matrix3x4d model{};
  };
  explicit RANSAC_P6L(const RANSACOptions& options) : options_(options) { options.Check(); }

  static size_t ComputeNumTrials(size_t num_inliers, size_t num_samples, double confidence,
                                 double num_trials_multiplier) {
    return ppsfm_compute_num_trials(num_inliers, num_samples, confidence, num_trials_multiplier);
  }

  Report Estimate(const FeatureLines& X, const std::vector<Vector3d>& Y) {
    PPSFM_CHECK(X.size() == Y.size());
    std::vector<double> l, p;
    std::vector<uint8_t> a, mask(X.size());
    internal::Flatten(X, &l, &a);
    internal::Flatten(Y, &p);
    const ppsfm_ransac_options o = options_.ToC();
    ppsfm_ransac_report r;
    internal::CheckRc(ppsfm_ransac_p6l(ThreadContext(), l.data(), a.data(), p.data(), X.size(), &o,
                                       &r, mask.data()));
    Report report;
    report.success = r.success != 0;
    report.num_trials = r.num_trials;
    report.support.num_inliers = r.num_inliers;
    report.support.residual_sum = r.residual_sum;
    for (int j = 0; j < 12; ++j) report.model.data()[j] = r.model[j];
    if (report.success) report.inlier_mask.assign(mask.begin(), mask.end());
    return report;
  }

  P6LEstimator estimator;

 private:
  RANSACOptions options_;
};

// ---- src/estimators/pose.h:110-115 -----------------------------------------------------------------
inline bool EstimateAbsolutePoseFromLines(const RANSACOptions& options, const FeatureLines& lines2D,
                                          const std::vector<Vector3d>& points3D, Vector4d* qvec,
                                          Vector3d* tvec, size_t* num_inliers,
                                          std::vector<char>* inlier_mask) {
  options.Check();
  PPSFM_CHECK(lines2D.size() == points3D.size());
  std::vector<double> l, p;
  std::vector<uint8_t> a, mask(lines2D.size());
  internal::Flatten(lines2D, &l, &a);
  internal::Flatten(points3D, &p);
  const ppsfm_ransac_options o = options.ToC();
  uint64_t ninl = 0;
  ppsfm_ransac_report report;
  const int rc = ppsfm_estimate_absolute_pose_from_lines(
      ThreadContext(), l.data(), a.data(), p.data(), lines2D.size(), &o, qvec->data(),
      tvec->data(), &ninl, mask.data(), &report);
  internal::CheckRc(rc);
  *num_inliers = ninl;
  if (report.success) inlier_mask->assign(mask.begin(), mask.end()); else inlier_mask->clear();
  return rc == PPSFM_OK;
}

// ---- src/estimators/pose.h:84-108, 117-122 -----------------------------------------------------------
struct AbsolutePoseRefinementOptions {
  double gradient_tolerance = 1.0;
  int max_num_iterations = 100;
  double loss_function_scale = 1.0;
  bool refine_focal_length = false;
  bool refine_extra_params = false;
  bool print_summary = true;
  void Check() const {
    PPSFM_CHECK(gradient_tolerance >= 0.0);
    PPSFM_CHECK(max_num_iterations >= 0);
    PPSFM_CHECK(loss_function_scale >= 0.0);
  }
};

// CameraT must provide ModelId() and ParamsData() like colmap::Camera (src/base/camera.h).
template <class CameraT>
inline bool RefineAbsolutePoseFromLines(const AbsolutePoseRefinementOptions& options,
                                        const std::vector<char>& inlier_mask,
                                        const std::vector<Vector3d>& lines2D,
                                        const std::vector<Vector3d>& points3D, Vector4d* qvec,
                                        Vector3d* tvec, CameraT* camera) {
  PPSFM_CHECK(inlier_mask.size() == lines2D.size());
  PPSFM_CHECK(lines2D.size() == points3D.size());
  options.Check();
  std::vector<double> l, p;
  internal::Flatten(lines2D, &l);
  internal::Flatten(points3D, &p);
  std::vector<uint8_t> mask(inlier_mask.begin(), inlier_mask.end());
  ppsfm_ba_summary summary;
  // (camera->ParamsData() is written only when a refine_* flag is set, pose.cc:149-183)
  const int rc = ppsfm_refine_absolute_pose_from_lines_ex(
      ThreadContext(), mask.data(), l.data(), p.data(), lines2D.size(), camera->ModelId(),
      camera->ParamsData(), options.refine_focal_length ? 1 : 0,
      options.refine_extra_params ? 1 : 0, options.gradient_tolerance, options.max_num_iterations,
      options.loss_function_scale, qvec->data(), tvec->data(), &summary);
  internal::CheckRc(rc);
  if (options.print_summary)
    std::printf("Pose refinement report: residuals %lld, iterations %d, cost %g -> %g\n",
                (long long)summary.num_residuals_reduced,
                summary.num_successful_steps + summary.num_unsuccessful_steps,
                summary.initial_cost, summary.final_cost);
  return rc == PPSFM_OK;
}


// ---- src/base/camera_models.h:533-543 (through Camera::ImageToWorldThreshold, camera.cc) -------
// CameraT: .ModelId(), .ParamsData()
template <class CameraT>
inline double ImageToWorldThreshold(const CameraT& camera, double threshold) {
  double out = 0.0;
  PPSFM_CHECK(ppsfm_image_to_world_threshold(camera.ModelId(), camera.ParamsData(), threshold,
                                             &out) == PPSFM_OK);
  return out;
}

// ---- src/estimators/triangulation.h:57-147 ---------------------------------------------------
// EstimateTriangulation: LORANSAC over the 3-combinations of the views of ONE track.  The GPU
// entry point takes many tracks at once (ppsfm_estimate_triangulation_batch); this is the
// reference's per-track signature on top of it, and EstimateTriangulationBatch the form the
// triangulator should call once per image.  CameraT: .ModelId(), .NumParams(), .ParamsData(),
// .Width(), .Height().
struct TriangulationEstimator {
  enum class ResidualType { ANGULAR_ERROR, REPROJECTION_ERROR };
  struct PointData {
    PointData() = default;
    explicit PointData(const Vector3d& l) : line(l) {}
    Vector3d line{};
  };
  template <class CameraT>
  struct PoseData {
    PoseData() = default;
    PoseData(const Matrix3x4d& P, const Vector3d& c, const CameraT* cam)
        : proj_matrix(P), proj_center(c), camera(cam) {}
    Matrix3x4d proj_matrix{};  // [R | t], column-major
    Vector3d proj_center{};
    const CameraT* camera = nullptr;
  };
  static const int kMinNumSamples = 3;
};

struct EstimateTriangulationOptions {
  double min_tri_angle = 0.0;  // radians
  TriangulationEstimator::ResidualType residual_type =
      TriangulationEstimator::ResidualType::ANGULAR_ERROR;
  RANSACOptions ransac_options;
  // tracks up to this length sample all C(n, 3) combinations
  // (src/sfm/incremental_triangulator.cc:527-531 sets min_num_trials that way per track)
  int exhaustive_threshold = 0;
  void Check() const {
    PPSFM_CHECK(min_tri_angle >= 0.0);
    ransac_options.Check();
  }
};

namespace internal {
// a batch of tracks as the flat problem of the C-ABI
template <class CameraT>
struct TrackBatch {
  std::vector<double> qvecs, tvecs, params, lines;
  std::vector<int32_t> image_camera, model, width, height, obs_image;
  std::vector<int64_t> track_start{0};
  std::vector<uint8_t> aligned;
  std::unordered_map<const CameraT*, int> camera_index;
  void AddView(const Vector3d& line, const Matrix3x4d& P, const CameraT* camera) {
    PPSFM_CHECK(camera != nullptr);
    auto it = camera_index.find(camera);
    if (it == camera_index.end()) {
      it = camera_index.emplace(camera, (int)model.size()).first;
      model.push_back(camera->ModelId());
      width.push_back((int32_t)camera->Width());
      height.push_back((int32_t)camera->Height());
      const size_t base = params.size();
      params.resize(base + 12, 0.0);
      for (size_t k = 0; k < camera->NumParams() && k < 12; ++k)
        params[base + k] = camera->ParamsData()[k];
    }
    double q[4];
    ppsfm_rotation_matrix_to_quaternion(P.data(), q);  // first 9 doubles = R, column-major
    obs_image.push_back((int32_t)image_camera.size());
    image_camera.push_back(it->second);
    for (int k = 0; k < 4; ++k) qvecs.push_back(q[k]);
    for (int k = 0; k < 3; ++k) tvecs.push_back(P.data()[9 + k]);
    for (int k = 0; k < 3; ++k) lines.push_back(line.data()[k]);
    aligned.push_back(0);
  }
  void EndTrack() { track_start.push_back((int64_t)obs_image.size()); }
  ppsfm_filter_problem Problem() const {
    ppsfm_filter_problem pb{};
    pb.num_images = (int32_t)image_camera.size();
    pb.qvecs = qvecs.data();
    pb.tvecs = tvecs.data();
    pb.image_camera = image_camera.data();
    pb.num_cameras = (int32_t)model.size();
    pb.camera_model = model.data();
    pb.camera_params = params.data();
    pb.camera_width = width.data();
    pb.camera_height = height.data();
    pb.num_points = (int32_t)track_start.size() - 1;
    pb.points = nullptr;
    pb.track_start = track_start.data();
    pb.num_obs = (int64_t)obs_image.size();
    pb.obs_image = obs_image.data();
    pb.obs_line = lines.data();
    pb.obs_aligned = aligned.data();
    return pb;
  }
};
inline ppsfm_triangulation_options ToC(const EstimateTriangulationOptions& o) {
  ppsfm_triangulation_options c;
  ppsfm_triangulation_options_default(&c);
  c.min_tri_angle = o.min_tri_angle;
  c.residual_type = o.residual_type == TriangulationEstimator::ResidualType::ANGULAR_ERROR ? 0 : 1;
  c.max_error = o.ransac_options.max_error;
  c.min_inlier_ratio = o.ransac_options.min_inlier_ratio;
  c.confidence = o.ransac_options.confidence;
  c.dyn_num_trials_multiplier = o.ransac_options.dyn_num_trials_multiplier;
  c.min_num_trials = o.ransac_options.min_num_trials;
  c.max_num_trials = o.ransac_options.max_num_trials;
  c.exhaustive_threshold = o.exhaustive_threshold;
  return c;
}
}  // namespace internal

// All tracks of a batch in ONE GPU call: point_data[t] / pose_data[t] are the views of track t.
template <class CameraT>
inline void EstimateTriangulationBatch(
    const EstimateTriangulationOptions& options,
    const std::vector<std::vector<TriangulationEstimator::PointData>>& point_data,
    const std::vector<std::vector<TriangulationEstimator::PoseData<CameraT>>>& pose_data,
    std::vector<char>* success, std::vector<std::vector<char>>* inlier_masks,
    std::vector<Vector3d>* xyz) {
  options.Check();
  PPSFM_CHECK(point_data.size() == pose_data.size());
  internal::TrackBatch<CameraT> batch;
  for (size_t t = 0; t < point_data.size(); ++t) {
    PPSFM_CHECK(point_data[t].size() == pose_data[t].size());
    for (size_t i = 0; i < point_data[t].size(); ++i)
      batch.AddView(point_data[t][i].line, pose_data[t][i].proj_matrix, pose_data[t][i].camera);
    batch.EndTrack();
  }
  const size_t T = point_data.size(), O = batch.obs_image.size();
  std::vector<double> x(3 * (T ? T : 1));
  std::vector<uint8_t> ok(T ? T : 1), mask(O ? O : 1);
  const ppsfm_filter_problem pb = batch.Problem();
  const ppsfm_triangulation_options o = internal::ToC(options);
  internal::CheckRc(ppsfm_estimate_triangulation_batch(ThreadContext(), &pb, &o, x.data(),
                                                       ok.data(), mask.data(), nullptr));
  success->assign(T, 0);
  inlier_masks->assign(T, {});
  xyz->assign(T, Vector3d{});
  for (size_t t = 0; t < T; ++t) {
    (*success)[t] = (char)ok[t];
    for (int k = 0; k < 3; ++k) (*xyz)[t].data()[k] = x[3 * t + k];
    for (int64_t k = batch.track_start[t]; k < batch.track_start[t + 1]; ++k)
      (*inlier_masks)[t].push_back((char)mask[k]);
  }
}

// bool EstimateTriangulation(options, point_data, pose_data, &inlier_mask, &xyz)
// (src/estimators/triangulation.h:143-147)
template <class CameraT>
inline bool EstimateTriangulation(
    const EstimateTriangulationOptions& options,
    const std::vector<TriangulationEstimator::PointData>& point_data,
    const std::vector<TriangulationEstimator::PoseData<CameraT>>& pose_data,
    std::vector<char>* inlier_mask, Vector3d* xyz) {
  // src/estimators/triangulation.cc:122-130: two views are a contract violation only below 2;
  // exactly two return false
  PPSFM_CHECK(inlier_mask != nullptr && xyz != nullptr);
  PPSFM_CHECK(point_data.size() >= 2);
  PPSFM_CHECK(point_data.size() == pose_data.size());
  options.Check();
  if (point_data.size() < 3) return false;
  std::vector<char> ok;
  std::vector<std::vector<char>> masks;
  std::vector<Vector3d> pts;
  EstimateTriangulationBatch<CameraT>(options, {point_data}, {pose_data}, &ok, &masks, &pts);
  if (!ok[0]) return false;
  *inlier_mask = masks[0];
  *xyz = pts[0];
  return true;
}

// ---- src/base/reconstruction.cc:425-460, 594-719 ---------------------------------------------
// Reconstruction::FilterPoints3D / FilterObservationsWithNegativeDepth as free functions over a
// Reconstruction-like object: the tracks of `point3D_ids` go to the GPU as one problem, the
// delete masks come back and are applied through the reconstruction's own DeleteObservation /
// DeletePoint3D / Point3D::SetError.  ReconstructionT needs, beyond what BundleAdjuster uses:
//   ExistsPoint3D(id), DeleteObservation(image_id, line_idx), DeletePoint3D(id),
//   Point3D(id).SetError(e), Image(id).Lines()[i].IsAligned(), Camera(id).Width() / .Height().
namespace internal {
template <class ReconstructionT>
struct TrackProblem {
  std::vector<double> qvecs, tvecs, params, points, lines;
  std::vector<int32_t> image_camera, model, width, height, obs_image;
  std::vector<int64_t> track_start{0};
  std::vector<uint8_t> aligned;
  std::vector<point3D_t> point_ids;
  std::vector<image_t> image_ids;
  std::vector<std::pair<image_t, uint32_t>> obs_ref;
  std::unordered_map<image_t, int> image_index;
  std::unordered_map<camera_t, int> camera_index;
  TrackProblem(ReconstructionT* rec, const std::vector<point3D_t>& ids) {
    for (const point3D_t pid : ids) {
      if (!rec->ExistsPoint3D(pid)) continue;
      auto& point3D = rec->Point3D(pid);
      point_ids.push_back(pid);
      for (int k = 0; k < 3; ++k) points.push_back(point3D.XYZ().data()[k]);
      for (const auto& el : point3D.Track().Elements()) {
        auto it = image_index.find(el.image_id);
        if (it == image_index.end()) {
          auto& image = rec->Image(el.image_id);
          it = image_index.emplace(el.image_id, (int)image_ids.size()).first;
          image_ids.push_back(el.image_id);
          for (int k = 0; k < 4; ++k) qvecs.push_back(image.Qvec().data()[k]);
          for (int k = 0; k < 3; ++k) tvecs.push_back(image.Tvec().data()[k]);
          const camera_t cid = image.CameraId();
          auto ct = camera_index.find(cid);
          if (ct == camera_index.end()) {
            auto& camera = rec->Camera(cid);
            ct = camera_index.emplace(cid, (int)model.size()).first;
            model.push_back(camera.ModelId());
            width.push_back((int32_t)camera.Width());
            height.push_back((int32_t)camera.Height());
            const size_t base = params.size();
            params.resize(base + 12, 0.0);
            for (size_t k = 0; k < camera.NumParams() && k < 12; ++k)
              params[base + k] = camera.ParamsData()[k];
          }
          image_camera.push_back(ct->second);
        }
        const auto& line = rec->Image(el.image_id).Lines()[el.line_idx];
        obs_image.push_back(it->second);
        for (int k = 0; k < 3; ++k) lines.push_back(line.Line().data()[k]);
        aligned.push_back(line.IsAligned() ? 1 : 0);
        obs_ref.emplace_back(el.image_id, el.line_idx);
      }
      track_start.push_back((int64_t)obs_image.size());
    }
  }
  ppsfm_filter_problem Problem() const {
    ppsfm_filter_problem pb{};
    pb.num_images = (int32_t)image_ids.size();
    pb.qvecs = qvecs.data();
    pb.tvecs = tvecs.data();
    pb.image_camera = image_camera.data();
    pb.num_cameras = (int32_t)model.size();
    pb.camera_model = model.data();
    pb.camera_params = params.data();
    pb.camera_width = width.data();
    pb.camera_height = height.data();
    pb.num_points = (int32_t)point_ids.size();
    pb.points = points.data();
    pb.track_start = track_start.data();
    pb.num_obs = (int64_t)obs_image.size();
    pb.obs_image = obs_image.data();
    pb.obs_line = lines.data();
    pb.obs_aligned = aligned.data();
    return pb;
  }
  // applies the masks: whole points first, then single observations of the survivors
  void Apply(ReconstructionT* rec, const std::vector<uint8_t>& obs_deleted,
             const std::vector<uint8_t>& point_deleted, const double* point_error) const {
    for (size_t p = 0; p < point_ids.size(); ++p) {
      if (point_deleted[p]) {
        rec->DeletePoint3D(point_ids[p]);
        continue;
      }
      for (int64_t k = track_start[p]; k < track_start[p + 1]; ++k)
        if (obs_deleted[k]) rec->DeleteObservation(obs_ref[k].first, obs_ref[k].second);
      if (point_error) rec->Point3D(point_ids[p]).SetError(point_error[p]);
    }
  }
};
}  // namespace internal

template <class ReconstructionT, class IdContainer>
inline size_t FilterPoints3D(ReconstructionT* rec, double max_reproj_error, double min_tri_angle,
                             const IdContainer& point3D_ids) {
  const std::vector<point3D_t> ids(point3D_ids.begin(), point3D_ids.end());
  internal::TrackProblem<ReconstructionT> tp(rec, ids);
  const size_t P = tp.point_ids.size(), O = tp.obs_image.size();
  if (P == 0) return 0;
  std::vector<uint8_t> od(O ? O : 1), pd(P);
  std::vector<double> err(P, -1.0);
  size_t num_filtered = 0;
  const ppsfm_filter_problem pb = tp.Problem();
  internal::CheckRc(ppsfm_filter_points3d(ThreadContext(), &pb, max_reproj_error, min_tri_angle,
                                          od.data(), pd.data(), err.data(), &num_filtered));
  tp.Apply(rec, od, pd, err.data());
  return num_filtered;
}

// `point3D_ids`: every point of the reconstruction (the reference walks the registered images;
// the set of observations visited is the same)
template <class ReconstructionT, class IdContainer>
inline size_t FilterObservationsWithNegativeDepth(ReconstructionT* rec,
                                                  const IdContainer& point3D_ids) {
  const std::vector<point3D_t> ids(point3D_ids.begin(), point3D_ids.end());
  internal::TrackProblem<ReconstructionT> tp(rec, ids);
  const size_t P = tp.point_ids.size(), O = tp.obs_image.size();
  if (P == 0) return 0;
  std::vector<uint8_t> od(O ? O : 1), pd(P);
  size_t num_filtered = 0;
  const ppsfm_filter_problem pb = tp.Problem();
  internal::CheckRc(ppsfm_filter_observations_with_negative_depth(
      ThreadContext(), &pb, od.data(), pd.data(), &num_filtered));
  tp.Apply(rec, od, pd, nullptr);
  return num_filtered;
}

// ---- src/optim/bundle_adjustment.h:49-100 --------------------------------------------------------------
struct BundleAdjustmentOptions {
  enum class LossFunctionType { TRIVIAL, SOFT_L1, CAUCHY };
  LossFunctionType loss_function_type = LossFunctionType::TRIVIAL;
  double loss_function_scale = 1.0;
  bool refine_focal_length = false;
  bool refine_principal_point = false;
  bool refine_extra_params = false;
  bool refine_extrinsics = true;
  bool print_summary = true;
  int min_num_residuals_for_multi_threading = 50000;
  // stands in for ceres::Solver::Options (same field names for the members the reference sets,
  // src/controllers/incremental_mapper.cc:196-243)
  ppsfm_ba_options solver_options;
  BundleAdjustmentOptions() { ppsfm_ba_options_default(&solver_options); }
  bool Check() const {
    if (!(loss_function_scale >= 0)) {
      std::fprintf(stderr, "CHECK_OPTION_GE(loss_function_scale, 0) failed\n");
      return false;
    }
    return true;
  }
};

// ---- src/optim/bundle_adjustment.h:103-167 -------------------------------------------------------------
class BundleAdjustmentConfig {
 public:
  size_t NumImages() const { return image_ids_.size(); }
  size_t NumPoints() const { return variable_point3D_ids_.size() + constant_point3D_ids_.size(); }
  size_t NumConstantCameras() const { return constant_camera_ids_.size(); }
  size_t NumConstantPoses() const { return constant_poses_.size(); }
  size_t NumConstantTvecs() const { return constant_tvecs_.size(); }
  size_t NumVariablePoints() const { return variable_point3D_ids_.size(); }
  size_t NumConstantPoints() const { return constant_point3D_ids_.size(); }
  void AddImage(image_t id) { image_ids_.insert(id); }
  bool HasImage(image_t id) const { return image_ids_.count(id) > 0; }
  void RemoveImage(image_t id) { image_ids_.erase(id); }
  void SetConstantCamera(camera_t id) { constant_camera_ids_.insert(id); }
  void SetVariableCamera(camera_t id) { constant_camera_ids_.erase(id); }
  bool IsConstantCamera(camera_t id) const { return constant_camera_ids_.count(id) > 0; }
  void SetConstantPose(image_t id) {
    PPSFM_CHECK(HasImage(id));
    PPSFM_CHECK(!HasConstantTvec(id));
    constant_poses_.insert(id);
  }
  void SetVariablePose(image_t id) { constant_poses_.erase(id); }
  bool HasConstantPose(image_t id) const { return constant_poses_.count(id) > 0; }
  void SetConstantTvec(image_t id, const std::vector<int>& idxs) {
    PPSFM_CHECK(idxs.size() > 0 && idxs.size() <= 3);
    PPSFM_CHECK(HasImage(id));
    PPSFM_CHECK(!HasConstantPose(id));
    for (size_t i = 0; i < idxs.size(); ++i)
      for (size_t j = i + 1; j < idxs.size(); ++j) PPSFM_CHECK(idxs[i] != idxs[j]);
    constant_tvecs_[id] = idxs;
  }
  void RemoveConstantTvec(image_t id) { constant_tvecs_.erase(id); }
  bool HasConstantTvec(image_t id) const { return constant_tvecs_.count(id) > 0; }
  void AddVariablePoint(point3D_t id) {
    PPSFM_CHECK(!HasConstantPoint(id));
    variable_point3D_ids_.insert(id);
  }
  void AddConstantPoint(point3D_t id) {
    PPSFM_CHECK(!HasVariablePoint(id));
    constant_point3D_ids_.insert(id);
  }
  bool HasPoint(point3D_t id) const { return HasVariablePoint(id) || HasConstantPoint(id); }
  bool HasVariablePoint(point3D_t id) const { return variable_point3D_ids_.count(id) > 0; }
  bool HasConstantPoint(point3D_t id) const { return constant_point3D_ids_.count(id) > 0; }
  void RemoveVariablePoint(point3D_t id) { variable_point3D_ids_.erase(id); }
  void RemoveConstantPoint(point3D_t id) { constant_point3D_ids_.erase(id); }
  const std::unordered_set<image_t>& Images() const { return image_ids_; }
  const std::unordered_set<point3D_t>& VariablePoints() const { return variable_point3D_ids_; }
  const std::unordered_set<point3D_t>& ConstantPoints() const { return constant_point3D_ids_; }
  const std::vector<int>& ConstantTvec(image_t id) const { return constant_tvecs_.at(id); }

 private:
  std::unordered_set<camera_t> constant_camera_ids_;
  std::unordered_set<image_t> image_ids_;
  std::unordered_set<point3D_t> variable_point3D_ids_;
  std::unordered_set<point3D_t> constant_point3D_ids_;
  std::unordered_set<image_t> constant_poses_;
  std::unordered_map<image_t, std::vector<int>> constant_tvecs_;
};

// ---- src/optim/bundle_adjustment.h:171-215 -------------------------------------------------------------
// ReconstructionT must offer the accessors BundleAdjuster uses on colmap::Reconstruction:
//   Image(image_t)   -> .CameraId(), .NormalizeQvec(), .Qvec().data(), .Tvec().data(),
//                       .Lines() (vector of FeatureLine-like: HasPoint3D(), Point3DId(), Line())
//   Camera(camera_t) -> .ModelId(), .NumParams(), .ParamsData()
//   Point3D(id)      -> .XYZ().data(), .Track().Length(), .Track().Elements() ({image_id, line_idx})
template <class ReconstructionT>
class BundleAdjuster {
 public:
  BundleAdjuster(const BundleAdjustmentOptions& options, const BundleAdjustmentConfig& config)
      : options_(options), config_(config) {
    PPSFM_CHECK(options_.Check());
  }

  const ppsfm_ba_summary& Summary() const { return summary_; }

  // The assembly alone (what SetUp, bundle_adjustment.cc:326-542, decides), without a solve: the
  // flat problem Solve() would hand to ppsfm_ba_solve.  For tests of the assembly rules against
  // the reference's own SetUp (tests/test_ref_ba_setup.py); uses up the adjuster like Solve().
  struct Assembly {
    std::vector<image_t> image_ids;      // problem order
    std::vector<camera_t> camera_ids;
    std::vector<point3D_t> point_ids;
    std::vector<uint8_t> pose_flags;     // per image: 1 constant pose, 2 << k: tvec[k] constant
    std::vector<uint8_t> point_const;    // per point
    std::vector<uint8_t> camera_const;   // per camera
    std::vector<int32_t> image_camera, obs_image, obs_point;
    std::vector<double> obs_line;
  };
  Assembly AssembleOnly(ReconstructionT* reconstruction) {
    PPSFM_CHECK(reconstruction != nullptr);
    PPSFM_CHECK(!used_);
    used_ = true;
    SetUp(reconstruction);
    return Assembly{image_ids_,  camera_ids_,   point_ids_, pose_flags_, point_const_,
                    camera_const_, image_camera_, obs_image_, obs_point_,  obs_line_};
  }

  bool Solve(ReconstructionT* reconstruction) {
    PPSFM_CHECK(reconstruction != nullptr);
    PPSFM_CHECK(!used_);  // "Cannot use the same BundleAdjuster multiple times"
    used_ = true;
    SetUp(reconstruction);
    if (obs_image_.empty()) return false;  // problem_->NumResiduals() == 0
    ppsfm_ba_problem pb;
    pb.num_images = (int32_t)image_ids_.size();
    pb.qvecs = qvecs_.data();
    pb.tvecs = tvecs_.data();
    pb.pose_flags = pose_flags_.data();
    pb.image_camera = image_camera_.data();
    pb.num_cameras = (int32_t)camera_model_.size();
    pb.camera_model = camera_model_.data();
    pb.camera_params = camera_params_.data();
    pb.num_points = (int32_t)point_ids_.size();
    pb.points = points_.data();
    pb.point_const = point_const_.data();
    pb.num_obs = (int64_t)obs_image_.size();
    pb.obs_image = obs_image_.data();
    pb.obs_point = obs_point_.data();
    pb.obs_line = obs_line_.data();
    pb.camera_const = camera_const_.data();  // ParameterizeCameras (bundle_adjustment.cc:490-528)
    ppsfm_ba_options o = options_.solver_options;
    o.refine_focal_length = options_.refine_focal_length ? 1 : 0;
    o.refine_principal_point = options_.refine_principal_point ? 1 : 0;
    o.refine_extra_params = options_.refine_extra_params ? 1 : 0;
    o.loss_type = (int32_t)options_.loss_function_type;
    o.loss_scale = options_.loss_function_scale;
    const int rc = ppsfm_ba_solve(ThreadContext(), &pb, &o, &summary_);
    internal::CheckRc(rc);
    // write the result back in place (image.Qvec().data() etc., bundle_adjustment.cc:357-359)
    for (size_t i = 0; i < image_ids_.size(); ++i) {
      auto& image = reconstruction->Image(image_ids_[i]);
      for (int k = 0; k < 4; ++k) image.Qvec().data()[k] = qvecs_[4 * i + k];
      for (int k = 0; k < 3; ++k) image.Tvec().data()[k] = tvecs_[3 * i + k];
    }
    for (size_t i = 0; i < point_ids_.size(); ++i) {
      double* xyz = reconstruction->Point3D(point_ids_[i]).XYZ().data();
      for (int k = 0; k < 3; ++k) xyz[k] = points_[3 * i + k];
    }
    if (o.refine_focal_length || o.refine_principal_point || o.refine_extra_params)
      for (size_t c = 0; c < camera_ids_.size(); ++c) {  // camera.ParamsData() updated in place
        auto& camera = reconstruction->Camera(camera_ids_[c]);
        for (size_t k = 0; k < camera.NumParams() && k < 12; ++k)
          camera.ParamsData()[k] = camera_params_[12 * c + k];
      }
    if (options_.print_summary)
      std::printf("Bundle adjustment report: residuals %lld, iterations %d, cost %g -> %g\n",
                  (long long)summary_.num_residuals_reduced,
                  summary_.num_successful_steps + summary_.num_unsuccessful_steps,
                  summary_.initial_cost, summary_.final_cost);
    return rc == PPSFM_OK;
  }

 private:
  int ImageIndex(ReconstructionT* rec, image_t id, bool constant) {
    auto it = image_index_.find(id);
    if (it != image_index_.end()) return it->second;
    auto& image = rec->Image(id);
    const int idx = (int)image_ids_.size();
    image_index_[id] = idx;
    image_ids_.push_back(id);
    for (int k = 0; k < 4; ++k) qvecs_.push_back(image.Qvec().data()[k]);
    for (int k = 0; k < 3; ++k) tvecs_.push_back(image.Tvec().data()[k]);
    uint8_t f = constant ? 1 : 0;
    if (!constant && config_.HasConstantTvec(id))
      for (int k : config_.ConstantTvec(id)) f |= (uint8_t)(2 << k);
    pose_flags_.push_back(f);
    const camera_t cid = image.CameraId();
    auto ct = camera_index_.find(cid);
    if (ct == camera_index_.end()) {
      auto& camera = rec->Camera(cid);
      ct = camera_index_.emplace(cid, (int)camera_model_.size()).first;
      camera_ids_.push_back(cid);
      camera_model_.push_back(camera.ModelId());
      const size_t base = camera_params_.size();
      camera_params_.resize(base + 12, 0.0);
      for (size_t k = 0; k < camera.NumParams() && k < 12; ++k)
        camera_params_[base + k] = camera.ParamsData()[k];
    }
    image_camera_.push_back(ct->second);
    return idx;
  }
  int PointIndex(ReconstructionT* rec, point3D_t id) {
    auto it = point_index_.find(id);
    if (it != point_index_.end()) return it->second;
    const int idx = (int)point_ids_.size();
    point_index_[id] = idx;
    point_ids_.push_back(id);
    const double* xyz = rec->Point3D(id).XYZ().data();
    for (int k = 0; k < 3; ++k) points_.push_back(xyz[k]);
    point_const_.push_back(0);
    return idx;
  }
  void AddObservation(int img, int pt, const double* line) {
    obs_image_.push_back(img);
    obs_point_.push_back(pt);
    for (int k = 0; k < 3; ++k) obs_line_.push_back(line[k]);
  }
  // bundle_adjustment.cc:326-542
  void SetUp(ReconstructionT* rec) {
    std::unordered_map<point3D_t, size_t> num_obs;
    std::unordered_set<camera_t> config_cameras, outside_cameras;
    for (const image_t image_id : config_.Images()) {  // AddImageToProblem (:348-435)
      auto& image = rec->Image(image_id);
      image.NormalizeQvec();
      const bool constant_pose = !options_.refine_extrinsics || config_.HasConstantPose(image_id);
      for (const auto& line : image.Lines()) {
        if (!line.HasPoint3D()) continue;
        num_obs[line.Point3DId()] += 1;
        const int ii = ImageIndex(rec, image_id, constant_pose);
        AddObservation(ii, PointIndex(rec, line.Point3DId()), line.Line().data());
        config_cameras.insert(image.CameraId());  // camera_ids_.insert (:432-434)
      }
    }
    auto add_point = [&](point3D_t pid) {  // AddPointToProblem (:437-488)
      auto& point3D = rec->Point3D(pid);
      if (num_obs[pid] == point3D.Track().Length()) return;
      for (const auto& el : point3D.Track().Elements()) {
        if (config_.HasImage(el.image_id)) continue;
        num_obs[pid] += 1;
        auto& image = rec->Image(el.image_id);
        // a camera that enters only through images outside the configuration is constant
        // (config_.SetConstantCamera, :476-479)
        if (config_cameras.count(image.CameraId()) == 0) outside_cameras.insert(image.CameraId());
        const int ii = ImageIndex(rec, el.image_id, /*constant=*/true);
        AddObservation(ii, PointIndex(rec, pid), image.Lines()[el.line_idx].Line().data());
      }
    };
    for (const point3D_t pid : config_.VariablePoints()) add_point(pid);
    for (const point3D_t pid : config_.ConstantPoints()) add_point(pid);
    for (const auto& el : num_obs) {  // ParameterizePoints (:530-542)
      auto it = point_index_.find(el.first);
      if (it == point_index_.end()) continue;
      if (rec->Point3D(el.first).Track().Length() > el.second) point_const_[it->second] = 1;
    }
    for (const point3D_t pid : config_.ConstantPoints()) {
      auto it = point_index_.find(pid);
      if (it != point_index_.end()) point_const_[it->second] = 1;
    }
    // ParameterizeCameras (:490-528): constant if the config says so or AddPointToProblem did
    camera_const_.assign(camera_ids_.size(), 0);
    for (size_t c = 0; c < camera_ids_.size(); ++c)
      camera_const_[c] = (config_.IsConstantCamera(camera_ids_[c]) ||
                          outside_cameras.count(camera_ids_[c]) > 0) ? 1 : 0;
  }

  BundleAdjustmentOptions options_;
  BundleAdjustmentConfig config_;
  ppsfm_ba_summary summary_{};
  bool used_ = false;
  std::unordered_map<image_t, int> image_index_;
  std::unordered_map<camera_t, int> camera_index_;
  std::unordered_map<point3D_t, int> point_index_;
  std::vector<image_t> image_ids_;
  std::vector<camera_t> camera_ids_;
  std::vector<point3D_t> point_ids_;
  std::vector<double> qvecs_, tvecs_, points_, camera_params_, obs_line_;
  std::vector<uint8_t> pose_flags_, point_const_, camera_const_;
  std::vector<int32_t> image_camera_, camera_model_, obs_image_, obs_point_;
};

}  // namespace ppsfm

#endif  // PPSFM_ADAPTOR_H_
