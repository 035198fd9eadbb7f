// ref_cost.cc — TEST INFRASTRUCTURE.  C entry points around the REFERENCE's own line cost
// functors and camera models, compiled from where they lie under /root/reference (never copied):
//   src/base/cost_functions.h    BundleAdjustmentLineCostFunction<CameraModel>::operator()
//                                BundleAdjustmentConstantPoseLineCostFunction<CameraModel>::operator()
//                                                                          
//   src/base/camera_models.{h,cc}  WorldToImage of the 11 models, ImageToWorldThreshold,
//                                parameter index groups                    (A13)
// Ceres, Eigen, glog and Boost are absent in this image: the sources compile against the
// stand-ins of oracle/ref/shim/ (ceres/ceres.h: Jet, UnitQuaternionRotatePoint and the scalar
// functions restated from Ceres' published sources).  The functors are evaluated the way
// ceres::AutoDiffCostFunction does: once with every parameter block seeded as a dual number.
// tests/test_ref_cost.py compares residuals and Jacobians with the oracle's restatement
// (oracle/ba_oracle.cc) and, on the GPU, with the analytic kernel.  Built by oracle/build_ref.sh
// into oracle/_ref/libref_cost.so.
#include <cstring>
#include <cstdint>
#include <vector>

#include "base/camera_models.h"
#include "base/cost_functions.h"

namespace {

constexpr int kMaxParams = 12;

// (2; 4, 3, 3, k): Jacobians row-major 2x4, 2x3, 2x3, 2x12 (columns >= k zero)
template <typename CameraModel>
bool LineCost(const double* cam, const double* line, const double* q, const double* t,
              const double* X, double* r, double* Jq, double* Jt, double* JX, double* Jcam) {
  constexpr int k = static_cast<int>(CameraModel::kNumParams);
  const colmap::BundleAdjustmentLineCostFunction<CameraModel> f(
      Eigen::Vector3d(line[0], line[1], line[2]));
  // residual only: plain doubles (what AutoDiffCostFunction::Evaluate does without Jacobians)
  if (Jq == nullptr) return f(q, t, X, cam, r);
  typedef ceres::Jet<double, 10 + kMaxParams> J;
  J jq[4], jt[3], jX[3], jc[kMaxParams], res[2];
  for (int i = 0; i < 4; ++i) jq[i] = J(q[i], i);
  for (int i = 0; i < 3; ++i) jt[i] = J(t[i], 4 + i);
  for (int i = 0; i < 3; ++i) jX[i] = J(X[i], 7 + i);
  for (int i = 0; i < k; ++i) jc[i] = J(cam[i], 10 + i);
  if (!f(jq, jt, jX, jc, res)) return false;
  for (int row = 0; row < 2; ++row) {
    r[row] = res[row].a;  // with Jacobians, the residual is the value part of the dual numbers
    for (int i = 0; i < 4; ++i) Jq[4 * row + i] = res[row].v[i];
    for (int i = 0; i < 3; ++i) Jt[3 * row + i] = res[row].v[4 + i];
    for (int i = 0; i < 3; ++i) JX[3 * row + i] = res[row].v[7 + i];
    if (Jcam != nullptr)
      for (int i = 0; i < kMaxParams; ++i) Jcam[kMaxParams * row + i] = res[row].v[10 + i];
  }
  return true;
}

// (2; 3, k) with the pose baked into the functor
template <typename CameraModel>
bool ConstantPoseLineCost(const double* cam, const double* line, const double* q, const double* t,
                          const double* X, double* r, double* JX, double* Jcam) {
  constexpr int k = static_cast<int>(CameraModel::kNumParams);
  const colmap::BundleAdjustmentConstantPoseLineCostFunction<CameraModel> f(
      Eigen::Vector4d(q[0], q[1], q[2], q[3]), Eigen::Vector3d(t[0], t[1], t[2]),
      Eigen::Vector3d(line[0], line[1], line[2]));
  if (JX == nullptr) return f(X, cam, r);
  typedef ceres::Jet<double, 3 + kMaxParams> J;
  J jX[3], jc[kMaxParams], res[2];
  for (int i = 0; i < 3; ++i) jX[i] = J(X[i], i);
  for (int i = 0; i < k; ++i) jc[i] = J(cam[i], 3 + i);
  if (!f(jX, jc, res)) return false;
  for (int row = 0; row < 2; ++row) {
    r[row] = res[row].a;
    for (int i = 0; i < 3; ++i) JX[3 * row + i] = res[row].v[i];
    if (Jcam != nullptr)
      for (int i = 0; i < kMaxParams; ++i) Jcam[kMaxParams * row + i] = res[row].v[3 + i];
  }
  return true;
}

}  // namespace

extern "C" {

int ref_camera_num_params(int model) {
  switch (model) {
#define CAMERA_MODEL_CASE(CameraModel) \
  case colmap::CameraModel::kModelId:  \
    return static_cast<int>(colmap::CameraModel::kNumParams);
    CAMERA_MODEL_CASES
#undef CAMERA_MODEL_CASE
  }
  return -1;
}

// groups: 0 focal length, 1 principal point, 2 extra parameters; returns the count
int ref_camera_param_idxs(int model, int group, int* idxs_out) {
  const std::vector<size_t>* v = nullptr;
  switch (model) {
#define CAMERA_MODEL_CASE(CameraModel)                                            \
  case colmap::CameraModel::kModelId:                                             \
    v = group == 0   ? &colmap::CameraModel::focal_length_idxs                    \
        : group == 1 ? &colmap::CameraModel::principal_point_idxs                 \
                     : &colmap::CameraModel::extra_params_idxs;                   \
    break;
    CAMERA_MODEL_CASES
#undef CAMERA_MODEL_CASE
  }
  if (v == nullptr) return -1;
  for (size_t i = 0; i < v->size(); ++i) idxs_out[i] = static_cast<int>((*v)[i]);
  return static_cast<int>(v->size());
}

void ref_world_to_image(int model, const double* params, double u, double v, double* xy) {
  colmap::CameraModelWorldToImage(model, std::vector<double>(params, params + ref_camera_num_params(model)),
                                  u, v, &xy[0], &xy[1]);
}

// CameraModelImageToWorld (src/base/camera_models.cc) on n pixels: xy [n, 2] -> uv [n, 2]
void ref_image_to_world(int model, const double* params, int n, const double* xy, double* uv) {
  const std::vector<double> p(params, params + ref_camera_num_params(model));
  for (int i = 0; i < n; ++i)
    colmap::CameraModelImageToWorld(model, p, xy[2 * i], xy[2 * i + 1], &uv[2 * i], &uv[2 * i + 1]);
}

// CameraModelHasBogusParams (src/base/camera_models.cc:233-253)
int ref_has_bogus_params(int model, const double* params, uint64_t width, uint64_t height,
                         double min_focal_length_ratio, double max_focal_length_ratio,
                         double max_extra_param) {
  return colmap::CameraModelHasBogusParams(
             model, std::vector<double>(params, params + ref_camera_num_params(model)), width, height,
             min_focal_length_ratio, max_focal_length_ratio, max_extra_param)
             ? 1 : 0;
}

double ref_image_to_world_threshold(int model, const double* params, double threshold) {
  return colmap::CameraModelImageToWorldThreshold(
      model, std::vector<double>(params, params + ref_camera_num_params(model)), threshold);
}

int ref_line_cost(int model, const double* cam, const double* line, const double* q,
                  const double* t, const double* X, double* r, double* Jq, double* Jt,
                  double* JX, double* Jcam) {
  switch (model) {
#define CAMERA_MODEL_CASE(CameraModel) \
  case colmap::CameraModel::kModelId:  \
    return LineCost<colmap::CameraModel>(cam, line, q, t, X, r, Jq, Jt, JX, Jcam) ? 1 : 0;
    CAMERA_MODEL_CASES
#undef CAMERA_MODEL_CASE
  }
  return -1;
}

int ref_constant_pose_line_cost(int model, const double* cam, const double* line, const double* q,
                                const double* t, const double* X, double* r, double* JX,
                                double* Jcam) {
  switch (model) {
#define CAMERA_MODEL_CASE(CameraModel) \
  case colmap::CameraModel::kModelId:  \
    return ConstantPoseLineCost<colmap::CameraModel>(cam, line, q, t, X, r, JX, Jcam) ? 1 : 0;
    CAMERA_MODEL_CASES
#undef CAMERA_MODEL_CASE
  }
  return -1;
}

}  // extern "C"
