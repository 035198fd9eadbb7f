// ref_ba_setup.cc — TEST INFRASTRUCTURE.  The REFERENCE's own bundle-adjustment ASSEMBLY
// (SURVEY §8 rows A14 / A15), compiled from where it lies under /root/reference:
//   src/optim/bundle_adjustment.cc   BundleAdjustmentOptions / BundleAdjustmentConfig,
//                                    BundleAdjuster::Solve -> SetUp -> AddImageToProblem,
//                                    AddPointToProblem, ParameterizeCameras, ParameterizePoints
//   src/base/{reconstruction,image,point3d,track,camera,camera_models,pose,projection,
//   triangulation}.cc, src/util/{string,misc,threading,timer,logging,math}.cc
//                                    the classes SetUp reads
// against the stand-ins of oracle/ref/shim/.  ceres::Problem is a RECORDER there (it keeps the
// residual blocks, constant blocks and parameterisations SetUp creates; ceres::Solve keeps the
// options and solves nothing), so one call of the reference's BundleAdjuster::Solve yields what
// the reference WOULD hand to Ceres.  The same colmap::Reconstruction then goes through the
// product's adaptor (ppsfm::BundleAdjuster<colmap::Reconstruction>::AssembleOnly, the flat problem
// of the C-ABI) and both are written out in one canonical form for tests/test_ref_ba_setup.py.
//
// The colmap::Reconstruction is built through the reference's own AddCamera / AddImage /
// RegisterImage / AddPoint3D (src/base/reconstruction.cc).  The access hack below opens the
// reference's classes only to READ what SetUp produced (BundleAdjuster::problem_, the functors'
// line / pose members); base/database.cc (SQLite) is not built: its one constant that the compiled
// code names is defined here.
#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <climits>
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <future>
#include <iomanip>
#include <iostream>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

// the stand-ins (include-guarded) before the access hack below, so that only the reference's own
// classes are opened up
#include <Eigen/Dense>
#include <boost/algorithm/string.hpp>
#include <boost/filesystem.hpp>
#include <ceres/ceres.h>
#include <glog/logging.h>

#define private public
#define protected public
#include "base/cost_functions.h"
#include "base/reconstruction.h"
#include "optim/bundle_adjustment.h"
#undef private
#undef protected

#include "base/camera_models.h"
#include "base/database.h"
#include "base/pose.h"

#ifndef PPSFM_WITH_EIGEN
#define PPSFM_WITH_EIGEN
#endif
#include "ppsfm_adaptor.h"

namespace colmap {  // database.cc:229-230
const size_t Database::kMaxNumImages = static_cast<size_t>(std::numeric_limits<int32_t>::max());
}  // namespace colmap

namespace {

// ---- flat scene description (all arrays owned by the caller) --------------------------------
struct Scene {
  int32_t num_cameras;
  const int32_t* camera_model;     // [num_cameras]
  const double* camera_params;     // [num_cameras][12]
  int32_t num_images;
  const int32_t* image_camera;     // [num_images] camera index
  const double* qvecs;             // [num_images][4]
  const double* tvecs;             // [num_images][3]
  int32_t num_points;
  const double* points;            // [num_points][3]
  int64_t num_lines;               // all lines of all images, image-major
  const int64_t* image_line_start; // [num_images + 1]
  const double* lines;             // [num_lines][3]
  const int32_t* line_point;       // [num_lines] point index or -1
};
struct Config {
  int32_t num_config_images; const int32_t* config_images;
  int32_t num_constant_poses; const int32_t* constant_poses;
  int32_t num_constant_tvecs; const int32_t* constant_tvec_image; const int32_t* constant_tvec_mask;
  int32_t num_variable_points; const int32_t* variable_points;
  int32_t num_constant_points; const int32_t* constant_points;
  int32_t num_constant_cameras; const int32_t* constant_cameras;
  int32_t loss_type; double loss_scale;
  int32_t refine_focal_length, refine_principal_point, refine_extra_params, refine_extrinsics;
};

// ids are index + 1 (AddPoint3D hands out 1, 2, ... in the order of the calls)
void BuildReconstruction(const Scene& s, colmap::Reconstruction* rec) {
  for (int c = 0; c < s.num_cameras; ++c) {
    colmap::Camera cam;
    cam.SetCameraId(c + 1);
    cam.SetModelId(s.camera_model[c]);
    cam.SetWidth(1000);
    cam.SetHeight(1000);
    cam.SetParams(std::vector<double>(s.camera_params + 12 * c,
                                      s.camera_params + 12 * c + cam.NumParams()));
    rec->AddCamera(cam);
  }
  std::vector<colmap::Track> tracks(s.num_points);
  for (int i = 0; i < s.num_images; ++i) {
    colmap::Image img;
    img.SetImageId(i + 1);
    img.SetCameraId(s.image_camera[i] + 1);
    img.Qvec() = Eigen::Vector4d(s.qvecs[4 * i], s.qvecs[4 * i + 1], s.qvecs[4 * i + 2], s.qvecs[4 * i + 3]);
    img.Tvec() = Eigen::Vector3d(s.tvecs[3 * i], s.tvecs[3 * i + 1], s.tvecs[3 * i + 2]);
    colmap::FeatureLines lines;
    for (int64_t k = s.image_line_start[i]; k < s.image_line_start[i + 1]; ++k) {
      lines.emplace_back(Eigen::Vector3d(s.lines[3 * k], s.lines[3 * k + 1], s.lines[3 * k + 2]), false);
      if (s.line_point[k] >= 0)
        tracks[s.line_point[k]].AddElement(i + 1, static_cast<colmap::point2D_t>(k - s.image_line_start[i]));
    }
    img.SetLines(lines);
    rec->AddImage(img);
    rec->RegisterImage(i + 1);
  }
  for (int p = 0; p < s.num_points; ++p) {
    const colmap::point3D_t id = rec->AddPoint3D(
        Eigen::Vector3d(s.points[3 * p], s.points[3 * p + 1], s.points[3 * p + 2]), tracks[p]);
    if (id != static_cast<colmap::point3D_t>(p + 1)) std::abort();
  }
}

template <class ConfigT>
void FillConfig(const Config& c, ConfigT* cfg) {
  for (int i = 0; i < c.num_config_images; ++i) cfg->AddImage(c.config_images[i] + 1);
  for (int i = 0; i < c.num_constant_poses; ++i) cfg->SetConstantPose(c.constant_poses[i] + 1);
  for (int i = 0; i < c.num_constant_tvecs; ++i) {
    std::vector<int> idxs;
    for (int k = 0; k < 3; ++k)
      if (c.constant_tvec_mask[i] & (1 << k)) idxs.push_back(k);
    cfg->SetConstantTvec(c.constant_tvec_image[i] + 1, idxs);
  }
  for (int i = 0; i < c.num_variable_points; ++i) cfg->AddVariablePoint(c.variable_points[i] + 1);
  for (int i = 0; i < c.num_constant_points; ++i) cfg->AddConstantPoint(c.constant_points[i] + 1);
  for (int i = 0; i < c.num_constant_cameras; ++i) cfg->SetConstantCamera(c.constant_cameras[i] + 1);
}

// canonical form of an assembled problem, printed as text lines (sorted)
struct Canon {
  std::vector<std::string> lines;
  void Add(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    lines.emplace_back(buf);
  }
  std::string Str() {
    std::sort(lines.begin(), lines.end());
    std::string out;
    for (const auto& l : lines) out += l + "\n";
    return out;
  }
};

template <class CameraModel>
void ConstantPoseFunctor(const void* f, double pose[7], double line[3]) {
  const auto* p = static_cast<const colmap::BundleAdjustmentConstantPoseLineCostFunction<CameraModel>*>(f);
  pose[0] = p->qw_; pose[1] = p->qx_; pose[2] = p->qy_; pose[3] = p->qz_;
  pose[4] = p->tx_; pose[5] = p->ty_; pose[6] = p->tz_;
  line[0] = p->a_; line[1] = p->b_; line[2] = p->c_;
}
template <class CameraModel>
void VariablePoseFunctor(const void* f, double line[3]) {
  const auto* p = static_cast<const colmap::BundleAdjustmentLineCostFunction<CameraModel>*>(f);
  line[0] = p->a_; line[1] = p->b_; line[2] = p->c_;
}

unsigned GroupMask(const std::vector<size_t>& idxs) {
  unsigned m = 0;
  for (size_t i : idxs) m |= 1u << i;
  return m;
}

}  // namespace

// Writes the two canonical descriptions (NUL-terminated) into ref_out / prod_out (capacity cap).
// Returns 0, 1 if the reference's Solve returned false (no residuals), -1 on overflow.
extern "C" __attribute__((visibility("default"))) int ref_ba_setup_compare(const Scene* scene, const Config* config, char* ref_out,
                                    char* prod_out, size_t cap) {
  // ------------------------------------------------------------------ the reference's SetUp
  colmap::Reconstruction rec;
  BuildReconstruction(*scene, &rec);
  colmap::BundleAdjustmentOptions opt;
  opt.loss_function_type = static_cast<colmap::BundleAdjustmentOptions::LossFunctionType>(config->loss_type);
  opt.loss_function_scale = config->loss_scale;
  opt.refine_focal_length = config->refine_focal_length != 0;
  opt.refine_principal_point = config->refine_principal_point != 0;
  opt.refine_extra_params = config->refine_extra_params != 0;
  opt.refine_extrinsics = config->refine_extrinsics != 0;
  opt.print_summary = false;
  colmap::BundleAdjustmentConfig cfg;
  FillConfig(*config, &cfg);
  colmap::BundleAdjuster adjuster(opt, cfg);
  const bool solved = adjuster.Solve(&rec);

  Canon ref;
  ref.Add("solve %d", solved ? 1 : 0);
  if (solved) {
    const ceres::Problem& pr = *adjuster.problem_;
    std::map<const double*, int> q_of, t_of, x_of, c_of;
    for (auto& kv : rec.images_) {
      q_of[kv.second.Qvec().data()] = static_cast<int>(kv.first);
      t_of[kv.second.Tvec().data()] = static_cast<int>(kv.first);
    }
    for (auto& kv : rec.points3D_) x_of[kv.second.XYZ().data()] = static_cast<int>(kv.first);
    for (auto& kv : rec.cameras_) c_of[kv.second.ParamsData()] = static_cast<int>(kv.first);
    std::set<int> images_variable, images_constant, cams_used, points_used;
    for (const auto& rb : pr.residual_blocks) {
      const auto& sizes = rb.cost->ParameterBlockSizes();
      double line[3], pose[7];
      int image_id = -1, point_id = -1, cam_id = -1;
      if (sizes.size() == 4) {  // (2; 4, 3, 3, k): variable pose
        image_id = q_of.at(rb.blocks[0]);
        if (t_of.at(rb.blocks[1]) != image_id) return -2;
        point_id = x_of.at(rb.blocks[2]);
        cam_id = c_of.at(rb.blocks[3]);
        switch (rec.Camera(cam_id).ModelId()) {
#define CAMERA_MODEL_CASE(CameraModel) \
  case colmap::CameraModel::kModelId:  \
    VariablePoseFunctor<colmap::CameraModel>(rb.cost->FunctorAddress(), line); break;
          CAMERA_MODEL_CASES
#undef CAMERA_MODEL_CASE
        }
        images_variable.insert(image_id);
      } else {  // (2; 3, k): the pose is inside the functor
        point_id = x_of.at(rb.blocks[0]);
        cam_id = c_of.at(rb.blocks[1]);
        switch (rec.Camera(cam_id).ModelId()) {
#define CAMERA_MODEL_CASE(CameraModel) \
  case colmap::CameraModel::kModelId:  \
    ConstantPoseFunctor<colmap::CameraModel>(rb.cost->FunctorAddress(), pose, line); break;
          CAMERA_MODEL_CASES
#undef CAMERA_MODEL_CASE
        }
        for (auto& kv : rec.images_) {  // identify the image by its (normalised) pose
          const auto& im = kv.second;
          bool same = true;
          for (int k = 0; k < 4; ++k) same = same && im.Qvec()(k) == pose[k];
          for (int k = 0; k < 3; ++k) same = same && im.Tvec()(k) == pose[4 + k];
          if (same && static_cast<int>(im.CameraId()) == cam_id) image_id = static_cast<int>(kv.first);
        }
        if (image_id < 0) return -3;
        images_constant.insert(image_id);
      }
      if (static_cast<int>(sizes.back()) != static_cast<int>(rec.Camera(cam_id).NumParams())) return -4;
      cams_used.insert(cam_id);
      points_used.insert(point_id);
      ref.Add("obs image %d point %d line %a %a %a pose_constant %d loss %d %a", image_id, point_id,
              line[0], line[1], line[2], sizes.size() == 4 ? 0 : 1, rb.loss->Kind(), rb.loss->Scale());
    }
    std::set<const double*> constant(pr.constant_blocks.begin(), pr.constant_blocks.end());
    std::map<const double*, const ceres::LocalParameterization*> par;
    for (const auto& p : pr.parameterizations) par[p.first] = p.second.get();
    for (int id : images_variable) {
      if (images_constant.count(id)) return -5;  // an image is either variable or constant
      const auto& im = rec.Image(id);
      unsigned tmask = 0;
      auto it = par.find(im.Tvec().data());
      if (it != par.end())
        for (int k : *it->second->ConstantIndices()) tmask |= 1u << k;
      const bool quat = par.count(im.Qvec().data()) && par.at(im.Qvec().data())->GlobalSize() == 4 &&
                        par.at(im.Qvec().data())->LocalSize() == 3;
      ref.Add("image %d camera %d constant 0 tvec_constant_mask %u quaternion %d", id,
              static_cast<int>(im.CameraId()), tmask, quat ? 1 : 0);
    }
    for (int id : images_constant)
      ref.Add("image %d camera %d constant 1 tvec_constant_mask 0 quaternion 1", id,
              static_cast<int>(rec.Image(id).CameraId()));
    for (int id : points_used)
      ref.Add("point %d constant %d", id, constant.count(rec.Point3D(id).XYZ().data()) ? 1 : 0);
    for (int id : cams_used) {
      const auto& cam = rec.Camera(id);
      const unsigned all = (1u << cam.NumParams()) - 1;
      unsigned variable = all;
      if (constant.count(cam.ParamsData())) {
        variable = 0;
      } else if (par.count(cam.ParamsData())) {
        for (int k : *par.at(cam.ParamsData())->ConstantIndices()) variable &= ~(1u << k);
      }
      ref.Add("camera %d model %d variable_mask %u", id, cam.ModelId(), variable);
    }
    // every camera the reference parameterised took part in a residual block
    for (const auto id : adjuster.camera_ids_)
      if (!cams_used.count(static_cast<int>(id))) return -6;
  }

  // ------------------------------------------------------------------ the product's assembly
  colmap::Reconstruction rec2;
  BuildReconstruction(*scene, &rec2);
  ppsfm::BundleAdjustmentOptions popt;
  popt.loss_function_type = static_cast<ppsfm::BundleAdjustmentOptions::LossFunctionType>(config->loss_type);
  popt.loss_function_scale = config->loss_scale;
  popt.refine_focal_length = config->refine_focal_length != 0;
  popt.refine_principal_point = config->refine_principal_point != 0;
  popt.refine_extra_params = config->refine_extra_params != 0;
  popt.refine_extrinsics = config->refine_extrinsics != 0;
  popt.print_summary = false;
  ppsfm::BundleAdjustmentConfig pcfg;
  FillConfig(*config, &pcfg);
  ppsfm::BundleAdjuster<colmap::Reconstruction> padjuster(popt, pcfg);
  const auto as = padjuster.AssembleOnly(&rec2);

  Canon prod;
  prod.Add("solve %d", as.obs_image.empty() ? 0 : 1);
  if (!as.obs_image.empty()) {
    for (size_t o = 0; o < as.obs_image.size(); ++o) {
      const int ii = as.obs_image[o];
      prod.Add("obs image %d point %d line %a %a %a pose_constant %d loss %d %a",
               static_cast<int>(as.image_ids[ii]), static_cast<int>(as.point_ids[as.obs_point[o]]),
               as.obs_line[3 * o], as.obs_line[3 * o + 1], as.obs_line[3 * o + 2],
               (as.pose_flags[ii] & 1) ? 1 : 0, config->loss_type,
               config->loss_type == 0 ? 0.0 : config->loss_scale);
    }
    for (size_t i = 0; i < as.image_ids.size(); ++i)
      prod.Add("image %d camera %d constant %d tvec_constant_mask %u quaternion 1",
               static_cast<int>(as.image_ids[i]), static_cast<int>(as.camera_ids[as.image_camera[i]]),
               (as.pose_flags[i] & 1) ? 1 : 0, (as.pose_flags[i] & 1) ? 0u : (unsigned)(as.pose_flags[i] >> 1));
    for (size_t p = 0; p < as.point_ids.size(); ++p)
      prod.Add("point %d constant %d", static_cast<int>(as.point_ids[p]), as.point_const[p] ? 1 : 0);
    for (size_t c = 0; c < as.camera_ids.size(); ++c) {
      // VariableIntrinsics of csrc/ba_host.cu: the refined groups unless the camera is constant
      const auto& cam = rec2.Camera(as.camera_ids[c]);
      unsigned variable = 0;
      if (popt.refine_focal_length) variable |= GroupMask(cam.FocalLengthIdxs());
      if (popt.refine_principal_point) variable |= GroupMask(cam.PrincipalPointIdxs());
      if (popt.refine_extra_params) variable |= GroupMask(cam.ExtraParamsIdxs());
      if (as.camera_const[c]) variable = 0;
      prod.Add("camera %d model %d variable_mask %u", static_cast<int>(as.camera_ids[c]),
               cam.ModelId(), variable);
    }
  }
  const std::string a = ref.Str(), b = prod.Str();
  if (a.size() + 1 > cap || b.size() + 1 > cap) return -1;
  std::memcpy(ref_out, a.c_str(), a.size() + 1);
  std::memcpy(prod_out, b.c_str(), b.size() + 1);
  return solved ? 0 : 1;
}

// The linear solver / thread choices of BundleAdjuster::Solve (bundle_adjustment.cc:273-301) are
// properties of Ceres the product has no counterpart for; exposed for the record only.
extern "C" __attribute__((visibility("default"))) void ref_ba_last_solver_options(int* linear_solver_type, int* num_threads,
                                           int* max_num_iterations) {
  const ceres::Solver::Options& o = ceres::LastSolveOptions();
  *linear_solver_type = static_cast<int>(o.linear_solver_type);
  *num_threads = o.num_threads;
  *max_num_iterations = o.max_num_iterations;
}
