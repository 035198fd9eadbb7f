// init_initializer.cc — the drop-in for the reference's four-view initialisation entry point.
//
// Compiled inside the reference tree in place of the body of
//     bool init::initialize_reconstruction(...)    src/init/initializer.cc:57-215
// (declared in src/init/initializer.h:103-108, included here, so the compiler checks the
// signature and the InitOptions / Pose / FeatureLines types against the reference's own).
// Caller: IncrementalMapper::RegisterInitialLineImages (src/sfm/incremental_mapper.cc:192-560).
// By default the control flow, the minimal solvers AND the model scoring run on the host
// (cpp/ppsfm_init.h, header-only: no library needed); with -DPPSFM_INIT_ON_GPU the candidate
// models of both LO-MSAC loops are scored on the GPU through libppsfm_b200.so
// (ppsfm_initialize_reconstruction_gpu; identical results).
// Contract violations the reference CHECK-aborts on (size mismatch, an aligned line that is not
// parallel to gravity, initializer.cc:83) abort here too.
#include "init/initializer.h"  // the reference's declarations

#include <cstdio>
#include <cstdlib>

#ifdef PPSFM_INIT_ON_GPU
#ifndef PPSFM_WITH_EIGEN
#define PPSFM_WITH_EIGEN
#endif
#include "ppsfm_adaptor.h"
#else
#include "ppsfm_init.h"
#endif

namespace colmap {
namespace init {

bool initialize_reconstruction(const std::vector<FeatureLines>& lines,
                               const std::vector<Eigen::Vector3d>& gravity,
                               const InitOptions& options, std::vector<Pose>* output,
                               double* inlier_ratio) {
  if (lines.size() != 4 || gravity.size() != 4) {
    std::fprintf(stderr, "Check failed: four images are required\n");
    std::abort();
  }
  const size_t n = lines[0].size();
#ifdef PPSFM_INIT_ON_GPU
  std::vector<double> l(4 * n * 3), g(12), poses(48);
  std::vector<uint8_t> a(4 * n);
  for (int i = 0; i < 4; ++i) {
    if (lines[i].size() != n) std::abort();
    for (size_t j = 0; j < n; ++j) {
      for (int k = 0; k < 3; ++k) l[3 * (i * n + j) + k] = lines[i][j].Line()(k);
      a[i * n + j] = lines[i][j].IsAligned() ? 1 : 0;
    }
    for (int k = 0; k < 3; ++k) g[3 * i + k] = gravity[i](k);
  }
  const ppsfm_init_options o = {options.min_tri_angle, options.min_num_inliers, options.max_error};
  const int rc = ppsfm_initialize_reconstruction_gpu(ppsfm::ThreadContext(), l.data(), a.data(), n,
                                                     g.data(), &o, poses.data(), inlier_ratio,
                                                     nullptr, nullptr);
  if (rc < 0) {
    std::fprintf(stderr, "Check failed: ppsfm_initialize_reconstruction_gpu rc=%d\n", rc);
    std::abort();
  }
  output->clear();
  if (rc != 0) return false;
  for (int i = 0; i < 4; ++i) {
    Pose P;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) P(r, c) = poses[12 * i + 4 * r + c];  // row-major on the ABI
    output->push_back(P);
  }
  return true;
#else
  namespace pi = ppsfm::init;
  std::vector<pi::ImageLines> img(4);
  std::vector<pi::Vec3> g(4);
  for (int i = 0; i < 4; ++i) {
    g[i] = pi::Vec3{gravity[i](0), gravity[i](1), gravity[i](2)};
    img[i].line.resize(lines[i].size());
    img[i].aligned.resize(lines[i].size());
    for (size_t j = 0; j < lines[i].size(); ++j) {
      const Eigen::Vector3d& v = lines[i][j].Line();
      img[i].line[j] = pi::Vec3{v(0), v(1), v(2)};
      img[i].aligned[j] = lines[i][j].IsAligned() ? 1 : 0;
    }
  }
  pi::InitOptions o;
  o.min_tri_angle = options.min_tri_angle;
  o.min_num_inliers = options.min_num_inliers;
  o.max_error = options.max_error;
  std::vector<pi::Pose> poses;
  const char* error = nullptr;
  const bool ok = pi::initialize_reconstruction(img, g, o, &poses, inlier_ratio, nullptr, &error);
  if (error != nullptr) {
    std::fprintf(stderr, "Check failed: %s\n", error);
    std::abort();
  }
  output->clear();
  for (const pi::Pose& p : poses) {
    Pose P;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) P(r, c) = p.m[r][c];
    output->push_back(P);
  }
  return ok;
#endif
}

}  // namespace init
}  // namespace colmap
