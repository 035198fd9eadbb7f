// dropin_pose_test.cc — a caller written against the REFERENCE's own headers
// (src/estimators/pose.h, src/feature/types.h, src/base/camera.h, src/optim/ransac.h), the way
// IncrementalMapper::RegisterNextImage calls them (src/sfm/incremental_mapper.cc:673-735), linked
// with the drop-in definitions of privacy_preserving_sfm_b200/cpp/dropin/estimators_pose_lines.cc
// and libppsfm_b200.so instead of the reference's pose.cc.  Built by tests/cpp/build_dropin.sh in
// the container that has /root/reference (Eigen / Ceres / glog / Boost: stand-in headers of
// oracle/ref/shim); the binary travels to the GPU box.
//   dropin_pose_test <in.bin> <out.bin>      (file layout of adaptor_test's `pose` mode)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "base/camera.h"
#include "base/camera_models.h"
#include "estimators/pose.h"
#include "feature/types.h"

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int64_t n64 = 0;
  if (std::fread(&n64, sizeof(n64), 1, f) != 1) return 2;
  const size_t n = static_cast<size_t>(n64);
  std::vector<double> l(3 * n), p(3 * n), a(n);
  if (std::fread(l.data(), 8, 3 * n, f) != 3 * n || std::fread(p.data(), 8, 3 * n, f) != 3 * n ||
      std::fread(a.data(), 8, n, f) != n)
    return 2;
  std::fclose(f);

  colmap::FeatureLines lines2D;
  std::vector<Eigen::Vector3d> points3D, line_vecs;
  for (size_t i = 0; i < n; ++i) {
    const Eigen::Vector3d line(l[3 * i], l[3 * i + 1], l[3 * i + 2]);
    lines2D.emplace_back(line, a[i] != 0.0);
    line_vecs.push_back(line);
    points3D.emplace_back(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  }

  // the mapper's settings (incremental_mapper.cc:673-681)
  colmap::RANSACOptions ransac_options;
  ransac_options.max_error = 12.0 / 1000.0;
  ransac_options.min_inlier_ratio = 0.25;
  ransac_options.confidence = 0.99999;
  ransac_options.min_num_trials = 100;
  ransac_options.max_num_trials = 10000;

  Eigen::Vector4d qvec;
  Eigen::Vector3d tvec;
  size_t num_inliers = 0;
  std::vector<char> inlier_mask;
  const bool ok = colmap::EstimateAbsolutePoseFromLines(ransac_options, lines2D, points3D, &qvec,
                                                        &tvec, &num_inliers, &inlier_mask);

  colmap::Camera camera;
  camera.SetModelId(colmap::PinholeCameraModel::model_id);
  camera.SetWidth(1000);
  camera.SetHeight(1000);
  camera.SetParams({1000.0, 1000.0, 500.0, 500.0});
  colmap::AbsolutePoseRefinementOptions refine_options;
  refine_options.print_summary = false;
  Eigen::Vector4d qref = qvec;
  Eigen::Vector3d tref = tvec;
  bool ok_ref = false;
  if (ok) {
    ok_ref = colmap::RefineAbsolutePoseFromLines(refine_options, inlier_mask, line_vecs, points3D,
                                                 &qref, &tref, &camera);
  }

  std::vector<double> out;
  out.push_back(ok ? 1.0 : 0.0);
  out.push_back(static_cast<double>(num_inliers));
  out.push_back(ok_ref ? 1.0 : 0.0);
  out.push_back(0.0);
  for (int k = 0; k < 4; ++k) out.push_back(qvec(k));
  for (int k = 0; k < 3; ++k) out.push_back(tvec(k));
  for (int k = 0; k < 4; ++k) out.push_back(qref(k));
  for (int k = 0; k < 3; ++k) out.push_back(tref(k));
  for (size_t i = 0; i < n; ++i) out.push_back(i < inlier_mask.size() && inlier_mask[i] ? 1.0 : 0.0);
  FILE* g = std::fopen(argv[2], "wb");
  if (!g) return 2;
  std::fwrite(out.data(), 8, out.size(), g);
  std::fclose(g);
  return 0;
}
