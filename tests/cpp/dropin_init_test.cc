// dropin_init_test.cc — a caller written against the REFERENCE's own src/init/initializer.h
// (colmap::FeatureLines, colmap::init::InitOptions, colmap::init::Pose), linked with the drop-in
// definition of privacy_preserving_sfm_b200/cpp/dropin/init_initializer.cc (host build: no GPU,
// no library).  Scene from stdin as doubles: lines[4][n][3], aligned[4][n], gravity[4][3];
// prints ok, the inlier ratio and the four poses (row-major 3x4) as hex doubles.
//   dropin_init_test <n>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "feature/types.h"
#include "init/initializer.h"

int main(int argc, char** argv) {
  if (argc != 2) return 2;
  const size_t n = static_cast<size_t>(std::atoll(argv[1]));
  std::vector<double> buf(4 * n * 3 + 4 * n + 12);
  if (std::fread(buf.data(), sizeof(double), buf.size(), stdin) != buf.size()) return 2;
  const double* l = buf.data();
  const double* a = l + 4 * n * 3;
  const double* g = a + 4 * n;
  std::vector<colmap::FeatureLines> lines(4);
  std::vector<Eigen::Vector3d> gravity;
  for (int i = 0; i < 4; ++i) {
    for (size_t j = 0; j < n; ++j) {
      const double* v = l + 3 * (i * n + j);
      lines[i].emplace_back(Eigen::Vector3d(v[0], v[1], v[2]), a[i * n + j] != 0.0);
    }
    gravity.emplace_back(g[3 * i], g[3 * i + 1], g[3 * i + 2]);
  }
  colmap::init::InitOptions options;  // the reference's defaults (initializer.h:49-58)
  std::vector<colmap::init::Pose> poses;
  double inlier_ratio = 0.0;
  const bool ok = colmap::init::initialize_reconstruction(lines, gravity, options, &poses, &inlier_ratio);
  std::printf("%d %a %zu\n", ok ? 1 : 0, inlier_ratio, poses.size());
  for (const auto& P : poses) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) std::printf("%a ", P(r, c));
    std::printf("\n");
  }
  return 0;
}
