// ref_init.cc — TEST INFRASTRUCTURE.  Runs the four-view initialisation with the REFERENCE's own
// LO-MSAC driver: ransac_lib::LocallyOptimizedMSAC from /root/reference/lib/RansacLib (header-only,
// compiles stand-alone; included from where it lies, never copied) around the product's estimator
// classes (privacy_preserving_sfm_b200/cpp/ppsfm_init.h).  tests/test_init.py requires the result
// to be bit-identical to the product's ppsfm::LocallyOptimizedMSAC, which pins SURVEY.md §8 row
// A17 against real reference code.  Built by oracle/build_ref.sh into oracle/_ref/libref_init.so.
#include <RansacLib/ransac.h>

#include <cstdint>
#include <cstring>

#include "../../privacy_preserving_sfm_b200/cpp/ppsfm_init.h"

namespace {
struct RefLomsac {
  using Options = ransac_lib::LORansacOptions;
  using Stats = ransac_lib::RansacStatistics;
  template <class M, class MV, class S>
  using Driver = ransac_lib::LocallyOptimizedMSAC<M, MV, S>;
};
}  // namespace

extern "C" int ref_initialize_reconstruction(const double* lines, const uint8_t* aligned, size_t n,
                                             const double* gravity, const double* options3,
                                             double* poses_out, double* inlier_ratio,
                                             double* report7) {
  using namespace ppsfm::init;
  std::vector<ImageLines> img(4);
  std::vector<Vec3> g(4);
  for (int i = 0; i < 4; ++i) {
    g[i] = Vec3{gravity[3 * i], gravity[3 * i + 1], gravity[3 * i + 2]};
    img[i].line.resize(n);
    img[i].aligned.assign(aligned + i * n, aligned + (i + 1) * n);
    for (size_t j = 0; j < n; ++j) {
      const double* l = lines + 3 * (i * n + j);
      img[i].line[j] = Vec3{l[0], l[1], l[2]};
    }
  }
  InitOptions opt;
  opt.min_tri_angle = options3[0];
  opt.min_num_inliers = options3[1];
  opt.max_error = options3[2];
  std::vector<Pose> poses;
  InitReport rep;
  const char* err = nullptr;
  const bool ok = initialize_reconstruction_t<RefLomsac>(img, g, opt, &poses, inlier_ratio, &rep, &err);
  if (report7) {
    report7[0] = rep.num_aligned; report7[1] = rep.num_unaligned;
    report7[2] = rep.inliers_2d; report7[3] = rep.inliers_3d;
    report7[4] = rep.iterations_2d; report7[5] = rep.iterations_3d;
    report7[6] = rep.mean_tri_angle_deg;
  }
  if (err) return -1;
  if (poses.size() == 4) std::memcpy(poses_out, poses.data(), sizeof(double) * 48);
  return ok ? 0 : 1;
}
