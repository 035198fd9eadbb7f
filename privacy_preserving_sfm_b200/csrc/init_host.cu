// init_host.cu — C-ABI of the four-view initialisation (host code; SURVEY.md §8 A17/A18, the
// "plumbing" configuration 1 of BASELINE.json).  The implementation lives in the header-only
// cpp/ppsfm_init.h + cpp/ppsfm_lomsac.h so that the C++ adaptor and the reference-driver check
// (oracle/ref/ref_init.cc) use the very same estimators.
#include <cstring>

#include "../../include/ppsfm_b200.h"
#include "../cpp/ppsfm_init.h"
#include "common.h"
#include "init_kernels.h"

namespace {

int InitializeReconstruction(ppsfm_ctx* ctx, const double* lines, const uint8_t* aligned, size_t n,
                             const double* gravity, const ppsfm_init_options* options,
                             double* poses_out, double* inlier_ratio, ppsfm_init_report* report,
                             int64_t* gpu_launches) {
  if (!lines || !aligned || !gravity || !options || !poses_out || !inlier_ratio)
    return PPSFM_ERR_INVALID;
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
  opt.min_tri_angle = options->min_tri_angle;
  opt.min_num_inliers = options->min_num_inliers;
  opt.max_error = options->max_error;
  std::vector<Pose> poses;
  InitReport rep;
  const char* err = nullptr;
  bool ok;
  if (ctx) {  // candidate models scored on the GPU (init_kernels.cu)
    cudaSetDevice(ctx->device);
    ppsfm::GpuScorerFactory scorers(ctx);
    ok = initialize_reconstruction(img, g, opt, &poses, inlier_ratio, &rep, &err, &scorers);
    if (gpu_launches) *gpu_launches = scorers.launches();
    if (scorers.error() != cudaSuccess)
      return ppsfm::fail(ctx, PPSFM_ERR_CUDA, "four-view initialisation: %s",
                         cudaGetErrorString(scorers.error()));
  } else {
    ok = initialize_reconstruction(img, g, opt, &poses, inlier_ratio, &rep, &err);
  }
  if (report) {
    report->num_aligned = rep.num_aligned;
    report->num_unaligned = rep.num_unaligned;
    report->inliers_2d = rep.inliers_2d;
    report->inliers_3d = rep.inliers_3d;
    report->iterations_2d = rep.iterations_2d;
    report->iterations_3d = rep.iterations_3d;
    report->mean_tri_angle_deg = rep.mean_tri_angle_deg;
  }
  if (err) return ctx ? ppsfm::fail(ctx, PPSFM_ERR_INVALID, "%s", err) : PPSFM_ERR_INVALID;
  if (poses.size() == 4) std::memcpy(poses_out, poses.data(), sizeof(double) * 48);
  return ok ? PPSFM_OK : PPSFM_NO_SOLUTION;
}

}  // namespace

extern "C" {

void ppsfm_init_options_default(ppsfm_init_options* o) {
  if (!o) return;
  const ppsfm::init::InitOptions d;  // src/init/initializer.h:49-58
  o->min_tri_angle = d.min_tri_angle;
  o->min_num_inliers = d.min_num_inliers;
  o->max_error = d.max_error;
}

// init::initialize_reconstruction (src/init/initializer.h:103-108) on the host.
int ppsfm_initialize_reconstruction(const double* lines, const uint8_t* aligned, size_t n,
                                    const double* gravity, const ppsfm_init_options* options,
                                    double* poses_out, double* inlier_ratio,
                                    ppsfm_init_report* report) {
  return InitializeReconstruction(nullptr, lines, aligned, n, gravity, options, poses_out,
                                  inlier_ratio, report, nullptr);
}

// The same with the candidate models of both LO-MSAC loops scored on the GPU: identical results
// (the scores are bit-identical to the host's), the host keeps the control flow.
int ppsfm_initialize_reconstruction_gpu(ppsfm_ctx* ctx, const double* lines,
                                        const uint8_t* aligned, size_t n, const double* gravity,
                                        const ppsfm_init_options* options, double* poses_out,
                                        double* inlier_ratio, ppsfm_init_report* report,
                                        int64_t* gpu_launches) {
  if (!ctx) return PPSFM_ERR_INVALID;
  return InitializeReconstruction(ctx, lines, aligned, n, gravity, options, poses_out,
                                  inlier_ratio, report, gpu_launches);
}

// Test hook: la::qr_solve (the generic host routine) and hd::qr_solve_fixed (the fixed-size routine
// the triangulations and the GPU kernel use) on one m x n system, (m, n) = (3, 2) or (4, 3).
int ppsfm_init_test_qr(const double* A, int m, int n, const double* b, double* x_generic,
                       double* x_fixed) {
  if (!A || !b || !x_generic || !x_fixed) return PPSFM_ERR_INVALID;
  if (!((m == 3 && n == 2) || (m == 4 && n == 3))) return PPSFM_ERR_INVALID;
  std::vector<double> Av(A, A + m * n), bv(b, b + m);
  ppsfm::init::la::qr_solve(Av, m, n, bv, x_generic);
  double Af[12], bf[4];
  std::memcpy(Af, A, sizeof(double) * m * n);
  std::memcpy(bf, b, sizeof(double) * m);
  if (m == 3) ppsfm::init::hd::qr_solve_fixed<3, 2>(Af, bf, x_fixed);
  else ppsfm::init::hd::qr_solve_fixed<4, 3>(Af, bf, x_fixed);
  return PPSFM_OK;
}

}  // extern "C"
