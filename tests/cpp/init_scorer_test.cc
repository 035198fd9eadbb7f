// CPU check of the batched-scoring hooks of the four-view initialisation (cpp/ppsfm_init.h,
// cpp/ppsfm_lomsac.h): a HOST implementation of init::BatchScorer — the loop the GPU kernel
// csrc/init_kernels.cu runs, on the shared arithmetic of cpp/ppsfm_init_math.h — is plugged
// into initialize_reconstruction; lazy models, ScoreModels and Materialize must reproduce the
// plain run bit for bit.  argv: n n_aligned n_outliers seed; the scene comes from stdin as
// doubles: lines[4][n][3], aligned[4][n], gravity[4][3].
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "ppsfm_init.h"

using namespace ppsfm::init;

namespace {

class HostScorer : public BatchScorer {
 public:
  HostScorer(const double* const* obs, int n, bool is3d) : n_(n), is3d_(is3d) {
    const size_t per = static_cast<size_t>(n) * (is3d ? 3 : 2);
    for (int v = 0; v < 4; ++v) obs_[v].assign(obs[v], obs[v] + per);
  }
  bool ScoreFourView2d(const double* cams, int m, double thr, bool thr_first,
                       double* scores) const override {
    if (is3d_) return false;
    for (int k = 0; k < m; ++k) {
      const double* cam = cams + 24 * k;
      double s = 0.0;
      for (int i = 0; i < n_; ++i) {
        const double *x0 = &obs_[0][2 * i], *x1 = &obs_[1][2 * i], *x2 = &obs_[2][2 * i],
                     *x3 = &obs_[3][2 * i];
        double X[2];
        hd::triangulate2d_point(cam, cam + 6, cam + 12, x0, x1, x2, X);
        const double e = hd::fourview2d_error(cam, x0, x1, x2, x3, X);
        s += thr_first ? ((e < thr) ? e : thr) : ((thr < e) ? thr : e);
      }
      scores[k] = s;
    }
    ++calls;
    return true;
  }
  bool ScorePlanarOffset(const double* cams, int m, double thr, bool thr_first,
                         double* scores) const override {
    if (!is3d_) return false;
    for (int k = 0; k < m; ++k) {
      const double* cam = cams + 48 * k;
      double s = 0.0;
      for (int i = 0; i < n_; ++i) {
        const double *l0 = &obs_[0][3 * i], *l1 = &obs_[1][3 * i], *l2 = &obs_[2][3 * i],
                     *l3 = &obs_[3][3 * i];
        double X[3];
        hd::triangulate3d_point(cam, l0, l1, l2, l3, X);
        const double e = hd::planar_offset_error(cam, l0, l1, l2, l3, X);
        s += thr_first ? ((e < thr) ? e : thr) : ((thr < e) ? thr : e);
      }
      scores[k] = s;
    }
    ++calls;
    return true;
  }
  mutable long calls = 0;

 private:
  int n_;
  bool is3d_;
  std::vector<double> obs_[4];
};

class HostFactory : public BatchScorerFactory {
 public:
  const BatchScorer* FourView2d(const double* const* x, int n) override {
    made.emplace_back(new HostScorer(x, n, false));
    return made.back().get();
  }
  const BatchScorer* PlanarOffset(const double* const* lines, int n) override {
    made.emplace_back(new HostScorer(lines, n, true));
    return made.back().get();
  }
  std::vector<std::unique_ptr<HostScorer>> made;
};

}  // namespace

int main(int argc, char** argv) {
  if (argc != 2) return 2;
  const int n = std::atoi(argv[1]);
  std::vector<double> buf(static_cast<size_t>(4) * n * 3 + 4 * n + 12);
  if (std::fread(buf.data(), sizeof(double), buf.size(), stdin) != buf.size()) return 3;
  std::vector<ImageLines> img(4);
  std::vector<Vec3> g(4);
  for (int i = 0; i < 4; ++i) {
    img[i].line.resize(n);
    img[i].aligned.resize(n);
    for (int j = 0; j < n; ++j) {
      const double* l = &buf[3 * (static_cast<size_t>(i) * n + j)];
      img[i].line[j] = Vec3{l[0], l[1], l[2]};
      img[i].aligned[j] = buf[static_cast<size_t>(12) * n + static_cast<size_t>(i) * n + j] != 0.0;
    }
    const double* gg = &buf[static_cast<size_t>(16) * n + 3 * i];
    g[i] = Vec3{gg[0], gg[1], gg[2]};
  }
  InitOptions opt;
  std::vector<Pose> p0, p1;
  double r0 = 0, r1 = 0;
  InitReport rep0, rep1;
  const bool ok0 = initialize_reconstruction(img, g, opt, &p0, &r0, &rep0);
  HostFactory factory;
  const bool ok1 = initialize_reconstruction(img, g, opt, &p1, &r1, &rep1, nullptr, &factory);
  long calls = 0;
  for (const auto& s : factory.made) calls += s->calls;
  bool same = ok0 == ok1 && r0 == r1 && p0.size() == p1.size() &&
              rep0.inliers_2d == rep1.inliers_2d && rep0.inliers_3d == rep1.inliers_3d &&
              rep0.iterations_2d == rep1.iterations_2d && rep0.iterations_3d == rep1.iterations_3d;
  for (size_t i = 0; same && i < p0.size(); ++i)
    same = std::memcmp(&p0[i], &p1[i], sizeof(Pose)) == 0;
  std::printf("ok %d %d ratio %.17g %.17g scorer_calls %ld identical %d\n", ok0, ok1, r0, r1,
              calls, same ? 1 : 0);
  return same && calls > 0 ? 0 : 1;
}
