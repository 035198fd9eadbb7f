// ppsfm_oracle.cc — CPU oracle for the line-lifted absolute-pose RANSAC path.
//
// TEST INFRASTRUCTURE ONLY (see ppsfm_oracle.h).  Dependency-free C++17 restatement of the
// reference's CPU path; every function cites the reference file:line it follows
// (paths relative to the upstream repository colmap/privacy_preserving_sfm).
//
// Build: g++ -O2 -ffp-contract=off -shared -fPIC   (no FMA contraction: the reference is built
// for baseline x86-64, CMakeLists.txt has no -march flag, so every mul/add rounds separately).
//
// PARITY: scoring / support / sampler / RANSAC loop follow in-tree reference code 1:1.
// P6L + re3q3 call Eigen in the reference (absent here): elimination, LU and the companion-matrix
// eigenvalue step are restated from the published algorithms (Francis double-shift QR on the
// Hessenberg companion matrix, EISPACK `hqr` lineage, as used by Eigen::RealSchur) — pinned by the
// reference's known-answer properties only; bit-level parity with an Eigen build is UNPINNED
// (eigen_restated.h).  Everything else is pinned bit for bit against the reference's own sources
// compiled here against stand-ins (oracle/build_ref.sh, tests/test_ref_p6l.py).
// The two `rand()`-driven degenerate fallbacks (absolute_pose.cc:128-134, re3q3.h:39-64) use a
// FIXED generic matrix instead of C rand() so that results are reproducible.

#include "ppsfm_oracle.h"

#include "eigen_restated.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// util/random.{h,cc}: one std::mt19937, default seed 0 (random.h:46), lazily created (:90-92).
// ---------------------------------------------------------------------------------------------
std::mt19937* g_prng = nullptr;

std::mt19937& Prng() {
  if (g_prng == nullptr) g_prng = new std::mt19937(0u);
  return *g_prng;
}

// util/random.h:88-97 RandomInteger<uint32_t>
uint32_t RandomInteger(uint32_t lo, uint32_t hi) {
  std::uniform_int_distribution<uint32_t> distribution(lo, hi);
  return distribution(Prng());
}

// optim/random_sampler.cc:40-62 — persistent permutation, partial Fisher-Yates of the first k.
struct RandomSampler {
  size_t k;
  std::vector<size_t> idxs;
  explicit RandomSampler(size_t num_samples) : k(num_samples) {}
  void Initialize(size_t total) {
    idxs.resize(total);
    std::iota(idxs.begin(), idxs.end(), 0);
  }
  // util/random.h:120-128 Shuffle(num_to_shuffle, &elems)
  void Sample(size_t* out) {
    const uint32_t last = static_cast<uint32_t>(idxs.size() - 1);
    for (uint32_t i = 0; i < static_cast<uint32_t>(k); ++i) {
      const uint32_t j = RandomInteger(i, last);
      std::swap(idxs[i], idxs[j]);
    }
    for (size_t i = 0; i < k; ++i) out[i] = idxs[i];
  }
};

// Det3, SolvePartialPiv3, Hqr8, Poly8Roots: the stand-ins for the Eigen calls of the reference
// live in eigen_restated.h (shared with the Eigen stand-in the real reference sources compile
// against, oracle/ref/shim).
using namespace eigen_restated;


// The hidden-variable resultant of re3q3 (re3q3.h:84-150, :177-188) as generated straight-line
// code: one statement per IEEE operation of the reference's expressions, in their order.
#define RE3Q3_FN inline
#include "re3q3_resultant.inc"
#undef RE3Q3_FN


// ---------------------------------------------------------------------------------------------
// re3q3  (lib/re3q3/re3q3/re3q3.h:16-200)
// coeffs[k][m], monomial order x^2 xy xz y^2 yz z^2 x y z 1.
// ---------------------------------------------------------------------------------------------

// Fixed generic change of variables replacing the reference's rand()-driven one (re3q3.h:41-42).
const double kVarChangeA[3][4] = {
    // rotation of unit quaternion (0.42,-0.31,0.56,0.64)/|.| ; translation (0.3,-0.5,0.8)/|.|
    {-0.45264637943155561, -0.88862107060359552, 0.073917846740986226, 0.30304576336566319},
    {0.19122225569950785, -0.015767801546650473, 0.98142010645776834, -0.5050762722761053},
    {-0.8709450637742292, 0.45837099527970276, 0.1770613638646179, 0.80812203564176865}};

int Re3q3Impl(double coeffs[3][10], double solutions[3][8], bool try_var_change) {
  // Choose the elimination variable by the largest |det| of the quadratic block (:19-37).
  double Ax[3][3], Ay[3][3], Az[3][3];
  for (int k = 0; k < 3; ++k) {
    Ax[k][0] = coeffs[k][3]; Ax[k][1] = coeffs[k][5]; Ax[k][2] = coeffs[k][4];  // y^2 z^2 yz
    Ay[k][0] = coeffs[k][0]; Ay[k][1] = coeffs[k][5]; Ay[k][2] = coeffs[k][2];  // x^2 z^2 xz
    Az[k][0] = coeffs[k][3]; Az[k][1] = coeffs[k][0]; Az[k][2] = coeffs[k][1];  // y^2 x^2 yx
  }
  const double detx = std::fabs(Det3(Ax));
  const double dety = std::fabs(Det3(Ay));
  const double detz = std::fabs(Det3(Az));
  int elim_var = 1;
  double det = detx;
  if (det < dety) { det = dety; elim_var = 2; }
  if (det < detz) { det = detz; elim_var = 3; }

  if (try_var_change && det < 1e-10) {
    // Affine change of variables v = A[:, :3] v' + A[:, 3]  (:39-64): every quadric
    // q(v) = [v;1]^T Q [v;1] becomes q'(v') with Q' = G^T Q G, G = [A; 0 0 0 1].
    const double(*A)[4] = kVarChangeA;
    double G[4][4] = {{A[0][0], A[0][1], A[0][2], A[0][3]},
                      {A[1][0], A[1][1], A[1][2], A[1][3]},
                      {A[2][0], A[2][1], A[2][2], A[2][3]},
                      {0.0, 0.0, 0.0, 1.0}};
    double c2[3][10];
    for (int k = 0; k < 3; ++k) {
      const double* c = coeffs[k];
      const double Q[4][4] = {{c[0], 0.5 * c[1], 0.5 * c[2], 0.5 * c[6]},
                              {0.5 * c[1], c[3], 0.5 * c[4], 0.5 * c[7]},
                              {0.5 * c[2], 0.5 * c[4], c[5], 0.5 * c[8]},
                              {0.5 * c[6], 0.5 * c[7], 0.5 * c[8], c[9]}};
      double QG[4][4], Qp[4][4];
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          double s = 0.0;
          for (int l = 0; l < 4; ++l) s = s + Q[i][l] * G[l][j];
          QG[i][j] = s;
        }
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
          double s = 0.0;
          for (int l = 0; l < 4; ++l) s = s + G[l][i] * QG[l][j];
          Qp[i][j] = s;
        }
      c2[k][0] = Qp[0][0];
      c2[k][1] = Qp[0][1] + Qp[1][0];
      c2[k][2] = Qp[0][2] + Qp[2][0];
      c2[k][3] = Qp[1][1];
      c2[k][4] = Qp[1][2] + Qp[2][1];
      c2[k][5] = Qp[2][2];
      c2[k][6] = Qp[0][3] + Qp[3][0];
      c2[k][7] = Qp[1][3] + Qp[3][1];
      c2[k][8] = Qp[2][3] + Qp[3][2];
      c2[k][9] = Qp[3][3];
    }
    const int n = Re3q3Impl(c2, solutions, false);
    for (int s = 0; s < n; ++s) {  // revert (:60-62)
      const double x = solutions[0][s], y = solutions[1][s], z = solutions[2][s];
      for (int i = 0; i < 3; ++i)
        solutions[i][s] = A[i][0] * x + A[i][1] * y + A[i][2] * z + A[i][3];
    }
    return n;
  }

  // Column order so that the eliminated variable plays the role of "x" (:68-80).
  // P columns: [x^2, xy, xz, x, y, z, 1] in the permuted naming.
  static const int kCols[3][7] = {{0, 1, 2, 6, 7, 8, 9},   // eliminate x
                                  {3, 1, 4, 7, 6, 8, 9},   // eliminate y (x<->y)
                                  {5, 4, 2, 8, 7, 6, 9}};  // eliminate z (x<->z)
  double A[3][3], P[3][7];
  for (int k = 0; k < 3; ++k) {
    for (int j = 0; j < 3; ++j)
      A[k][j] = (elim_var == 1) ? Ax[k][j] : (elim_var == 2) ? Ay[k][j] : Az[k][j];
    for (int j = 0; j < 7; ++j) P[k][j] = coeffs[k][kCols[elim_var - 1][j]];
  }
  SolvePartialPiv3<7>(A, P);  // P = -A.lu().solve(P)  (:71-79)
  for (int k = 0; k < 3; ++k)
    for (int j = 0; j < 7; ++j) P[k][j] = -P[k][j];

  // a11 ... a313 and c(0) ... c(8) = det M(x), literally the reference's expressions (:84-150)
  double a[33], c[9];
  re3q3_resultant(P, a, c);

  double re[8], im[8];
  Poly8Roots(c, re, im);

  int root_cnt = 0;
  for (int i = 0; i < 8; ++i) {
    if (std::fabs(im[i]) > 1e-8) continue;  // (:173)
    const double xs1 = re[i];
    solutions[0][root_cnt] = xs1;  // (:177-188)
    re3q3_backsubstitute(a, xs1, &solutions[1][root_cnt], &solutions[2][root_cnt]);
    ++root_cnt;
  }
  if (elim_var == 2) {
    for (int s = 0; s < root_cnt; ++s) std::swap(solutions[0][s], solutions[1][s]);
  } else if (elim_var == 3) {
    for (int s = 0; s < root_cnt; ++s) std::swap(solutions[0][s], solutions[2][s]);
  }
  return root_cnt;
}

// ---------------------------------------------------------------------------------------------
// P6L  (src/estimators/absolute_pose.cc:46-162)
// ---------------------------------------------------------------------------------------------
// Fixed substitute for Eigen's setRandom() in the degenerate-translation-block branch (:128-134).
const double kMixA[3][3] = {{0.680375, -0.211234, 0.566198},
                            {0.596880, 0.823295, -0.604897},
                            {-0.329554, 0.536459, -0.444451}};

int P6LEstimate(const double lines[6][3], const uint8_t aligned[6], const double points[6][3],
                double models[8][12]) {
  bool all_aligned = true;
  for (int i = 0; i < 6; ++i) all_aligned = all_aligned && (aligned[i] != 0);
  if (all_aligned) return 0;  // (:87-97)

  // l^T t + kron(X^T, l^T) vec(R) = 0  (:101-123), vec(R) column-major.
  double tt[3][9], Rc[3][9];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k)
      for (int j = 0; j < 3; ++j) {
        tt[i][3 * k + j] = points[i][k] * lines[i][j];
        Rc[i][3 * k + j] = points[i + 3][k] * lines[i + 3][j];
      }

  // B = [l0 l1 l2] (columns); |det| test (:126-134)
  double B[3][3], L1[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      B[r][c] = lines[c][r];
      L1[r][c] = lines[c + 3][r];
    }
  const double det_tt = std::fabs(Det3(B));
  if (det_tt < 1e-10) {
    double tt2[3][9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 9; ++j) {
        const double s = kMixA[i][0] * Rc[0][j] + kMixA[i][1] * Rc[1][j] + kMixA[i][2] * Rc[2][j];
        tt2[i][j] = tt[i][j] + s;
      }
    std::memcpy(tt, tt2, sizeof(tt));
    double B2[3][3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        const double s = L1[r][0] * kMixA[c][0] + L1[r][1] * kMixA[c][1] + L1[r][2] * kMixA[c][2];
        B2[r][c] = B[r][c] + s;
      }
    std::memcpy(B, B2, sizeof(B));
  }

  // tt <- B^-T tt  (:137)
  double Bt[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Bt[r][c] = B[c][r];
  SolvePartialPiv3<9>(Bt, tt);
  // Rc <- Rc - L1^T tt  (:138)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 9; ++j) {
      const double s = L1[0][i] * tt[0][j] + L1[1][i] * tt[1][j] + L1[2][i] * tt[2][j];
      Rc[i][j] = Rc[i][j] - s;
    }

  // rotation_to_e3q3 (:46-62): Cayley parametrisation turns each linear form into a quadric.
  double coeffs[3][10];
  for (int k = 0; k < 3; ++k) {
    const double* r = Rc[k];
    coeffs[k][0] = r[0] - r[4] - r[8];
    coeffs[k][1] = 2 * r[1] + 2 * r[3];
    coeffs[k][2] = 2 * r[2] + 2 * r[6];
    coeffs[k][3] = r[4] - r[0] - r[8];
    coeffs[k][4] = 2 * r[5] + 2 * r[7];
    coeffs[k][5] = r[8] - r[4] - r[0];
    coeffs[k][6] = 2 * r[5] - 2 * r[7];
    coeffs[k][7] = 2 * r[6] - 2 * r[2];
    coeffs[k][8] = 2 * r[1] - 2 * r[3];
    coeffs[k][9] = r[0] + r[4] + r[8];
  }

  double sols[3][8];
  const int n_sols = Re3q3Impl(coeffs, sols, true);

  for (int s = 0; s < n_sols; ++s) {
    // cayley_param (:64-75), R row-major here
    const double c0 = sols[0][s], c1 = sols[1][s], c2 = sols[2][s];
    double R[3][3];
    R[0][0] = c0 * c0 - c1 * c1 - c2 * c2 + 1;
    R[0][1] = 2 * c0 * c1 - 2 * c2;
    R[0][2] = 2 * c1 + 2 * c0 * c2;
    R[1][0] = 2 * c2 + 2 * c0 * c1;
    R[1][1] = c1 * c1 - c0 * c0 - c2 * c2 + 1;
    R[1][2] = 2 * c1 * c2 - 2 * c0;
    R[2][0] = 2 * c0 * c2 - 2 * c1;
    R[2][1] = 2 * c0 + 2 * c1 * c2;
    R[2][2] = c2 * c2 - c1 * c1 - c0 * c0 + 1;
    const double nrm = 1 + c0 * c0 + c1 * c1 + c2 * c2;
    double* m = models[s];
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) m[3 * c + r] = R[r][c] / nrm;
    // t = -tt * vec(R)  (:154)
    for (int i = 0; i < 3; ++i) {
      double acc = (-tt[i][0]) * m[0];
      for (int j = 1; j < 9; ++j) acc = acc + (-tt[i][j]) * m[j];
      m[9 + i] = acc;
    }
  }
  return n_sols;
}

// ---------------------------------------------------------------------------------------------
// Scoring (src/estimators/utils.cc:40-89) and support (src/optim/support_measurement.cc:36-60)
// ---------------------------------------------------------------------------------------------
void LineResiduals(const double* lines, const double* points, size_t n, const double* m,
                   double* res) {
  const double P_00 = m[0], P_10 = m[1], P_20 = m[2];
  const double P_01 = m[3], P_11 = m[4], P_21 = m[5];
  const double P_02 = m[6], P_12 = m[7], P_22 = m[8];
  const double P_03 = m[9], P_13 = m[10], P_23 = m[11];
  for (size_t i = 0; i < n; ++i) {
    const double X_0 = points[3 * i + 0];
    const double X_1 = points[3 * i + 1];
    const double X_2 = points[3 * i + 2];
    const double px_2 = P_20 * X_0 + P_21 * X_1 + P_22 * X_2 + P_23;
    if (px_2 > std::numeric_limits<double>::epsilon()) {
      const double px_0 = P_00 * X_0 + P_01 * X_1 + P_02 * X_2 + P_03;
      const double px_1 = P_10 * X_0 + P_11 * X_1 + P_12 * X_2 + P_13;
      const double l_0 = lines[3 * i + 0];
      const double l_1 = lines[3 * i + 1];
      const double l_2 = lines[3 * i + 2];
      const double inv_px_2 = 1.0 / px_2;
      const double r = px_0 * l_0 * inv_px_2 + px_1 * l_1 * inv_px_2 + l_2;
      res[i] = r * r;
    } else {
      res[i] = std::numeric_limits<double>::max();
    }
  }
}

struct Support {
  size_t num_inliers = 0;
  double residual_sum = std::numeric_limits<double>::max();  // support_measurement.h:51-52
};

Support EvaluateSupport(const double* res, size_t n, double max_residual) {
  Support s;
  s.num_inliers = 0;
  s.residual_sum = 0;
  for (size_t i = 0; i < n; ++i) {
    if (res[i] <= max_residual) {
      s.num_inliers += 1;
      s.residual_sum += res[i];
    }
  }
  return s;
}

bool CompareSupport(const Support& a, const Support& b) {  // support_measurement.cc:52-60
  if (a.num_inliers > b.num_inliers) return true;
  return a.num_inliers == b.num_inliers && a.residual_sum < b.residual_sum;
}

// src/optim/ransac.h:158-176
size_t ComputeNumTrials(size_t num_inliers, size_t num_samples, double confidence,
                        double multiplier) {
  const double inlier_ratio = num_inliers / static_cast<double>(num_samples);
  const double nom = 1 - confidence;
  if (nom <= 0) return std::numeric_limits<size_t>::max();
  const double denom = 1 - std::pow(inlier_ratio, 6);
  if (denom <= 0) return 1;
  return static_cast<size_t>(std::ceil(std::log(nom) / std::log(denom) * multiplier));
}

// src/optim/ransac.h:144-156 (constructor) + :178-278 (Estimate)
void RansacP6L(const double* lines, const uint8_t* aligned, const double* points, size_t n,
               const orc_ransac_options& opt_in, orc_ransac_report* report, uint8_t* mask,
               bool adaptive) {
  orc_ransac_options opt = opt_in;
  {
    const size_t kNumSamples = 100000;
    const size_t dyn = ComputeNumTrials(static_cast<size_t>(opt.min_inlier_ratio * kNumSamples),
                                        kNumSamples, opt.confidence,
                                        opt.dyn_num_trials_multiplier);
    opt.max_num_trials = std::min<uint64_t>(opt.max_num_trials, dyn);
  }
  std::memset(report, 0, sizeof(*report));
  report->best_trial = -1;
  report->best_model_idx = -1;
  report->residual_sum = std::numeric_limits<double>::max();
  if (n < 6) return;

  Support best;
  double best_model[12] = {0};
  bool abort = false;
  const double max_residual = opt.max_error * opt.max_error;
  std::vector<double> residuals(n);

  RandomSampler sampler(6);
  sampler.Initialize(n);

  size_t max_num_trials = opt.max_num_trials;
  size_t dyn_max_num_trials = max_num_trials;
  size_t num_trials = 0;
  uint64_t scored = 0;
  for (num_trials = 0; num_trials < max_num_trials; ++num_trials) {
    if (abort) {
      num_trials += 1;
      break;
    }
    size_t idx[6];
    sampler.Sample(idx);
    double l6[6][3], p6[6][3];
    uint8_t a6[6];
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 3; ++j) {
        l6[i][j] = lines[3 * idx[i] + j];
        p6[i][j] = points[3 * idx[i] + j];
      }
      a6[i] = aligned ? aligned[idx[i]] : 0;
    }
    double models[8][12];
    const int nm = P6LEstimate(l6, a6, p6, models);
    for (int m = 0; m < nm; ++m) {
      // P6LEstimator::Residuals (absolute_pose.cc:165-174) first copies every line into a fresh
      // std::vector<Eigen::Vector3d> (emplace_back without reserve) — restated so that the CPU
      // baseline pays the same per-model allocation + copy as the reference.
      {
        struct V3 { double v[3]; };
        std::vector<V3> line_params;
        for (size_t i = 0; i < n; ++i)
          line_params.emplace_back(V3{{lines[3 * i], lines[3 * i + 1], lines[3 * i + 2]}});
        LineResiduals(&line_params[0].v[0], points, n, models[m], residuals.data());
      }
      ++scored;
      const Support s = EvaluateSupport(residuals.data(), n, max_residual);
      if (CompareSupport(s, best)) {
        best = s;
        std::memcpy(best_model, models[m], sizeof(best_model));
        report->best_trial = static_cast<int64_t>(num_trials);
        report->best_model_idx = m;
        if (adaptive)
          dyn_max_num_trials = ComputeNumTrials(best.num_inliers, n, opt.confidence,
                                                opt.dyn_num_trials_multiplier);
      }
      if (adaptive && num_trials >= dyn_max_num_trials && num_trials >= opt.min_num_trials) {
        abort = true;
        break;
      }
    }
  }
  report->num_trials = num_trials;
  report->num_inliers = best.num_inliers;
  report->residual_sum = best.residual_sum;
  report->num_models_scored = scored;
  std::memcpy(report->model, best_model, sizeof(best_model));
  if (best.num_inliers < 6) return;
  report->success = 1;
  if (mask != nullptr) {
    LineResiduals(lines, points, n, best_model, residuals.data());
    for (size_t i = 0; i < n; ++i) mask[i] = residuals[i] <= max_residual ? 1 : 0;
  }
}

}  // namespace

// =============================================================================================
// C interface
// =============================================================================================
extern "C" {

void orc_set_prng_seed(uint32_t seed) {
  delete g_prng;
  g_prng = new std::mt19937(seed);
}

uint32_t orc_prng_peek(void) {
  std::mt19937 copy = Prng();
  return static_cast<uint32_t>(copy());
}

void orc_line_residuals(const double* lines, const double* points, size_t n, const double* model,
                        double* residuals_out) {
  LineResiduals(lines, points, n, model, residuals_out);
}

void orc_inlier_support(const double* residuals, size_t n, double max_residual,
                        uint64_t* num_inliers, double* residual_sum) {
  const Support s = EvaluateSupport(residuals, n, max_residual);
  *num_inliers = s.num_inliers;
  *residual_sum = s.residual_sum;
}

void orc_mestimator_support(const double* residuals, size_t n, double max_residual,
                            uint64_t* num_inliers, double* score) {
  uint64_t cnt = 0;
  double sc = 0;
  for (size_t i = 0; i < n; ++i) {
    if (residuals[i] <= max_residual) {
      cnt += 1;
      sc += residuals[i];
    } else {
      sc += max_residual;
    }
  }
  *num_inliers = cnt;
  *score = sc;
}

uint64_t orc_compute_num_trials(uint64_t num_inliers, uint64_t num_samples, double confidence,
                                double num_trials_multiplier) {
  return ComputeNumTrials(num_inliers, num_samples, confidence, num_trials_multiplier);
}

void orc_sample_table(size_t n, size_t num_trials, uint32_t* table_out) {
  RandomSampler sampler(6);
  sampler.Initialize(n);
  for (size_t t = 0; t < num_trials; ++t) {
    size_t idx[6];
    sampler.Sample(idx);
    for (int i = 0; i < 6; ++i) table_out[6 * t + i] = static_cast<uint32_t>(idx[i]);
  }
}

int orc_re3q3(const double* coeffs, double* solutions) {
  double c[3][10], s[3][8];
  std::memcpy(c, coeffs, sizeof(c));
  std::memset(s, 0, sizeof(s));
  const int n = Re3q3Impl(c, s, true);
  for (int k = 0; k < 8; ++k)
    for (int i = 0; i < 3; ++i) solutions[3 * k + i] = s[i][k];
  return n;
}

int orc_poly8_all_roots(const double* c, double* re_im_out) {
  double re[8], im[8];
  const bool ok = Poly8Roots(c, re, im);
  for (int i = 0; i < 8; ++i) {
    re_im_out[2 * i] = re[i];
    re_im_out[2 * i + 1] = im[i];
  }
  return ok ? 8 : -1;
}

/* diagnostics: Francis sweeps the eigenvalue step of the last P6L / re3q3 / poly8 call took */
int orc_last_qr_sweeps(void) { return g_last_qr_sweeps; }

void orc_re3q3_resultant(const double* P, double* a_out, double* c_out) {
  double Pm[3][7], a[33], c[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 7; ++j) Pm[i][j] = P[7 * i + j];
  re3q3_resultant(Pm, a, c);
  for (int i = 0; i < 33; ++i) a_out[i] = a[i];
  for (int i = 0; i < 9; ++i) c_out[i] = c[i];
}

void orc_re3q3_backsubstitute(const double* a_in, double x, double* yz_out) {
  double a[33];
  for (int i = 0; i < 33; ++i) a[i] = a_in[i];
  re3q3_backsubstitute(a, x, &yz_out[0], &yz_out[1]);
}

int orc_poly8_real_roots(const double* c, double* roots_out) {
  double re[8], im[8];
  Poly8Roots(c, re, im);
  int n = 0;
  for (int i = 0; i < 8; ++i)
    if (!(std::fabs(im[i]) > 1e-8)) roots_out[n++] = re[i];
  return n;
}

int orc_p6l_estimate(const double* lines6, const uint8_t* aligned6, const double* points6,
                     double* models_out) {
  double l[6][3], p[6][3], m[8][12];
  std::memcpy(l, lines6, sizeof(l));
  std::memcpy(p, points6, sizeof(p));
  const int n = P6LEstimate(l, aligned6, p, m);
  std::memcpy(models_out, m, sizeof(double) * 12 * static_cast<size_t>(n));
  return n;
}

void orc_rotation_matrix_to_quaternion(const double* R, double* q) {
  QuaternionFromRotationMatrix(R, q);  // Eigen::Quaterniond(Matrix3d): eigen_restated.h
}

void orc_ransac_p6l(const double* lines, const uint8_t* aligned, const double* points, size_t n,
                    const orc_ransac_options* options, orc_ransac_report* report,
                    uint8_t* inlier_mask) {
  RansacP6L(lines, aligned, points, n, *options, report, inlier_mask, true);
}

int orc_estimate_absolute_pose_from_lines(const double* lines, const uint8_t* aligned,
                                          const double* points, size_t n,
                                          const orc_ransac_options* options, double* qvec,
                                          double* tvec, uint64_t* num_inliers,
                                          uint8_t* inlier_mask, orc_ransac_report* report_out) {
  // src/estimators/pose.cc:52-94
  orc_ransac_report report;
  std::vector<uint8_t> mask(n, 0);
  RansacP6L(lines, aligned, points, n, *options, &report, mask.data(), true);
  if (report_out) *report_out = report;
  *num_inliers = report.num_inliers;
  // report.inlier_mask is empty unless success; callers index it only when num_inliers > 0.
  if (inlier_mask) std::memcpy(inlier_mask, mask.data(), n);
  if (*num_inliers == 0) return 0;
  size_t num_aligned_inliers = 0;
  for (size_t i = 0; i < n; ++i)
    if (mask[i] && aligned && aligned[i]) num_aligned_inliers += 1;
  if (num_aligned_inliers > *num_inliers * 0.9) return 0;
  orc_rotation_matrix_to_quaternion(report.model, qvec);
  tvec[0] = report.model[9];
  tvec[1] = report.model[10];
  tvec[2] = report.model[11];
  for (int i = 0; i < 4; ++i)
    if (std::isnan(qvec[i])) return 0;
  for (int i = 0; i < 3; ++i)
    if (std::isnan(tvec[i])) return 0;
  return 1;
}

uint64_t orc_ransac_p6l_fixed_trials(const double* lines, const uint8_t* aligned,
                                     const double* points, size_t n, double max_error,
                                     uint64_t num_trials, orc_ransac_report* report) {
  orc_ransac_options opt;
  opt.max_error = max_error;
  opt.min_inlier_ratio = 0.0;  // constructor cap becomes SIZE_MAX-like (denom = 1 -> log(1)=0)
  opt.confidence = 0.99999;
  opt.dyn_num_trials_multiplier = 3.0;
  opt.min_num_trials = num_trials;
  opt.max_num_trials = num_trials;
  orc_ransac_report local;
  // min_inlier_ratio = 0 makes ComputeNumTrials divide by log(1) = 0 -> +inf -> cast UB; avoid it
  // by using a tiny ratio whose cap far exceeds any realistic num_trials.
  opt.min_inlier_ratio = 0.01;
  RansacP6L(lines, aligned, points, n, opt, report ? report : &local, nullptr, false);
  return (report ? report : &local)->num_models_scored;
}

}  // extern "C"
