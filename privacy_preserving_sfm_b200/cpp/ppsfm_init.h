// ppsfm_init.h — four-view gravity-aided initialisation from lifted lines, host
// side.  Dependency-free C++17 (no Eigen / Ceres); the estimators implement the RansacLib Solver
// concept, so they run under ppsfm::LocallyOptimizedMSAC (ppsfm_lomsac.h) and, unchanged, under the
// reference's own ransac_lib::LocallyOptimizedMSAC (oracle/ref_init.cc).
//
// Mirrors, with the reference's names and argument meaning:
//   init::FourView2dEstimator        src/init/sfm2d.h:48-100, sfm2d.cc:302-489
//   init::AbsolutePose2dEstimator    src/init/sfm2d.h:102-148, sfm2d.cc:491-530
//   init::PlanarOffsetEstimator      src/init/initializer.h:62-101, initializer.cc:219-333, 450-467
//   init::initialize_reconstruction  src/init/initializer.h:103-108, initializer.cc:57-215
//   init::InitOptions                src/init/initializer.h:49-58
// Deviations (documented in DESIGN.md): Eigen's JacobiSVD / colPivHouseholderQr / partialPivLu are
// replaced by the one-sided Jacobi SVD, pivoted Householder QR and pivoted LU below; the random
// projective change of variables in factorize_trifocal_tensor (Matrix2d::setRandom, i.e. C rand(),
// sfm2d.cc:234-237) uses FIXED generic matrices so that results are reproducible; the Ceres
// problems (2-D bundle adjustment, point refinement; sfm2d.cc:76-175) are solved by the small
// Levenberg-Marquardt below with the same cost, gauge and tolerances.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

#include "ppsfm_init_math.h"
#include "ppsfm_lomsac.h"

namespace ppsfm {
namespace init {

using Vec2 = std::array<double, 2>;
using Vec3 = std::array<double, 3>;
struct Pose2d {  // 2x3, row-major: [R(2x2) | t]
  double m[2][3] = {{0, 0, 0}, {0, 0, 0}};
};
struct Pose {    // 3x4, row-major: [R | t]
  double m[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
};
struct Mat3 {
  double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
};

struct InitOptions {  // initializer.h:49-58
  double min_tri_angle = 0.1;
  double min_num_inliers = 6;
  double max_error = 0.005;
};

// ------------------------------------------------------------------------------------------
// small dense linear algebra
// ------------------------------------------------------------------------------------------
namespace la {

inline double norm2(const Vec2& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1]); }
inline double norm3(const Vec3& v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
inline Vec2 apply(const Pose2d& P, const Vec2& X) {
  return {P.m[0][0] * X[0] + P.m[0][1] * X[1] + P.m[0][2],
          P.m[1][0] * X[0] + P.m[1][1] * X[1] + P.m[1][2]};
}
inline Vec3 apply(const Pose& P, const Vec3& X) {
  Vec3 z;
  for (int r = 0; r < 3; ++r) z[r] = P.m[r][0] * X[0] + P.m[r][1] * X[1] + P.m[r][2] * X[2] + P.m[r][3];
  return z;
}
inline Vec3 mul(const Mat3& R, const Vec3& v) {
  Vec3 z;
  for (int r = 0; r < 3; ++r) z[r] = R.m[r][0] * v[0] + R.m[r][1] * v[1] + R.m[r][2] * v[2];
  return z;
}

// Right singular vectors of the m x n matrix A (row-major), n <= 8, by one-sided Jacobi
// (Hestenes): columns of V sorted by descending singular value, like Eigen::JacobiSVD.
inline void svd_right_vectors(const std::vector<double>& A, int m, int n, double* V /* n*n */,
                              double* sv /* n */) {
  std::vector<double> W(A);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < m; ++r) {
          const double wp = W[r * n + p], wq = W[r * n + q];
          alpha += wp * wp;
          beta += wq * wq;
          gamma += wp * wq;
        }
        if (std::fabs(gamma) <= 1e-300 || std::fabs(gamma) <= 2.3e-16 * std::sqrt(alpha * beta))
          continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < m; ++r) {
          const double wp = W[r * n + p], wq = W[r * n + q];
          W[r * n + p] = c * wp - s * wq;
          W[r * n + q] = s * wp + c * wq;
        }
        for (int r = 0; r < n; ++r) {
          const double vp = V[r * n + p], vq = V[r * n + q];
          V[r * n + p] = c * vp - s * vq;
          V[r * n + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  int order[8];
  double s2[8];
  for (int j = 0; j < n; ++j) {
    order[j] = j;
    s2[j] = 0;
    for (int r = 0; r < m; ++r) s2[j] += W[r * n + j] * W[r * n + j];
  }
  std::stable_sort(order, order + n, [&](int a, int b) { return s2[a] > s2[b]; });
  double Vs[64];
  for (int j = 0; j < n; ++j) {
    sv[j] = std::sqrt(s2[order[j]]);
    for (int r = 0; r < n; ++r) Vs[r * n + j] = V[r * n + order[j]];
  }
  std::copy(Vs, Vs + n * n, V);
}

// min |A x - b| for the m x n matrix A (row-major, n <= 4, m >= 1) by Householder QR with column
// pivoting; directions whose pivot is negligible get a zero coefficient.
inline void qr_solve(std::vector<double> A, int m, int n, std::vector<double> b, double* x) {
  int perm[4] = {0, 1, 2, 3};
  int rank = 0;
  double first_pivot = 0.0;
  const int steps = std::min(m, n);
  for (int k = 0; k < steps; ++k) {
    int best = k;
    double best_norm = -1.0;
    for (int j = k; j < n; ++j) {
      double s = 0;
      for (int r = k; r < m; ++r) s += A[r * n + j] * A[r * n + j];
      if (s > best_norm) {
        best_norm = s;
        best = j;
      }
    }
    if (best != k) {
      for (int r = 0; r < m; ++r) std::swap(A[r * n + k], A[r * n + best]);
      std::swap(perm[k], perm[best]);
    }
    const double nrm = std::sqrt(best_norm);
    if (k == 0) first_pivot = nrm;
    if (nrm <= 1e-14 * first_pivot || nrm == 0.0) break;
    ++rank;
    const double akk = A[k * n + k];
    const double alpha = akk >= 0 ? -nrm : nrm;
    std::vector<double> v(m - k);
    v[0] = akk - alpha;
    for (int r = k + 1; r < m; ++r) v[r - k] = A[r * n + k];
    double vtv = 0;
    for (double e : v) vtv += e * e;
    if (vtv > 0) {
      for (int j = k; j < n; ++j) {
        double d = 0;
        for (int r = k; r < m; ++r) d += v[r - k] * A[r * n + j];
        d = 2.0 * d / vtv;
        for (int r = k; r < m; ++r) A[r * n + j] -= d * v[r - k];
      }
      double d = 0;
      for (int r = k; r < m; ++r) d += v[r - k] * b[r];
      d = 2.0 * d / vtv;
      for (int r = k; r < m; ++r) b[r] -= d * v[r - k];
    }
  }
  double y[4] = {0, 0, 0, 0};
  for (int k = rank - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < rank; ++j) s -= A[k * n + j] * y[j];
    y[k] = s / A[k * n + k];
  }
  for (int j = 0; j < n; ++j) x[j] = 0.0;
  for (int k = 0; k < rank; ++k) x[perm[k]] = y[k];
}

// X = A^-1 B for a 3x3 A and a 3 x nb B (row-major) by LU with partial pivoting
inline void lu3_solve(Mat3 A, double* B, int nb) {
  for (int k = 0; k < 3; ++k) {
    int piv = k;
    for (int r = k + 1; r < 3; ++r)
      if (std::fabs(A.m[r][k]) > std::fabs(A.m[piv][k])) piv = r;
    if (piv != k) {
      for (int c = 0; c < 3; ++c) std::swap(A.m[k][c], A.m[piv][c]);
      for (int c = 0; c < nb; ++c) std::swap(B[k * nb + c], B[piv * nb + c]);
    }
    for (int r = k + 1; r < 3; ++r) {
      const double f = A.m[r][k] / A.m[k][k];
      for (int c = k; c < 3; ++c) A.m[r][c] -= f * A.m[k][c];
      for (int c = 0; c < nb; ++c) B[r * nb + c] -= f * B[k * nb + c];
    }
  }
  for (int c = 0; c < nb; ++c)
    for (int k = 2; k >= 0; --k) {
      double s = B[k * nb + c];
      for (int j = k + 1; j < 3; ++j) s -= A.m[k][j] * B[j * nb + c];
      B[k * nb + c] = s / A.m[k][k];
    }
}

// rotation taking direction a onto direction b (Eigen::Quaterniond::FromTwoVectors)
inline Mat3 rotation_from_two_vectors(const Vec3& a, const Vec3& b) {
  const double na = norm3(a), nb = norm3(b);
  const Vec3 u{a[0] / na, a[1] / na, a[2] / na}, v{b[0] / nb, b[1] / nb, b[2] / nb};
  const double c = u[0] * v[0] + u[1] * v[1] + u[2] * v[2];
  double w, x, y, z;
  if (c < -1.0 + 1e-12) {  // opposite: half turn about any axis orthogonal to u
    Vec3 ax = std::fabs(u[0]) < 0.9 ? Vec3{0, -u[2], u[1]} : Vec3{-u[2], 0, u[0]};
    const double n = norm3(ax);
    w = 0;
    x = ax[0] / n; y = ax[1] / n; z = ax[2] / n;
  } else {
    const double s = std::sqrt((1.0 + c) * 2.0);
    x = (u[1] * v[2] - u[2] * v[1]) / s;
    y = (u[2] * v[0] - u[0] * v[2]) / s;
    z = (u[0] * v[1] - u[1] * v[0]) / s;
    w = 0.5 * s;
  }
  Mat3 R;
  R.m[0][0] = 1 - 2 * (y * y + z * z); R.m[0][1] = 2 * (x * y - w * z); R.m[0][2] = 2 * (x * z + w * y);
  R.m[1][0] = 2 * (x * y + w * z); R.m[1][1] = 1 - 2 * (x * x + z * z); R.m[1][2] = 2 * (y * z - w * x);
  R.m[2][0] = 2 * (x * z - w * y); R.m[2][1] = 2 * (y * z + w * x); R.m[2][2] = 1 - 2 * (x * x + y * y);
  return R;
}

}  // namespace la

// ------------------------------------------------------------------------------------------
// 2-D structure from motion on the gravity plane
// ------------------------------------------------------------------------------------------
namespace detail {

// residual of BundleAdjustment2DCostFunction (sfm2d.cc:42-74) and its derivatives
struct Res2d {
  double r, dq[2], dt[2], dX[2];
};
inline Res2d residual2d(const double q[2], const double t[2], const Vec2& X, const Vec2& x) {
  const double p0 = q[0] * X[0] - q[1] * X[1] + t[0], p1 = q[1] * X[0] + q[0] * X[1] + t[1];
  Res2d o;
  o.r = p0 / p1 - x[0] / x[1];
  const double g0 = 1.0 / p1, g1 = -p0 / (p1 * p1);
  o.dq[0] = g0 * X[0] + g1 * X[1];
  o.dq[1] = -g0 * X[1] + g1 * X[0];
  o.dt[0] = g0;
  o.dt[1] = g1;
  o.dX[0] = g0 * q[0] + g1 * q[1];
  o.dX[1] = -g0 * q[1] + g1 * q[0];
  return o;
}

// Levenberg-Marquardt on sum r^2 with `nc` dense "camera" unknowns and one 2-vector per point,
// points eliminated by the Schur complement (what ceres DENSE_SCHUR does for these problems).
// eval(xc, X, &r, &Jc(row-major nres x nc), &Jp(nres x 2), &pt(nres)) linearises at the state;
// plus(xc, dc) applies a camera step.  Tolerances / radius rule follow ceres' defaults with
// function = gradient = parameter tolerance 1e-10 (sfm2d.cc:105-108, 152-156).
template <class Eval, class Plus>
inline void lm_schur2(int nc, std::vector<double>* cam_state, std::vector<Vec2>* X, Eval eval,
                      Plus plus) {
  const int np = static_cast<int>(X->size());
  std::vector<double> r, Jc, Jp;
  std::vector<int> pt;
  auto cost_of = [&](const std::vector<double>& rr) {
    double c = 0;
    for (double e : rr) c += e * e;
    return 0.5 * c;
  };
  eval(*cam_state, *X, &r, &Jc, &Jp, &pt);
  double cost = cost_of(r);
  double radius = 1e4, decrease = 2.0;
  for (int iter = 0; iter < 50; ++iter) {
    const int nres = static_cast<int>(r.size());
    std::vector<double> U(nc * nc, 0.0), gc(nc, 0.0), V(3 * np, 0.0), gp(2 * np, 0.0), W(nc * 2 * np, 0.0);
    for (int k = 0; k < nres; ++k) {
      const double* jc = nc ? &Jc[(size_t)k * nc] : nullptr;
      const double* jp = &Jp[(size_t)k * 2];
      const int p = pt[k];
      for (int a = 0; a < nc; ++a) {
        gc[a] += jc[a] * r[k];
        for (int b = 0; b < nc; ++b) U[a * nc + b] += jc[a] * jc[b];
        W[(size_t)a * 2 * np + 2 * p] += jc[a] * jp[0];
        W[(size_t)a * 2 * np + 2 * p + 1] += jc[a] * jp[1];
      }
      V[3 * p] += jp[0] * jp[0];
      V[3 * p + 1] += jp[0] * jp[1];
      V[3 * p + 2] += jp[1] * jp[1];
      gp[2 * p] += jp[0] * r[k];
      gp[2 * p + 1] += jp[1] * r[k];
    }
    double gmax = 0;
    for (double g : gc) gmax = std::max(gmax, std::fabs(g));
    for (double g : gp) gmax = std::max(gmax, std::fabs(g));
    if (gmax <= 1e-10) break;
    auto damp = [&](double d) { return std::min(std::max(d, 1e-6), 1e32) / radius; };
    // damped point blocks and their inverses
    std::vector<double> Vi(3 * np);
    for (int p = 0; p < np; ++p) {
      const double a = V[3 * p] + damp(V[3 * p]), b = V[3 * p + 1], c = V[3 * p + 2] + damp(V[3 * p + 2]);
      const double det = a * c - b * b;
      Vi[3 * p] = c / det;
      Vi[3 * p + 1] = -b / det;
      Vi[3 * p + 2] = a / det;
    }
    std::vector<double> S(U), rhs(nc);
    for (int a = 0; a < nc; ++a) {
      S[a * nc + a] += damp(U[a * nc + a]);
      rhs[a] = -gc[a];
    }
    for (int p = 0; p < np; ++p) {
      const double v0 = Vi[3 * p], v1 = Vi[3 * p + 1], v2 = Vi[3 * p + 2];
      for (int a = 0; a < nc; ++a) {
        const double w0 = W[(size_t)a * 2 * np + 2 * p], w1 = W[(size_t)a * 2 * np + 2 * p + 1];
        const double y0 = w0 * v0 + w1 * v1, y1 = w0 * v1 + w1 * v2;
        rhs[a] += y0 * gp[2 * p] + y1 * gp[2 * p + 1];
        for (int b = 0; b < nc; ++b)
          S[a * nc + b] -= y0 * W[(size_t)b * 2 * np + 2 * p] + y1 * W[(size_t)b * 2 * np + 2 * p + 1];
      }
    }
    // dense Cholesky solve of the reduced system
    std::vector<double> dc(rhs);
    bool ok = true;
    {
      std::vector<double> L(S);
      for (int j = 0; j < nc && ok; ++j) {
        double d = L[j * nc + j];
        for (int k = 0; k < j; ++k) d -= L[j * nc + k] * L[j * nc + k];
        if (!(d > 0)) { ok = false; break; }
        L[j * nc + j] = std::sqrt(d);
        for (int i = j + 1; i < nc; ++i) {
          double s = L[i * nc + j];
          for (int k = 0; k < j; ++k) s -= L[i * nc + k] * L[j * nc + k];
          L[i * nc + j] = s / L[j * nc + j];
        }
      }
      if (ok) {
        for (int i = 0; i < nc; ++i) {
          double s = dc[i];
          for (int k = 0; k < i; ++k) s -= L[i * nc + k] * dc[k];
          dc[i] = s / L[i * nc + i];
        }
        for (int i = nc - 1; i >= 0; --i) {
          double s = dc[i];
          for (int k = i + 1; k < nc; ++k) s -= L[k * nc + i] * dc[k];
          dc[i] = s / L[i * nc + i];
        }
      }
    }
    std::vector<double> dp(2 * np, 0.0);
    double model_change = 0, step2 = 0, x2 = 0;
    std::vector<double> cand_cam(*cam_state);
    std::vector<Vec2> cand_X(*X);
    if (ok) {
      for (int p = 0; p < np; ++p) {
        double a0 = gp[2 * p], a1 = gp[2 * p + 1];
        for (int a = 0; a < nc; ++a) {
          a0 += W[(size_t)a * 2 * np + 2 * p] * dc[a];
          a1 += W[(size_t)a * 2 * np + 2 * p + 1] * dc[a];
        }
        dp[2 * p] = -(Vi[3 * p] * a0 + Vi[3 * p + 1] * a1);
        dp[2 * p + 1] = -(Vi[3 * p + 1] * a0 + Vi[3 * p + 2] * a1);
      }
      for (int k = 0; k < nres; ++k) {
        double m = Jp[2 * (size_t)k] * dp[2 * pt[k]] + Jp[2 * (size_t)k + 1] * dp[2 * pt[k] + 1];
        for (int a = 0; a < nc; ++a) m += Jc[(size_t)k * nc + a] * dc[a];
        model_change -= m * (r[k] + 0.5 * m);
      }
      plus(&cand_cam, dc);
      for (int p = 0; p < np; ++p) {
        cand_X[p][0] += dp[2 * p];
        cand_X[p][1] += dp[2 * p + 1];
        step2 += dp[2 * p] * dp[2 * p] + dp[2 * p + 1] * dp[2 * p + 1];
        x2 += (*X)[p][0] * (*X)[p][0] + (*X)[p][1] * (*X)[p][1];
      }
      for (size_t a = 0; a < cand_cam.size(); ++a) {
        step2 += (cand_cam[a] - (*cam_state)[a]) * (cand_cam[a] - (*cam_state)[a]);
        x2 += (*cam_state)[a] * (*cam_state)[a];
      }
    }
    if (!ok || !(model_change > 0)) {
      radius /= decrease;
      decrease *= 2;
      if (radius < 1e-32) break;
      continue;
    }
    if (std::sqrt(step2) <= 1e-10 * (std::sqrt(x2) + 1e-10)) break;
    std::vector<double> r2, Jc2, Jp2;
    std::vector<int> pt2;
    eval(cand_cam, cand_X, &r2, &Jc2, &Jp2, &pt2);
    const double cost2 = cost_of(r2);
    const double change = cost - cost2;
    const double rho = change / model_change;
    if (rho > 1e-3 && std::isfinite(cost2)) {
      *cam_state = cand_cam;
      *X = cand_X;
      r.swap(r2); Jc.swap(Jc2); Jp.swap(Jp2); pt.swap(pt2);
      const double tmp = 2 * rho - 1;
      radius = std::min(1e16, radius / std::max(1.0 / 3.0, 1 - tmp * tmp * tmp));
      decrease = 2;
      const bool converged = std::fabs(change) <= 1e-10 * cost;
      cost = cost2;
      if (converged) break;
    } else {
      radius /= decrease;
      decrease *= 2;
      if (radius < 1e-32) break;
    }
  }
}

// refine all points against fixed cameras (optimize_points2d, sfm2d.cc:76-116)
inline void optimize_points2d(const std::vector<Pose2d>& cams, const std::vector<std::vector<Vec2>>& x,
                              std::vector<Vec2>* X) {
  if (x.empty()) return;
  std::vector<double> none;
  auto eval = [&](const std::vector<double>&, const std::vector<Vec2>& Xs, std::vector<double>* r,
                  std::vector<double>* Jc, std::vector<double>* Jp, std::vector<int>* pt) {
    r->clear(); Jc->clear(); Jp->clear(); pt->clear();
    for (size_t i = 0; i < cams.size(); ++i) {
      const double q[2] = {cams[i].m[0][0], cams[i].m[1][0]}, t[2] = {cams[i].m[0][2], cams[i].m[1][2]};
      for (size_t j = 0; j < x[i].size(); ++j) {
        const Res2d o = residual2d(q, t, Xs[j], x[i][j]);
        r->push_back(o.r);
        Jp->push_back(o.dX[0]);
        Jp->push_back(o.dX[1]);
        pt->push_back((int)j);
      }
    }
  };
  lm_schur2(0, &none, X, eval, [](std::vector<double>*, const std::vector<double>&) {});
}

// 2-D bundle adjustment (bundle_adjust2d, sfm2d.cc:118-175): camera 0 fixed, rotations move on
// their circle (HomogeneousVectorParameterization(2)), so does the translation of camera 1 (scale
// gauge); state per camera = (q0, q1, t0, t1).
inline void bundle_adjust2d(std::vector<Pose2d>* cams, const std::vector<std::vector<Vec2>>& x,
                            std::vector<Vec2>* X) {
  if (x.empty() || x[0].size() < 10) return;
  const int ncam = static_cast<int>(cams->size());
  std::vector<double> state(4 * ncam);
  for (int i = 0; i < ncam; ++i) {
    state[4 * i] = (*cams)[i].m[0][0];
    state[4 * i + 1] = (*cams)[i].m[1][0];
    state[4 * i + 2] = (*cams)[i].m[0][2];
    state[4 * i + 3] = (*cams)[i].m[1][2];
  }
  // tangent layout: camera i >= 1: [theta_q] then camera 1: [theta_t], cameras >= 2: [t0, t1]
  std::vector<int> off(ncam, -1);
  int nc = 0;
  for (int i = 1; i < ncam; ++i) {
    off[i] = nc;
    nc += (i == 1) ? 2 : 3;
  }
  auto eval = [&](const std::vector<double>& s, const std::vector<Vec2>& Xs, std::vector<double>* r,
                  std::vector<double>* Jc, std::vector<double>* Jp, std::vector<int>* pt) {
    r->clear(); Jc->clear(); Jp->clear(); pt->clear();
    for (int i = 0; i < ncam; ++i) {
      const double* q = &s[4 * i];
      const double* t = &s[4 * i + 2];
      for (size_t j = 0; j < x[i].size(); ++j) {
        const Res2d o = residual2d(q, t, Xs[j], x[i][j]);
        r->push_back(o.r);
        const size_t base = Jc->size();
        Jc->resize(base + nc, 0.0);
        if (i >= 1) {
          double* row = &(*Jc)[base + off[i]];
          row[0] = o.dq[0] * (-q[1]) + o.dq[1] * q[0];
          if (i == 1) {
            row[1] = o.dt[0] * (-t[1]) + o.dt[1] * t[0];
          } else {
            row[1] = o.dt[0];
            row[2] = o.dt[1];
          }
        }
        Jp->push_back(o.dX[0]);
        Jp->push_back(o.dX[1]);
        pt->push_back((int)j);
      }
    }
  };
  auto rotate = [](double* v, double ang) {
    const double c = std::cos(ang), s = std::sin(ang), a = v[0], b = v[1];
    v[0] = c * a - s * b;
    v[1] = s * a + c * b;
  };
  auto plus = [&](std::vector<double>* s, const std::vector<double>& d) {
    for (int i = 1; i < ncam; ++i) {
      rotate(&(*s)[4 * i], d[off[i]]);
      if (i == 1) {
        rotate(&(*s)[4 * i + 2], d[off[i] + 1]);
      } else {
        (*s)[4 * i + 2] += d[off[i] + 1];
        (*s)[4 * i + 3] += d[off[i] + 2];
      }
    }
  };
  lm_schur2(nc, &state, X, eval, plus);
  for (int i = 0; i < ncam; ++i) {
    Pose2d& c = (*cams)[i];
    c.m[0][0] = state[4 * i]; c.m[0][1] = -state[4 * i + 1];
    c.m[1][0] = state[4 * i + 1]; c.m[1][1] = state[4 * i];
    c.m[0][2] = state[4 * i + 2]; c.m[1][2] = state[4 * i + 3];
  }
}

// sfm2d.cc:178-193
inline void metric_upgrade(const Pose2d& P2, const Pose2d& P3, double H[3][3]) {
  const std::vector<double> A = {P2.m[0][2], -P2.m[1][2], P2.m[1][2], P2.m[0][2],
                                 P3.m[0][2], -P3.m[1][2], P3.m[1][2], P3.m[0][2]};
  const std::vector<double> b = {P2.m[1][1] - P2.m[0][0], -P2.m[0][1] - P2.m[1][0],
                                 P3.m[1][1] - P3.m[0][0], -P3.m[0][1] - P3.m[1][0]};
  double x[2];
  la::qr_solve(A, 4, 2, b, x);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) H[r][c] = (r == c) ? 1.0 : 0.0;
  H[2][0] = x[0];
  H[2][1] = x[1];
}

// sfm2d.cc:196-215
inline void three_view_triangulate2d(const Pose2d& P1, const Pose2d& P2, const Pose2d& P3,
                                     const std::vector<Vec2>& x1, const std::vector<Vec2>& x2,
                                     const std::vector<Vec2>& x3, std::vector<Vec2>* X) {
  // (per point: hd::triangulate2d_point, the arithmetic the GPU scoring kernel shares)
  for (size_t i = 0; i < x1.size(); ++i) {
    Vec2 sol;
    hd::triangulate2d_point(&P1.m[0][0], &P2.m[0][0], &P3.m[0][0], x1[i].data(), x2[i].data(),
                            x3[i].data(), sol.data());
    X->push_back(sol);
  }
}

// sfm2d.cc:217-226 (A = 2x2 column-major: A(0) A(2) / A(1) A(3))
inline void trifocal_coord_change(const double T[8], const double A1[4], const double A2[4],
                                  const double A3[4], double out[8]) {
  for (int k = 0; k < 8; ++k) {
    const int i1 = (k & 1) ? 2 : 0, i2 = (k & 2) ? 2 : 0, i3 = (k & 4) ? 2 : 0;
    auto inner = [&](int base) {
      return A2[i2] * (A1[i1] * T[base] + A1[i1 + 1] * T[base + 1]) +
             A2[i2 + 1] * (A1[i1] * T[base + 2] + A1[i1 + 1] * T[base + 3]);
    };
    out[k] = A3[i3] * inner(0) + A3[i3 + 1] * inner(4);
  }
}

inline void mul22_left(const double A[4] /* col-major */, Pose2d* P) {  // P <- A P
  for (int c = 0; c < 3; ++c) {
    const double p0 = P->m[0][c], p1 = P->m[1][c];
    P->m[0][c] = A[0] * p0 + A[2] * p1;
    P->m[1][c] = A[1] * p0 + A[3] * p1;
  }
}
inline void mul22_right_inv(const double A[4] /* col-major */, Pose2d* P) {  // P(:,0:2) <- P(:,0:2) A^-1
  const double det = A[0] * A[3] - A[2] * A[1];
  const double i00 = A[3] / det, i01 = -A[2] / det, i10 = -A[1] / det, i11 = A[0] / det;
  for (int r = 0; r < 2; ++r) {
    const double p0 = P->m[r][0], p1 = P->m[r][1];
    P->m[r][0] = p0 * i00 + p1 * i10;
    P->m[r][1] = p0 * i01 + p1 * i11;
  }
}

// sfm2d.cc:229-300.  The reference draws A1..A3 with Matrix2d::setRandom(); fixed generic
// matrices here (see file header).
inline int factorize_trifocal_tensor(const double T[8], Pose2d P1[2], Pose2d P2[2], Pose2d P3[2]) {
  static const double A1[4] = {0.6803754343094190, -0.2112341463618254, 0.5661984475172117, 0.5968800669521466};
  static const double A2[4] = {0.8232947158735686, -0.6048972614132321, -0.3295544885702220, 0.5364591896238079};
  static const double A3[4] = {-0.4444505783936810, 0.1079399049282288, -0.0452058962756795, 0.2577418495238488};
  double AT[8];
  trifocal_coord_change(T, A1, A2, A3, AT);
  const double alpha = AT[2] * AT[7] - AT[3] * AT[6];
  const double beta = AT[1] * AT[6] + AT[3] * AT[4] - AT[0] * AT[7] - AT[2] * AT[5];
  const double gamma = AT[0] * AT[5] - AT[1] * AT[4];
  const double disc = beta * beta - 4.0 * alpha * gamma;
  if (disc < 0) return 0;
  const double sq = std::sqrt(disc);
  double aa1[2];
  aa1[0] = (beta > 0) ? (2 * gamma) / (-beta - sq) : (2.0 * gamma) / (-beta + sq);
  aa1[1] = gamma / (alpha * aa1[0]);
  int n_sols = 0;
  for (int i = 0; i < 2; ++i) {
    double a1 = aa1[i];
    const double s = std::sqrt(1 + a1 * a1);
    a1 /= s;
    const double a2 = 1 / s;
    const double rho = -(AT[1] * a2 - AT[3] * a1) / (AT[2] * a1 - AT[0] * a2);
    const double b1 = rho * a1, b2 = rho * a2, c1 = -a2, c2 = a1;
    const std::vector<double> G = {
        0, AT[7] * c2, -AT[0] * c1, 0, AT[0] * b1, -AT[7] * a2,
        0, 0, -AT[1] * c1, AT[7] * c2, AT[1] * b1, -AT[7] * b2,
        0, -AT[7] * c1, -AT[2] * c1, 0, AT[2] * b1, AT[7] * a1,
        0, 0, -AT[3] * c1, -AT[7] * c1, AT[3] * b1, AT[7] * b1,
        -AT[7] * c2, 0, -AT[4] * c1, 0, AT[7] * a2 + AT[4] * b1, 0,
        0, 0, -AT[5] * c1 - AT[7] * c2, 0, AT[7] * b2 + AT[5] * b1, 0,
        AT[7] * c1, 0, -AT[6] * c1, 0, -AT[7] * a1 + AT[6] * b1, 0};
    double V[36], sv[6];
    la::svd_right_vectors(G, 7, 6, V, sv);
    double def[6];
    for (int r = 0; r < 6; ++r) def[r] = V[r * 6 + 5];
    P1[n_sols] = Pose2d();
    P1[n_sols].m[0][0] = 1;
    P1[n_sols].m[1][1] = 1;
    P2[n_sols].m[0][0] = a1; P2[n_sols].m[0][1] = b1; P2[n_sols].m[0][2] = c1;
    P2[n_sols].m[1][0] = a2; P2[n_sols].m[1][1] = b2; P2[n_sols].m[1][2] = c2;
    P3[n_sols].m[0][0] = def[0]; P3[n_sols].m[0][1] = def[2]; P3[n_sols].m[0][2] = def[4];
    P3[n_sols].m[1][0] = def[1]; P3[n_sols].m[1][1] = def[3]; P3[n_sols].m[1][2] = def[5];
    ++n_sols;
  }
  for (int i = 0; i < n_sols; ++i) {  // revert the change of coordinates, first camera = [I 0]
    mul22_left(A2, &P2[i]);
    mul22_left(A3, &P3[i]);
    mul22_right_inv(A1, &P2[i]);
    mul22_right_inv(A1, &P3[i]);
  }
  return n_sols;
}

}  // namespace detail

class FourView2dEstimator {
 public:
  struct Reconstruction {
    std::vector<Pose2d> cams;
    std::vector<Vec2> X;
  };
  typedef std::vector<Reconstruction> ReconstructionVector;

  FourView2dEstimator(const std::vector<Vec2>& x1, const std::vector<Vec2>& x2,
                      const std::vector<Vec2>& x3, const std::vector<Vec2>& x4,
                      double inlier_threshold)
      : x1_(x1), x2_(x2), x3_(x3), x4_(x4), inlier_threshold_(inlier_threshold) {
    for (auto* xs : {&x1_, &x2_, &x3_, &x4_})
      for (Vec2& v : *xs) {
        const double n = la::norm2(v);
        v[0] /= n;
        v[1] /= n;
      }
  }
  int min_sample_size() const { return 5; }
  int non_minimal_sample_size() const { return 2 * min_sample_size(); }
  int num_data() const { return static_cast<int>(x1_.size()); }

  // sfm2d.cc:302-319: max over the four views of the 1-D reprojection error
  double EvaluateModelOnPoint(const Reconstruction& model, int i) const {
    static_assert(sizeof(Pose2d) == 6 * sizeof(double), "cams[] is 4 x 6 contiguous doubles");
    const double* cams = &model.cams[0].m[0][0];
    return hd::fourview2d_error(cams, x1_[i].data(), x2_[i].data(), x3_[i].data(), x4_[i].data(),
                                model.X[i].data());
  }

  // ---- batched scoring (optional; the GPU path of ppsfm_initialize_reconstruction_gpu) --------
  // With a scorer attached, MinimalSolver leaves X empty ("lazy" models: only the cameras are
  // known), ScoreModels scores all candidates of a sample in one call — every track triangulated
  // and evaluated on the device, the sums in track order like LocallyOptimizedMSAC::Score — and
  // Materialize triangulates the tracks of the one model that is kept.
  void set_batch_scorer(const BatchScorer* scorer) { scorer_ = scorer; }
  static void PackCams(const Reconstruction& model, double* cams) {
    for (int v = 0; v < 4; ++v)
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) cams[6 * v + 3 * r + c] = model.cams[v].m[r][c];
  }
  bool ScoreModels(const ReconstructionVector& models, int num_models, double threshold,
                   double* scores, bool threshold_first = false) const {
    if (!scorer_ || num_models <= 0) return false;
    std::vector<double> cams(24 * static_cast<size_t>(num_models));
    for (int m = 0; m < num_models; ++m) PackCams(models[m], &cams[24 * static_cast<size_t>(m)]);
    return scorer_->ScoreFourView2d(cams.data(), num_models, threshold, threshold_first, scores);
  }
  void Materialize(Reconstruction* model) const {
    if (!model->X.empty()) return;
    detail::three_view_triangulate2d(model->cams[0], model->cams[1], model->cams[2], x1_, x2_, x3_,
                                     &model->X);
  }

  // sfm2d.cc:321-361.  compact: X_ holds only the sample's points, in sample order (lazy models)
  int AbsPoseSolver(const std::vector<int>& sample, const std::vector<Vec2>& x_,
                    const std::vector<Vec2>& X_, Pose2d* model, bool compact = false) const {
    const int n = static_cast<int>(sample.size());
    std::vector<double> A(2 * n), B(2 * n);
    double btb[3] = {0, 0, 0}, bta[4] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      const double x1 = x_[sample[i]][0], x2 = x_[sample[i]][1];
      const Vec2& Xi = X_[compact ? i : sample[i]];
      const double X1 = Xi[0], X2 = Xi[1];
      A[2 * i] = X1 * x2 - X2 * x1;
      A[2 * i + 1] = -X1 * x1 - X2 * x2;
      B[2 * i] = x2;
      B[2 * i + 1] = -x1;
      btb[0] += B[2 * i] * B[2 * i];
      btb[1] += B[2 * i] * B[2 * i + 1];
      btb[2] += B[2 * i + 1] * B[2 * i + 1];
      bta[0] += B[2 * i] * A[2 * i];
      bta[1] += B[2 * i] * A[2 * i + 1];
      bta[2] += B[2 * i + 1] * A[2 * i];
      bta[3] += B[2 * i + 1] * A[2 * i + 1];
    }
    const double det = btb[0] * btb[2] - btb[1] * btb[1];
    const double i00 = btb[2] / det, i01 = -btb[1] / det, i11 = btb[0] / det;
    const double C[4] = {-(i00 * bta[0] + i01 * bta[2]), -(i00 * bta[1] + i01 * bta[3]),
                         -(i01 * bta[0] + i11 * bta[2]), -(i01 * bta[1] + i11 * bta[3])};
    std::vector<double> M(2 * n);
    for (int i = 0; i < n; ++i) {
      M[2 * i] = A[2 * i] + B[2 * i] * C[0] + B[2 * i + 1] * C[2];
      M[2 * i + 1] = A[2 * i + 1] + B[2 * i] * C[1] + B[2 * i + 1] * C[3];
    }
    double V[4], sv[2];
    la::svd_right_vectors(M, n, 2, V, sv);
    double ab[2] = {V[1], V[3]};
    const double nab = std::sqrt(ab[0] * ab[0] + ab[1] * ab[1]);
    ab[0] /= nab;
    ab[1] /= nab;
    const double t[2] = {C[0] * ab[0] + C[1] * ab[1], C[2] * ab[0] + C[3] * ab[1]};
    model->m[0][0] = ab[0]; model->m[1][1] = ab[0];
    model->m[0][1] = -ab[1]; model->m[1][0] = ab[1];
    model->m[0][2] = t[0]; model->m[1][2] = t[1];
    const Vec2& X0 = X_[compact ? 0 : sample[0]];
    if (model->m[1][0] * X0[0] + model->m[1][1] * X0[1] + model->m[1][2] < 0)
      for (auto& row : model->m)
        for (double& e : row) e *= -1.0;
    return 1;
  }

  // sfm2d.cc:363-444
  int MinimalSolver(const std::vector<int>& sample, ReconstructionVector* models) const {
    const int n = static_cast<int>(sample.size());
    std::vector<double> A(6 * n);
    for (int i = 0; i < n; ++i) {
      const double a1 = x1_[sample[i]][0], a2 = x1_[sample[i]][1];
      const double b1 = x2_[sample[i]][0], b2 = x2_[sample[i]][1];
      const double c1 = x3_[sample[i]][0], c2 = x3_[sample[i]][1];
      double* row = &A[6 * i];
      row[0] = a1 * b2 * c1 - a2 * b1 * c1;
      row[1] = a1 * b1 * c1 + a2 * b2 * c1;
      row[2] = a1 * b1 * c2 - a2 * b1 * c1;
      row[3] = a1 * b1 * c1 + a2 * b1 * c2;
      row[4] = a1 * b1 * c1 + a1 * b2 * c2;
      row[5] = a2 * b1 * c1 + a2 * b2 * c2;
    }
    double V[36], sv[6];
    la::svd_right_vectors(A, n, 6, V, sv);
    double t[6];
    for (int r = 0; r < 6; ++r) t[r] = V[r * 6 + 5];
    double tensor[8];
    tensor[0] = t[1] + t[3] + t[4];
    tensor[1] = -t[2] - t[0] + t[5];
    for (int r = 0; r < 6; ++r) tensor[2 + r] = t[r];
    Pose2d P1[2], P2[2], P3[2];
    const int n_fact = detail::factorize_trifocal_tensor(tensor, P1, P2, P3);
    if (n_fact == 0) return 0;
    models->clear();
    for (int f = 0; f < n_fact; ++f) {
      double H[3][3];
      detail::metric_upgrade(P2[f], P3[f], H);
      auto times_H = [&](const Pose2d& P) {
        Pose2d o;
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3; ++c)
            o.m[r][c] = P.m[r][0] * H[0][c] + P.m[r][1] * H[1][c] + P.m[r][2] * H[2][c];
        return o;
      };
      auto scale_all = [](Pose2d* P, double s) {
        for (auto& row : P->m)
          for (double& e : row) e *= s;
      };
      auto col_norm = [](const Pose2d& P, int c) {
        return std::sqrt(P.m[0][c] * P.m[0][c] + P.m[1][c] * P.m[1][c]);
      };
      const Pose2d P1f = P1[f];
      Pose2d P2f = times_H(P2[f]), P3f = times_H(P3[f]);
      scale_all(&P2f, 1.0 / col_norm(P2f, 0));
      scale_all(&P3f, 1.0 / col_norm(P3f, 0));
      const double s = col_norm(P2f, 2);
      for (int r = 0; r < 2; ++r) {
        P2f.m[r][2] /= s;
        P3f.m[r][2] /= s;
      }
      for (int flip1 = 0; flip1 < 2; ++flip1)
        for (int flip2 = 0; flip2 < 2; ++flip2)
          for (int flip3 = 0; flip3 < 2; ++flip3) {
            Reconstruction rec;
            rec.cams.resize(4);
            rec.cams[0] = P1f; rec.cams[1] = P2f; rec.cams[2] = P3f;
            const double n1 = col_norm(rec.cams[1], 2);
            for (int r = 0; r < 2; ++r) rec.cams[2].m[r][2] /= n1;
            const double n1b = col_norm(rec.cams[1], 2);
            for (int r = 0; r < 2; ++r) rec.cams[1].m[r][2] /= n1b;
            if (flip1)
              for (int r = 0; r < 2; ++r) {
                rec.cams[1].m[r][2] *= -1.0;
                rec.cams[2].m[r][2] *= -1.0;
              }
            if (flip2) scale_all(&rec.cams[1], -1.0);
            if (flip3) scale_all(&rec.cams[2], -1.0);
            if (scorer_) {  // lazy: only the sample's points, for the fourth camera
              std::vector<Vec2> Xs(sample.size());
              for (size_t i = 0; i < sample.size(); ++i)
                hd::triangulate2d_point(&rec.cams[0].m[0][0], &rec.cams[1].m[0][0],
                                        &rec.cams[2].m[0][0], x1_[sample[i]].data(),
                                        x2_[sample[i]].data(), x3_[sample[i]].data(), Xs[i].data());
              AbsPoseSolver(sample, x4_, Xs, &rec.cams[3], true);
            } else {
              detail::three_view_triangulate2d(rec.cams[0], rec.cams[1], rec.cams[2], x1_, x2_, x3_, &rec.X);
              AbsPoseSolver(sample, x4_, rec.X, &rec.cams[3]);
            }
            models->push_back(rec);
          }
    }
    return static_cast<int>(models->size());
  }

  // sfm2d.cc:446-467
  int NonMinimalSolver(const std::vector<int>& sample, Reconstruction* model) const {
    ReconstructionVector models;
    MinimalSolver(sample, &models);
    double best = std::numeric_limits<double>::max();
    std::vector<double> scores(models.size());
    // (std::min(threshold, error) here, std::min(error, threshold) in the driver)
    const bool batched = ScoreModels(models, static_cast<int>(models.size()), inlier_threshold_,
                                     scores.data(), true);
    for (size_t mi = 0; mi < models.size(); ++mi) {
      if (!batched) Materialize(&models[mi]);
      const Reconstruction& m = models[mi];
      double score = 0;
      if (batched) {
        score = scores[mi];
      } else {
        for (int j = 0, nd = num_data(); j < nd; ++j)
          score += std::min(inlier_threshold_, EvaluateModelOnPoint(m, j));
      }
      if (score < best) {
        best = score;
        *model = m;
      }
    }
    if (!models.empty()) Materialize(model);
    return models.empty() ? 0 : 1;
  }

  // sfm2d.cc:469-489
  void LeastSquares(const std::vector<int>& sample, Reconstruction* model) const {
    std::vector<std::vector<Vec2>> x(4);
    std::vector<Vec2> X;
    for (int s : sample) {
      x[0].push_back(x1_[s]); x[1].push_back(x2_[s]);
      x[2].push_back(x3_[s]); x[3].push_back(x4_[s]);
      X.push_back(model->X[s]);
    }
    detail::bundle_adjust2d(&model->cams, x, &X);
    for (size_t i = 0; i < sample.size(); ++i) model->X[sample[i]] = X[i];
    const std::vector<std::vector<Vec2>> all{x1_, x2_, x3_, x4_};
    detail::optimize_points2d(model->cams, all, &model->X);
  }

  const std::vector<Vec2>& x(int view) const {
    return view == 0 ? x1_ : view == 1 ? x2_ : view == 2 ? x3_ : x4_;
  }

 private:
  std::vector<Vec2> x1_, x2_, x3_, x4_;
  const double inlier_threshold_;
  const BatchScorer* scorer_ = nullptr;
};

class AbsolutePose2dEstimator {  // sfm2d.h:102-148
 public:
  typedef std::vector<Pose2d> Pose2dVector;
  AbsolutePose2dEstimator(const std::vector<Vec2>& x, const std::vector<Vec2>& X) : x_(x), X_(X) {
    for (Vec2& v : x_) {
      const double n = la::norm2(v);
      v[0] /= n;
      v[1] /= n;
    }
  }
  int min_sample_size() const { return 3; }
  int non_minimal_sample_size() const { return 2 * min_sample_size(); }
  int num_data() const { return static_cast<int>(x_.size()); }
  int MinimalSolver(const std::vector<int>& sample, Pose2dVector* models) const {
    Pose2d cam;
    NonMinimalSolver(sample, &cam);
    models->clear();
    models->push_back(cam);
    return 1;
  }
  // sfm2d.cc:491-522
  int NonMinimalSolver(const std::vector<int>& sample, Pose2d* model) const {
    const int n = static_cast<int>(sample.size());
    std::vector<double> A(4 * n);
    for (int i = 0; i < n; ++i) {
      const double x1 = x_[sample[i]][0], x2 = x_[sample[i]][1];
      const double X1 = X_[sample[i]][0], X2 = X_[sample[i]][1];
      A[4 * i] = X1 * x2 - X2 * x1;
      A[4 * i + 1] = -X1 * x1 - X2 * x2;
      A[4 * i + 2] = x2;
      A[4 * i + 3] = -x1;
    }
    double V[16], sv[4];
    la::svd_right_vectors(A, n, 4, V, sv);
    double t[4] = {V[3], V[7], V[11], V[15]};
    const double nt = std::sqrt(t[0] * t[0] + t[1] * t[1]);
    for (double& e : t) e /= nt;
    model->m[0][0] = t[0]; model->m[1][1] = t[0];
    model->m[0][1] = -t[1]; model->m[1][0] = t[1];
    model->m[0][2] = t[2]; model->m[1][2] = t[3];
    const Vec2& X0 = X_[sample[0]];
    if (model->m[1][0] * X0[0] + model->m[1][1] * X0[1] + model->m[1][2] < 0)
      for (auto& row : model->m)
        for (double& e : row) e *= -1.0;
    return 1;
  }
  // cosine error, sfm2d.cc:524-529
  double EvaluateModelOnPoint(const Pose2d& model, int i) const {
    Vec2 z = la::apply(model, X_[i]);
    const double n = la::norm2(z);
    return 1.0 - (x_[i][0] * z[0] + x_[i][1] * z[1]) / n;
  }
  void LeastSquares(const std::vector<int>& sample, Pose2d* model) const {
    NonMinimalSolver(sample, model);
  }

 private:
  std::vector<Vec2> x_, X_;
};

// initializer.cc:219-232
inline void four_view_triangulate(const std::vector<Pose>& cams,
                                  const std::vector<std::vector<Vec3>>& lines, std::vector<Vec3>* X) {
  double P[48];  // (per track: hd::triangulate3d_point, shared with the GPU scoring kernel)
  for (int j = 0; j < 4; ++j)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) P[12 * j + 4 * r + c] = cams[j].m[r][c];
  for (size_t i = 0; i < lines[0].size(); ++i) {
    Vec3 sol;
    hd::triangulate3d_point(P, lines[0][i].data(), lines[1][i].data(), lines[2][i].data(),
                            lines[3][i].data(), sol.data());
    X->push_back(sol);
  }
}

class PlanarOffsetEstimator {  // initializer.h:62-101
 public:
  struct Reconstruction {
    std::vector<Pose> cams;
    std::vector<Vec3> X;
  };
  typedef std::vector<Reconstruction> ReconstructionVector;

  PlanarOffsetEstimator(const std::vector<Pose>& poses, const std::vector<std::vector<Vec3>>& lines,
                        const std::vector<Mat3>& Rg, double inlier_threshold)
      : poses_(poses), lines_(lines), Rg_(Rg), inlier_threshold_(inlier_threshold) {}
  int min_sample_size() const { return 3; }
  int non_minimal_sample_size() const { return 20; }
  int num_data() const { return static_cast<int>(lines_[0].size()); }

  // initializer.cc:236-281: linear solve for the out-of-plane translations of cameras 1..3
  int MinimalSolver(const std::vector<int>& sample, ReconstructionVector* models) const {
    const int n = static_cast<int>(sample.size());
    std::vector<double> A(3 * n), b(n);
    for (int i = 0; i < n; ++i) {
      Mat3 A0;
      double B0[12] = {0};
      for (int j = 1; j < 4; ++j) {
        const Vec3 lg = la::mul(Rg_[j], lines_[j][sample[i]]);
        for (int c = 0; c < 3; ++c)
          A0.m[j - 1][c] = lg[0] * poses_[j].m[0][c] + lg[1] * poses_[j].m[1][c] + lg[2] * poses_[j].m[2][c];
        B0[(j - 1) * 4 + (j - 1)] = lg[1];
        B0[(j - 1) * 4 + 3] = lg[0] * poses_[j].m[0][3] + lg[2] * poses_[j].m[2][3];
      }
      la::lu3_solve(A0, B0, 4);
      double RB[12];  // Rg_[0]^T * B0
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
          RB[r * 4 + c] = Rg_[0].m[0][r] * B0[c] + Rg_[0].m[1][r] * B0[4 + c] + Rg_[0].m[2][r] * B0[8 + c];
      const Vec3& l0 = lines_[0][sample[i]];
      for (int c = 0; c < 3; ++c) A[3 * i + c] = l0[0] * RB[c] + l0[1] * RB[4 + c] + l0[2] * RB[8 + c];
      b[i] = -(l0[0] * RB[3] + l0[1] * RB[7] + l0[2] * RB[11]);
    }
    double tt[3];
    la::qr_solve(A, n, 3, b, tt);
    Reconstruction rec;
    rec.cams.resize(4);
    for (int i = 0; i < 4; ++i) {
      Pose P = poses_[i];
      if (i > 0) P.m[1][3] = tt[i - 1];
      for (int r = 0; r < 3; ++r)  // Rg^T * P
        for (int c = 0; c < 4; ++c)
          rec.cams[i].m[r][c] = Rg_[i].m[0][r] * P.m[0][c] + Rg_[i].m[1][r] * P.m[1][c] + Rg_[i].m[2][r] * P.m[2][c];
    }
    if (!scorer_) four_view_triangulate(rec.cams, lines_, &rec.X);  // (else lazy: see Materialize)
    models->clear();
    models->push_back(rec);
    return 1;
  }

  // batched scoring, as in FourView2dEstimator
  void set_batch_scorer(const BatchScorer* scorer) { scorer_ = scorer; }
  static void PackCams(const Reconstruction& model, double* cams) {
    for (int v = 0; v < 4; ++v)
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) cams[12 * v + 4 * r + c] = model.cams[v].m[r][c];
  }
  bool ScoreModels(const ReconstructionVector& models, int num_models, double threshold,
                   double* scores, bool threshold_first = false) const {
    if (!scorer_ || num_models <= 0) return false;
    std::vector<double> cams(48 * static_cast<size_t>(num_models));
    for (int m = 0; m < num_models; ++m) PackCams(models[m], &cams[48 * static_cast<size_t>(m)]);
    return scorer_->ScorePlanarOffset(cams.data(), num_models, threshold, threshold_first, scores);
  }
  void Materialize(Reconstruction* model) const {
    if (model->X.empty()) four_view_triangulate(model->cams, lines_, &model->X);
  }

  // initializer.cc:283-308
  int NonMinimalSolver(const std::vector<int>& sample, Reconstruction* model) const {
    ReconstructionVector models;
    MinimalSolver(sample, &models);
    double best = std::numeric_limits<double>::max();
    std::vector<double> scores(models.size());
    const bool batched = ScoreModels(models, static_cast<int>(models.size()), inlier_threshold_,
                                     scores.data());
    for (size_t mi = 0; mi < models.size(); ++mi) {
      if (!batched) Materialize(&models[mi]);
      const Reconstruction& m = models[mi];
      double score = 0;
      if (batched) {
        score = scores[mi];
      } else {
        for (int j = 0, nd = num_data(); j < nd; ++j)
          score += std::min(EvaluateModelOnPoint(m, j), inlier_threshold_);
      }
      if (score < best) {
        best = score;
        *model = m;
      }
    }
    if (models.empty()) return 0;
    Materialize(model);
    LeastSquares(sample, model);
    return 1;
  }

  // initializer.cc:310-333: max over the views of the point-to-line distance (normalised plane)
  double EvaluateModelOnPoint(const Reconstruction& model, int i) const {
    static_assert(sizeof(Pose) == 12 * sizeof(double), "cams[] is 4 x 12 contiguous doubles");
    const double* cams = &model.cams[0].m[0][0];
    return hd::planar_offset_error(cams, lines_[0][i].data(), lines_[1][i].data(),
                                   lines_[2][i].data(), lines_[3][i].data(), model.X[i].data());
  }

  // the reference returns before doing anything (initializer.cc:450-451)
  void LeastSquares(const std::vector<int>&, Reconstruction*) const {}

 private:
  std::vector<Pose> poses_;
  std::vector<std::vector<Vec3>> lines_;
  std::vector<Mat3> Rg_;
  const double inlier_threshold_;
  const BatchScorer* scorer_ = nullptr;
};

inline void lift_camera(const Pose2d& p, Pose* out) {  // initializer.cc:45-55
  *out = Pose();
  out->m[0][0] = p.m[0][0]; out->m[0][2] = p.m[0][1];
  out->m[2][0] = p.m[1][0]; out->m[2][2] = p.m[1][1];
  out->m[1][1] = 1.0;
  out->m[0][3] = p.m[0][2];
  out->m[2][3] = p.m[1][2];
}

// One image's lifted lines: line (a, b, c) in normalised camera coordinates and the aligned flag
// (FeatureLine, src/feature/types.h:98-138).
struct ImageLines {
  std::vector<Vec3> line;
  std::vector<unsigned char> aligned;
};

struct InitReport {  // what the reference prints; kept for tests
  int num_aligned = 0, num_unaligned = 0;
  int inliers_2d = 0, inliers_3d = 0;
  uint32_t iterations_2d = 0, iterations_3d = 0;
  double mean_tri_angle_deg = 0;
};

// initialize_reconstruction (initializer.cc:57-215), templated on the LO-MSAC driver so that the
// same estimators run under ppsfm::LocallyOptimizedMSAC and the reference's ransac_lib driver.
// Contract violations the reference CHECK-aborts on return false with *error set.
template <class LomsacTraits>
inline bool initialize_reconstruction_t(const std::vector<ImageLines>& lines,
                                        const std::vector<Vec3>& gravity, const InitOptions& options,
                                        std::vector<Pose>* output, double* inlier_ratio,
                                        InitReport* report = nullptr, const char** error = nullptr,
                                        BatchScorerFactory* scorers = nullptr) {
  *inlier_ratio = 0;
  if (error) *error = nullptr;
  auto fail = [&](const char* msg) {
    if (error) *error = msg;
    return false;
  };
  if (lines.size() != 4 || gravity.size() != 4) return fail("four images are required");
  std::vector<std::vector<Vec2>> x(4);
  std::vector<std::vector<Vec3>> lines_r(4);
  std::vector<Mat3> Rg(4);
  for (int i = 0; i < 4; ++i) {
    Rg[i] = la::rotation_from_two_vectors(gravity[i], Vec3{0.0, 1.0, 0.0});
    for (size_t j = 0; j < lines[i].line.size(); ++j) {
      Vec3 l = lines[i].line[j];
      if (lines[i].aligned[j]) {
        l = la::mul(Rg[i], l);
        if (std::fabs(l[1]) > 1e-6) return fail("aligned line is not parallel to gravity");  // CHECK_NEAR :83
        Vec2 xl{l[2], -l[0]};
        if (xl[1] < 0) {
          xl[0] = -xl[0];
          xl[1] = -xl[1];
        }
        const double n = la::norm2(xl);
        x[i].push_back(Vec2{xl[0] / n, xl[1] / n});
      } else {
        lines_r[i].push_back(l);
      }
    }
  }
  for (int i = 1; i < 4; ++i)
    if (x[i].size() != x[0].size() || lines_r[i].size() != lines_r[0].size())
      return fail("the four images must have the same aligned / unaligned split");  // CHECK_EQ :99-104
  InitReport local;
  InitReport& rep = report ? *report : local;
  rep.num_aligned = static_cast<int>(x[0].size());
  rep.num_unaligned = static_cast<int>(lines_r[0].size());

  typename LomsacTraits::Options ransac_options;
  ransac_options.final_least_squares_ = true;
  ransac_options.min_num_iterations_ = 1000;
  ransac_options.squared_inlier_threshold_ = options.max_error;
  FourView2dEstimator solver(x[0], x[1], x[2], x[3], ransac_options.squared_inlier_threshold_);
  if (scorers && !x[0].empty()) {  // (the estimator's own, normalised, observations)
    const double* obs[4];
    for (int v = 0; v < 4; ++v) obs[v] = solver.x(v).data()->data();
    const BatchScorer* sc = scorers->FourView2d(obs, static_cast<int>(x[0].size()));
    if (!sc) return fail("GPU scorer of the four-view solver could not be set up");
    solver.set_batch_scorer(sc);
  }
  typename LomsacTraits::template Driver<FourView2dEstimator::Reconstruction,
                                         FourView2dEstimator::ReconstructionVector,
                                         FourView2dEstimator> fourview_ransac;
  typename LomsacTraits::Stats stats;
  FourView2dEstimator::Reconstruction rec;
  int inliers = fourview_ransac.EstimateModel(ransac_options, solver, &rec, &stats);
  rep.inliers_2d = inliers;
  rep.iterations_2d = stats.num_iterations;
  if (inliers < options.min_num_inliers) return false;

  // mean (over the inliers) of the smallest triangulation angle among the first three views
  double angle_sum = 0;
  for (int idx : stats.inlier_indices) {
    double min_angle = std::numeric_limits<double>::max();
    for (int c1 = 0; c1 < 3; ++c1) {
      const Pose2d& A = rec.cams[c1];
      const Vec2 ca{-(A.m[0][0] * A.m[0][2] + A.m[1][0] * A.m[1][2]),
                    -(A.m[0][1] * A.m[0][2] + A.m[1][1] * A.m[1][2])};
      for (int c2 = c1 + 1; c2 < 3; ++c2) {
        const Pose2d& B = rec.cams[c2];
        const Vec2 cb{-(B.m[0][0] * B.m[0][2] + B.m[1][0] * B.m[1][2]),
                      -(B.m[0][1] * B.m[0][2] + B.m[1][1] * B.m[1][2])};
        const Vec2 v1{ca[0] - rec.X[idx][0], ca[1] - rec.X[idx][1]};
        const Vec2 v2{cb[0] - rec.X[idx][0], cb[1] - rec.X[idx][1]};
        const double angle =
            std::acos((v1[0] * v2[0] + v1[1] * v2[1]) / (la::norm2(v1) * la::norm2(v2)));
        if (angle < min_angle) min_angle = angle;
      }
    }
    angle_sum += min_angle;
  }
  rep.mean_tri_angle_deg = (angle_sum / stats.inlier_indices.size()) / M_PI * 180.0;
  if (rep.mean_tri_angle_deg < options.min_tri_angle) return false;

  std::vector<Pose> poses(4);
  for (int i = 0; i < 4; ++i) lift_camera(rec.cams[i], &poses[i]);
  typename LomsacTraits::Options planar_options;
  planar_options.final_least_squares_ = true;
  planar_options.min_num_iterations_ = 1000;
  planar_options.squared_inlier_threshold_ = options.max_error;
  PlanarOffsetEstimator planar_solver(poses, lines_r, Rg, planar_options.squared_inlier_threshold_);
  if (scorers && !lines_r[0].empty()) {
    const double* obs[4];
    for (int v = 0; v < 4; ++v) obs[v] = lines_r[v].data()->data();
    const BatchScorer* sc = scorers->PlanarOffset(obs, static_cast<int>(lines_r[0].size()));
    if (!sc) return fail("GPU scorer of the planar-offset solver could not be set up");
    planar_solver.set_batch_scorer(sc);
  }
  typename LomsacTraits::template Driver<PlanarOffsetEstimator::Reconstruction,
                                         PlanarOffsetEstimator::ReconstructionVector,
                                         PlanarOffsetEstimator> planar_ransac;
  PlanarOffsetEstimator::Reconstruction rec3d;
  inliers = planar_ransac.EstimateModel(planar_options, planar_solver, &rec3d, &stats);
  rep.inliers_3d = inliers;
  rep.iterations_3d = stats.num_iterations;
  if (inliers < options.min_tri_angle) return false;  // (sic) initializer.cc:207
  *output = rec3d.cams;
  *inlier_ratio = stats.inlier_ratio;
  return inliers >= options.min_num_inliers;
}

struct PpsfmLomsac {
  using Options = ppsfm::LORansacOptions;
  using Stats = ppsfm::RansacStatistics;
  template <class M, class MV, class S>
  using Driver = ppsfm::LocallyOptimizedMSAC<M, MV, S>;
};

inline bool initialize_reconstruction(const std::vector<ImageLines>& lines,
                                      const std::vector<Vec3>& gravity, const InitOptions& options,
                                      std::vector<Pose>* output, double* inlier_ratio,
                                      InitReport* report = nullptr, const char** error = nullptr,
                                      BatchScorerFactory* scorers = nullptr) {
  return initialize_reconstruction_t<PpsfmLomsac>(lines, gravity, options, output, inlier_ratio,
                                                  report, error, scorers);
}

}  // namespace init
}  // namespace ppsfm
