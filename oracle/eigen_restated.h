// eigen_restated.h — TEST INFRASTRUCTURE ONLY.
//
// The Eigen boundary of the oracle in ONE place: the four Eigen calls the reference's P6L / re3q3
// path delegates to the library (Eigen3 is absent here and unpinned upstream), restated from the
// published algorithms / Eigen 3.3's sources from memory:
//   * Matrix3d::determinant()                 (re3q3.h:24-26, absolute_pose.cc:126)
//   * PartialPivLU<Matrix3d>::solve           (absolute_pose.cc:137, re3q3.h:71,75,79)
//   * EigenSolver<Matrix<double,8,8>>         (re3q3.h:164-165)
//   * Quaterniond(Matrix3d)                   (base/pose.cc:41-44, used by estimators/pose.cc:86)
//   * JacobiSVD<Matrix<double,Dynamic,4>>     (base/triangulation.cc:51-53: only matrixV().col(3))
// Included by ppsfm_oracle.cc (the restatement) AND by the Eigen stand-in under oracle/ref/shim/
// that lets the reference's own sources compile (oracle/build_ref.sh -> oracle/_ref/libref_p6l.so),
// so both sides of tests/test_ref_p6l.py share exactly these operations and everything ELSE the
// test compares is the reference's source text against the restatement.  Bit-level parity with a
// real Eigen build stays UNPINNED for what is in this file.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <cfloat>
#include <utility>
#include <vector>

namespace eigen_restated {

// ---------------------------------------------------------------------------------------------
// Small dense helpers (stand-ins for the Eigen calls of the reference).
// Matrices are row-major double[r][c] unless stated.
// ---------------------------------------------------------------------------------------------

// Eigen 3x3 determinant (cofactor expansion along the first row).
inline double Det3(const double m[3][3]) {
  const double h0 = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]);
  const double h1 = m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]);
  const double h2 = m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  return h0 - h1 + h2;
}

// Solve A X = B for 3x3 A: Eigen::PartialPivLU<Matrix3d>(A).solve(B), which is what both
// `B.transpose().partialPivLu().solve(tt)` (absolute_pose.cc:137) and `Ax.lu().solve(P)`
// (re3q3.h:71,75,79 -- in Eigen 3 MatrixBase::lu() is a synonym of partialPivLu()) evaluate.
// Operation order restated from Eigen 3.3's sources (the library itself is absent here, so this
// is UNPINNED at the bit level): LU/PartialPivLU.h `unblocked_lu` (first maximum of |column| is the
// pivot, whole rows swapped, multipliers by true division, rank-1 update of the trailing block),
// then `solve` = row permutation, unit-lower and upper triangular solves by
// products/TriangularSolverMatrix.h (right-looking per pivot: x_i = b_i * (1 / u_ii), then
// b_r -= x_i * u_ri for the remaining rows).  B is 3 x NC, overwritten by X.
template <int NC>
void SolvePartialPiv3(double A[3][3], double B[3][NC]) {
  int piv[3];
  for (int k = 0; k < 3; ++k) {
    int p = k;
    double best = std::fabs(A[k][k]);
    for (int i = k + 1; i < 3; ++i) {
      const double v = std::fabs(A[i][k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    piv[k] = p;
    if (best != 0.0) {
      if (p != k)
        for (int j = 0; j < 3; ++j) std::swap(A[k][j], A[p][j]);
      for (int i = k + 1; i < 3; ++i) A[i][k] = A[i][k] / A[k][k];
    }
    for (int i = k + 1; i < 3; ++i)
      for (int j = k + 1; j < 3; ++j) A[i][j] = A[i][j] - A[i][k] * A[k][j];
  }
  for (int k = 0; k < 3; ++k)  // dst = P * rhs
    if (piv[k] != k)
      for (int j = 0; j < NC; ++j) std::swap(B[k][j], B[piv[k]][j]);
  for (int i = 0; i < 3; ++i)  // unit lower
    for (int j = 0; j < NC; ++j) {
      const double b = B[i][j];
      for (int r = i + 1; r < 3; ++r) B[r][j] = B[r][j] - b * A[r][i];
    }
  for (int i = 2; i >= 0; --i) {  // upper
    const double a = 1.0 / A[i][i];
    for (int j = 0; j < NC; ++j) {
      const double b = B[i][j] * a;
      B[i][j] = b;
      for (int r = 0; r < i; ++r) B[r][j] = B[r][j] - b * A[r][i];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Eigenvalues of the 8x8 companion matrix (re3q3.h:152-165: Eigen::EigenSolver<Matrix8d>).
// Real Schur form by Francis double-shift QR on an upper-Hessenberg matrix (EISPACK hqr /
// JAMA lineage, the algorithm inside Eigen::RealSchur); eigenvalues are read off the
// quasi-triangular diagonal top-to-bottom, which fixes the ORDER of the returned roots.
// Only the active window is updated (sufficient for eigenvalues; window values are identical).
// ---------------------------------------------------------------------------------------------
constexpr int kN = 8;
inline thread_local int g_last_qr_sweeps = 0;  // diagnostics only (orc_poly8_sweeps)

struct Hqr8 {
  double T[kN][kN];

  // Householder vector for a 3- or 2-vector (Eigen makeHouseholder): returns tau, beta, ess[].
  static void MakeHouseholder(const double* v, int n, double* ess, double* tau, double* beta) {
    double tail_sq = 0.0;
    for (int i = 1; i < n; ++i) tail_sq = tail_sq + v[i] * v[i];
    const double c0 = v[0];
    if (tail_sq <= std::numeric_limits<double>::min()) {
      *tau = 0.0;
      *beta = c0;
      for (int i = 0; i < n - 1; ++i) ess[i] = 0.0;
    } else {
      double b = std::sqrt(c0 * c0 + tail_sq);
      if (c0 >= 0.0) b = -b;
      for (int i = 0; i < n - 1; ++i) ess[i] = v[i + 1] / (c0 - b);
      *tau = (b - c0) / b;
      *beta = b;
    }
  }

  // rows r0..r0+ne (ne = #ess), columns c_lo..c_hi :  M <- (I - tau [1;ess][1;ess]^T) M
  void ApplyLeft(int r0, int ne, const double* ess, double tau, int c_lo, int c_hi) {
    if (tau == 0.0) return;
    for (int j = c_lo; j <= c_hi; ++j) {
      double tmp = ess[0] * T[r0 + 1][j];
      if (ne == 2) tmp = tmp + ess[1] * T[r0 + 2][j];
      tmp = tmp + T[r0][j];
      T[r0][j] = T[r0][j] - tau * tmp;
      T[r0 + 1][j] = T[r0 + 1][j] - (tau * ess[0]) * tmp;
      if (ne == 2) T[r0 + 2][j] = T[r0 + 2][j] - (tau * ess[1]) * tmp;
    }
  }
  // columns c0..c0+ne, rows r_lo..r_hi :  M <- M (I - tau [1;ess][1;ess]^T)
  void ApplyRight(int c0, int ne, const double* ess, double tau, int r_lo, int r_hi) {
    if (tau == 0.0) return;
    for (int i = r_lo; i <= r_hi; ++i) {
      double tmp = T[i][c0 + 1] * ess[0];
      if (ne == 2) tmp = tmp + T[i][c0 + 2] * ess[1];
      tmp = tmp + T[i][c0];
      T[i][c0] = T[i][c0] - tau * tmp;
      T[i][c0 + 1] = T[i][c0 + 1] - (tau * tmp) * ess[0];
      if (ne == 2) T[i][c0 + 2] = T[i][c0 + 2] - (tau * tmp) * ess[1];
    }
  }

  // Returns false if the iteration limit (40 per row) is hit.
  bool Reduce() {
    // Overall scaling by the largest magnitude (RealSchur::compute).
    double scale = 0.0;
    for (int i = 0; i < kN; ++i)
      for (int j = 0; j < kN; ++j) scale = std::max(scale, std::fabs(T[i][j]));
    if (!(scale > 0.0) || !std::isfinite(scale)) return std::isfinite(scale);
    for (int i = 0; i < kN; ++i)
      for (int j = 0; j < kN; ++j) T[i][j] = T[i][j] / scale;

    // The input is already upper Hessenberg (companion matrix) — no Householder reduction needed.
    double norm = 0.0;
    for (int j = 0; j < kN; ++j)
      for (int i = 0; i < std::min(kN, j + 2); ++i) norm = norm + std::fabs(T[i][j]);

    const int max_iters = 40 * kN;
    int iu = kN - 1, iter = 0, total_iter = 0;
    double exshift = 0.0;
    const double eps = std::numeric_limits<double>::epsilon();
    bool ok = true;
    if (norm != 0.0) {
      while (iu >= 0) {
        // findSmallSubdiagEntry
        int il = iu;
        while (il > 0) {
          const double s = std::fabs(T[il - 1][il - 1]) + std::fabs(T[il][il]);
          if (std::fabs(T[il][il - 1]) <= eps * s) break;
          --il;
        }
        if (il == iu) {  // one real root
          T[iu][iu] = T[iu][iu] + exshift;
          if (iu > 0) T[iu][iu - 1] = 0.0;
          --iu;
          iter = 0;
        } else if (il == iu - 1) {  // 2x2 block: split if its eigenvalues are real
          const double p = 0.5 * (T[iu - 1][iu - 1] - T[iu][iu]);
          const double q = p * p + T[iu][iu - 1] * T[iu - 1][iu];
          T[iu][iu] = T[iu][iu] + exshift;
          T[iu - 1][iu - 1] = T[iu - 1][iu - 1] + exshift;
          if (q >= 0.0) {
            const double z = std::sqrt(std::fabs(q));
            // Givens rotation G with G^T [a; b] = [r; 0], a = p +- z, b = T[iu][iu-1].
            const double a = (p >= 0.0) ? (p + z) : (p - z);
            const double b = T[iu][iu - 1];
            double c, s;
            if (b == 0.0) {
              c = (a < 0.0) ? -1.0 : 1.0;
              s = 0.0;
            } else if (a == 0.0) {
              c = 0.0;
              s = (b < 0.0) ? 1.0 : -1.0;
            } else if (std::fabs(a) > std::fabs(b)) {
              const double t = b / a;
              double u = std::sqrt(1.0 + t * t);
              if (a < 0.0) u = -u;
              c = 1.0 / u;
              s = -t * c;
            } else {
              const double t = a / b;
              double u = std::sqrt(1.0 + t * t);
              if (b < 0.0) u = -u;
              s = -1.0 / u;
              c = -t * s;
            }
            // rows (iu-1, iu) <- G^T rows  (x' = c x - s y ; y' = s x + c y), columns iu-1..iu
            for (int j = iu - 1; j <= iu; ++j) {
              const double x = T[iu - 1][j], y = T[iu][j];
              T[iu - 1][j] = c * x - s * y;
              T[iu][j] = s * x + c * y;
            }
            // columns (iu-1, iu) <- columns G, rows il..iu
            for (int i = iu - 1; i <= iu; ++i) {
              const double x = T[i][iu - 1], y = T[i][iu];
              T[i][iu - 1] = c * x - s * y;
              T[i][iu] = s * x + c * y;
            }
            T[iu][iu - 1] = 0.0;
          }
          if (iu > 1) T[iu - 1][iu - 2] = 0.0;
          iu -= 2;
          iter = 0;
        } else {
          // computeShift
          double sh0 = T[iu][iu];
          double sh1 = T[iu - 1][iu - 1];
          double sh2 = T[iu][iu - 1] * T[iu - 1][iu];
          if (iter == 10) {  // Wilkinson's ad hoc shift
            exshift = exshift + sh0;
            for (int i = 0; i <= iu; ++i) T[i][i] = T[i][i] - sh0;
            const double s = std::fabs(T[iu][iu - 1]) + std::fabs(T[iu - 1][iu - 2]);
            sh0 = 0.75 * s;
            sh1 = 0.75 * s;
            sh2 = -0.4375 * s * s;
          }
          if (iter == 30) {  // MATLAB's ad hoc shift
            double s = (sh1 - sh0) / 2.0;
            s = s * s + sh2;
            if (s > 0.0) {
              s = std::sqrt(s);
              if (sh1 < sh0) s = -s;
              s = s + (sh1 - sh0) / 2.0;
              s = sh0 - sh2 / s;
              exshift = exshift + s;
              for (int i = 0; i <= iu; ++i) T[i][i] = T[i][i] - s;
              sh0 = sh1 = sh2 = 0.964;
            }
          }
          ++iter;
          ++total_iter;
          if (total_iter > max_iters) {
            ok = false;
            break;
          }
          // initFrancisQRStep: look for two consecutive small sub-diagonal elements
          int im;
          double v[3] = {0.0, 0.0, 0.0};
          for (im = iu - 2; im >= il; --im) {
            const double Tmm = T[im][im];
            const double r = sh0 - Tmm;
            const double s = sh1 - Tmm;
            v[0] = (r * s - sh2) / T[im + 1][im] + T[im][im + 1];
            v[1] = T[im + 1][im + 1] - Tmm - r - s;
            v[2] = T[im + 2][im + 1];
            if (im == il) break;
            const double lhs = T[im][im - 1] * (std::fabs(v[1]) + std::fabs(v[2]));
            const double rhs = v[0] * (std::fabs(T[im - 1][im - 1]) + std::fabs(Tmm) +
                                       std::fabs(T[im + 1][im + 1]));
            if (std::fabs(lhs) < eps * rhs) break;
          }
          // performFrancisQRStep
          for (int k = im; k <= iu - 2; ++k) {
            const bool first = (k == im);
            double w[3];
            if (first) {
              w[0] = v[0];
              w[1] = v[1];
              w[2] = v[2];
            } else {
              w[0] = T[k][k - 1];
              w[1] = T[k + 1][k - 1];
              w[2] = T[k + 2][k - 1];
            }
            double ess[2], tau, beta;
            MakeHouseholder(w, 3, ess, &tau, &beta);
            if (beta != 0.0) {
              if (first && k > il)
                T[k][k - 1] = -T[k][k - 1];
              else if (!first)
                T[k][k - 1] = beta;
              ApplyLeft(k, 2, ess, tau, k, iu);
              ApplyRight(k, 2, ess, tau, il, std::min(iu, k + 3));
            }
          }
          {
            double w[2] = {T[iu - 1][iu - 2], T[iu][iu - 2]};
            double ess[1], tau, beta;
            MakeHouseholder(w, 2, ess, &tau, &beta);
            if (beta != 0.0) {
              T[iu - 1][iu - 2] = beta;
              ApplyLeft(iu - 1, 1, ess, tau, iu - 1, iu);
              ApplyRight(iu - 1, 1, ess, tau, il, iu);
            }
          }
          for (int i = im + 2; i <= iu; ++i) {
            T[i][i - 2] = 0.0;
            if (i > im + 2) T[i][i - 3] = 0.0;
          }
        }
      }
    }
    for (int i = 0; i < kN; ++i)
      for (int j = 0; j < kN; ++j) T[i][j] = T[i][j] * scale;
    g_last_qr_sweeps = total_iter;
    return ok;
  }

  // EigenSolver::compute eigenvalue extraction; returns count (8) and fills re/im.
  void Eigenvalues(double* re, double* im) const {
    int i = 0;
    while (i < kN) {
      if (i == kN - 1 || T[i + 1][i] == 0.0) {
        re[i] = T[i][i];
        im[i] = 0.0;
        ++i;
      } else {
        const double p = 0.5 * (T[i][i] - T[i + 1][i + 1]);
        double t0 = T[i + 1][i];
        double t1 = T[i][i + 1];
        const double maxval = std::max(std::fabs(p), std::max(std::fabs(t0), std::fabs(t1)));
        t0 = t0 / maxval;
        t1 = t1 / maxval;
        const double p0 = p / maxval;
        const double z = maxval * std::sqrt(std::fabs(p0 * p0 + t0 * t1));
        re[i] = T[i + 1][i + 1] + p;
        im[i] = z;
        re[i + 1] = T[i + 1][i + 1] + p;
        im[i + 1] = -z;
        i += 2;
      }
    }
  }
};

// c[0] x^8 + c[1] x^7 + ... + c[8]  ->  all eigenvalues of the companion matrix (re3q3.h:152-165)
inline bool Poly8Roots(const double* c, double* re, double* im) {
  Hqr8 h;
  for (int i = 0; i < kN; ++i)
    for (int j = 0; j < kN; ++j) h.T[i][j] = 0.0;
  for (int j = 0; j < kN; ++j) h.T[0][j] = -c[j + 1] / c[0];
  for (int i = 1; i < kN; ++i) h.T[i][i - 1] = 1.0;
  const bool ok = h.Reduce();
  h.Eigenvalues(re, im);
  return ok;
}

// Eigen::Quaterniond(Matrix3d) (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl):
// Ken Shoemake's trace-branch algorithm.  R is column-major: R(r,c) = R[3c + r]; q = (w, x, y, z).
inline void QuaternionFromRotationMatrix(const double* R, double* q) {
  auto at = [&](int r, int c) { return R[3 * c + r]; };
  double t = at(0, 0) + at(1, 1) + at(2, 2);
  double w, v[3];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    w = 0.5 * t;
    t = 0.5 / t;
    v[0] = (at(2, 1) - at(1, 2)) * t;
    v[1] = (at(0, 2) - at(2, 0)) * t;
    v[2] = (at(1, 0) - at(0, 1)) * t;
  } else {
    int i = 0;
    if (at(1, 1) > at(0, 0)) i = 1;
    if (at(2, 2) > at(i, i)) i = 2;
    const int j = (i + 1) % 3;
    const int k = (j + 1) % 3;
    t = std::sqrt(at(i, i) - at(j, j) - at(k, k) + 1.0);
    v[i] = 0.5 * t;
    t = 0.5 / t;
    w = (at(k, j) - at(j, k)) * t;
    v[j] = (at(j, i) + at(i, j)) * t;
    v[k] = (at(k, i) + at(i, k)) * t;
  }
  q[0] = w;
  q[1] = v[0];
  q[2] = v[1];
  q[3] = v[2];
}

// The right singular vector of the smallest singular value of an n x 4 matrix W (row-major,
// overwritten): one-sided Jacobi rotations of the columns until they are orthogonal, then the
// column of the accumulated V whose W-column is shortest.  (Eigen's JacobiSVD is two-sided with a
// QR preconditioner: same vector up to rounding and sign; UNPINNED like the rest of this file.)
inline void NullVectorNx4(std::vector<double>& W, int n, double v[4]) {
  double V[16];
  for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int r = 0; r < n; ++r) {
          alpha += W[4 * r + p] * W[4 * r + p];
          beta += W[4 * r + q] * W[4 * r + q];
          gamma += W[4 * r + p] * W[4 * r + q];
        }
        if (std::fabs(gamma) <= 1e-300 || std::fabs(gamma) <= 2.3e-16 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int r = 0; r < n; ++r) {
          const double wp = W[4 * r + p], wq = W[4 * r + q];
          W[4 * r + p] = c * wp - s * wq;
          W[4 * r + q] = s * wp + c * wq;
        }
        for (int r = 0; r < 4; ++r) {
          const double vp = V[4 * r + p], vq = V[4 * r + q];
          V[4 * r + p] = c * vp - s * vq;
          V[4 * r + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  int best = 0;
  double best_norm = DBL_MAX;
  for (int j = 0; j < 4; ++j) {
    double s2 = 0;
    for (int r = 0; r < n; ++r) s2 += W[4 * r + j] * W[4 * r + j];
    if (s2 < best_norm) { best_norm = s2; best = j; }
  }
  for (int r = 0; r < 4; ++r) v[r] = V[4 * r + best];
}

}  // namespace eigen_restated
