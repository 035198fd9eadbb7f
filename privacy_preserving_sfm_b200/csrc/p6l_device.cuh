// p6l_device.cuh — per-thread P6L minimal solver (6 line<->point correspondences -> <= 8 poses).
//
// Replaces P6LEstimator::Estimate (src/estimators/absolute_pose.cc:79-162) and re3q3
// (lib/re3q3/re3q3/re3q3.h:16-200) of the reference.  One CUDA thread solves one hypothesis.
//
// Numerical contract: this translation unit is compiled with --fmad=false and every expression
// below is evaluated in a fixed order, so that the models are bit-identical to the CPU oracle's
// (oracle/ppsfm_oracle.cc), which in turn follows the reference's FMA-free x86-64 build.
// The expressions the reference writes out itself (re3q3.h:84-150, :177-188) are generated code in
// the reference's own evaluation order (re3q3_resultant.inc).  The Eigen calls of the reference
// (3x3 determinant, PartialPivLU::solve -- `lu()` is its synonym --, EigenSolver<8x8>) are
// realised as: cofactor determinant, Eigen 3.3's unblocked row-pivoted LU + triangular solves, and
// Francis double-shift QR on the (already Hessenberg) companion matrix with eigenvalues read off
// the real Schur form top-to-bottom.
#pragma once
#include <cfloat>
#include <cstdint>

namespace ppsfm {
namespace dev {

#define PPSFM_DI __device__ __forceinline__

PPSFM_DI double det3(const double m[3][3]) {
  const double h0 = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]);
  const double h1 = m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]);
  const double h2 = m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  return h0 - h1 + h2;
}

PPSFM_DI void swapd(double& a, double& b) {
  const double t = a;
  a = b;
  b = t;
}

// A X = B, 3x3: Eigen::PartialPivLU<Matrix3d>(A).solve(B) -- absolute_pose.cc:137 and
// re3q3.h:71-79 (`lu()` is Eigen 3's synonym of partialPivLu()).  Operation order of Eigen 3.3's
// unblocked_lu + permutation + unit-lower / upper triangular solves (x_i = b_i * (1 / u_ii), then
// b_r -= x_i * u_ri), identical to oracle/ppsfm_oracle.cc SolvePartialPiv3.
template <int NC>
__device__ void solve_partial_piv3(double A[3][3], double B[3][NC]) {
  int piv[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int p = k;
    double best = fabs(A[k][k]);
#pragma unroll
    for (int i = k + 1; i < 3; ++i) {
      const double v = fabs(A[i][k]);
      if (v > best) {
        best = v;
        p = i;
      }
    }
    piv[k] = p;
    if (best != 0.0) {
#pragma unroll
      for (int i = k + 1; i < 3; ++i)
        if (p == i) {
#pragma unroll
          for (int j = 0; j < 3; ++j) swapd(A[k][j], A[i][j]);
        }
#pragma unroll
      for (int i = k + 1; i < 3; ++i) A[i][k] = A[i][k] / A[k][k];
    }
#pragma unroll
    for (int i = k + 1; i < 3; ++i)
#pragma unroll
      for (int j = k + 1; j < 3; ++j) A[i][j] = A[i][j] - A[i][k] * A[k][j];
  }
#pragma unroll
  for (int k = 0; k < 2; ++k)  // dst = P * rhs
#pragma unroll
    for (int i = k + 1; i < 3; ++i)
      if (piv[k] == i) {
#pragma unroll
        for (int j = 0; j < NC; ++j) swapd(B[k][j], B[i][j]);
      }
#pragma unroll
  for (int i = 0; i < 3; ++i)  // unit lower
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const double b = B[i][j];
#pragma unroll
      for (int r = i + 1; r < 3; ++r) B[r][j] = B[r][j] - b * A[r][i];
    }
#pragma unroll
  for (int i = 2; i >= 0; --i) {  // upper
    const double a = 1.0 / A[i][i];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      const double b = B[i][j] * a;
      B[i][j] = b;
#pragma unroll
      for (int r = 0; r < i; ++r) B[r][j] = B[r][j] - b * A[r][i];
    }
  }
}

// The hidden-variable resultant (re3q3.h:84-150) and the back-substitution for a root
// (:177-188): generated three-address code, one IEEE operation per statement in the reference's
// evaluation order (scripts/gen_re3q3_resultant.py); the oracle includes the same text.
#define RE3Q3_FN __device__ __forceinline__
#include "re3q3_resultant.inc"
#undef RE3Q3_FN

// ---- 8x8 Hessenberg QR (Francis double shift), eigenvalues only ------------------------------
struct Hqr8 {
  double T[8][8];

  __device__ static void make_householder(const double* v, int n, double* ess, double* tau,
                                          double* beta) {
    double tail_sq = 0.0;
    for (int i = 1; i < n; ++i) tail_sq = tail_sq + v[i] * v[i];
    const double c0 = v[0];
    if (tail_sq <= DBL_MIN) {
      *tau = 0.0;
      *beta = c0;
      for (int i = 0; i < n - 1; ++i) ess[i] = 0.0;
    } else {
      double b = sqrt(c0 * c0 + tail_sq);
      if (c0 >= 0.0) b = -b;
      for (int i = 0; i < n - 1; ++i) ess[i] = v[i + 1] / (c0 - b);
      *tau = (b - c0) / b;
      *beta = b;
    }
  }

  __device__ void apply_left(int r0, int ne, const double* ess, double tau, int c_lo, int c_hi) {
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
  __device__ void apply_right(int c0, int ne, const double* ess, double tau, int r_lo,
                              int r_hi) {
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

  __device__ bool reduce() {
    double scale = 0.0;
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) {
        const double v = fabs(T[i][j]);
        scale = (scale < v) ? v : scale;
      }
    if (!(scale > 0.0) || !isfinite(scale)) return isfinite(scale);
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) T[i][j] = T[i][j] / scale;

    double norm = 0.0;
    for (int j = 0; j < 8; ++j) {
      const int lim = (j + 2 < 8) ? (j + 2) : 8;
      for (int i = 0; i < lim; ++i) norm = norm + fabs(T[i][j]);
    }

    const int max_iters = 40 * 8;
    int iu = 7, iter = 0, total_iter = 0;
    double exshift = 0.0;
    const double eps = DBL_EPSILON;
    bool ok = true;
    if (norm != 0.0) {
      while (iu >= 0) {
        int il = iu;
        while (il > 0) {
          const double s = fabs(T[il - 1][il - 1]) + fabs(T[il][il]);
          if (fabs(T[il][il - 1]) <= eps * s) break;
          --il;
        }
        if (il == iu) {
          T[iu][iu] = T[iu][iu] + exshift;
          if (iu > 0) T[iu][iu - 1] = 0.0;
          --iu;
          iter = 0;
        } else if (il == iu - 1) {
          const double p = 0.5 * (T[iu - 1][iu - 1] - T[iu][iu]);
          const double q = p * p + T[iu][iu - 1] * T[iu - 1][iu];
          T[iu][iu] = T[iu][iu] + exshift;
          T[iu - 1][iu - 1] = T[iu - 1][iu - 1] + exshift;
          if (q >= 0.0) {
            const double z = sqrt(fabs(q));
            const double a = (p >= 0.0) ? (p + z) : (p - z);
            const double b = T[iu][iu - 1];
            double c, s;
            if (b == 0.0) {
              c = (a < 0.0) ? -1.0 : 1.0;
              s = 0.0;
            } else if (a == 0.0) {
              c = 0.0;
              s = (b < 0.0) ? 1.0 : -1.0;
            } else if (fabs(a) > fabs(b)) {
              const double t = b / a;
              double u = sqrt(1.0 + t * t);
              if (a < 0.0) u = -u;
              c = 1.0 / u;
              s = -t * c;
            } else {
              const double t = a / b;
              double u = sqrt(1.0 + t * t);
              if (b < 0.0) u = -u;
              s = -1.0 / u;
              c = -t * s;
            }
            for (int j = iu - 1; j <= iu; ++j) {
              const double x = T[iu - 1][j], y = T[iu][j];
              T[iu - 1][j] = c * x - s * y;
              T[iu][j] = s * x + c * y;
            }
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
          double sh0 = T[iu][iu];
          double sh1 = T[iu - 1][iu - 1];
          double sh2 = T[iu][iu - 1] * T[iu - 1][iu];
          if (iter == 10) {
            exshift = exshift + sh0;
            for (int i = 0; i <= iu; ++i) T[i][i] = T[i][i] - sh0;
            const double s = fabs(T[iu][iu - 1]) + fabs(T[iu - 1][iu - 2]);
            sh0 = 0.75 * s;
            sh1 = 0.75 * s;
            sh2 = -0.4375 * s * s;
          }
          if (iter == 30) {
            double s = (sh1 - sh0) / 2.0;
            s = s * s + sh2;
            if (s > 0.0) {
              s = sqrt(s);
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
            const double lhs = T[im][im - 1] * (fabs(v[1]) + fabs(v[2]));
            const double rhs =
                v[0] * (fabs(T[im - 1][im - 1]) + fabs(Tmm) + fabs(T[im + 1][im + 1]));
            if (fabs(lhs) < eps * rhs) break;
          }
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
            make_householder(w, 3, ess, &tau, &beta);
            if (beta != 0.0) {
              if (first && k > il)
                T[k][k - 1] = -T[k][k - 1];
              else if (!first)
                T[k][k - 1] = beta;
              apply_left(k, 2, ess, tau, k, iu);
              apply_right(k, 2, ess, tau, il, (iu < k + 3) ? iu : (k + 3));
            }
          }
          {
            double w[2] = {T[iu - 1][iu - 2], T[iu][iu - 2]};
            double ess[1], tau, beta;
            make_householder(w, 2, ess, &tau, &beta);
            if (beta != 0.0) {
              T[iu - 1][iu - 2] = beta;
              apply_left(iu - 1, 1, ess, tau, iu - 1, iu);
              apply_right(iu - 1, 1, ess, tau, il, iu);
            }
          }
          for (int i = im + 2; i <= iu; ++i) {
            T[i][i - 2] = 0.0;
            if (i > im + 2) T[i][i - 3] = 0.0;
          }
        }
      }
    }
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 8; ++j) T[i][j] = T[i][j] * scale;
    return ok;
  }

  __device__ void eigenvalues(double* re, double* im) const {
    int i = 0;
    while (i < 8) {
      if (i == 7 || T[i + 1][i] == 0.0) {
        re[i] = T[i][i];
        im[i] = 0.0;
        ++i;
      } else {
        const double p = 0.5 * (T[i][i] - T[i + 1][i + 1]);
        double t0 = T[i + 1][i];
        double t1 = T[i][i + 1];
        const double at0 = fabs(t0), at1 = fabs(t1), ap = fabs(p);
        const double m01 = (at0 < at1) ? at1 : at0;
        const double maxval = (ap < m01) ? m01 : ap;
        t0 = t0 / maxval;
        t1 = t1 / maxval;
        const double p0 = p / maxval;
        const double z = maxval * sqrt(fabs(p0 * p0 + t0 * t1));
        re[i] = T[i + 1][i + 1] + p;
        im[i] = z;
        re[i + 1] = T[i + 1][i + 1] + p;
        im[i + 1] = -z;
        i += 2;
      }
    }
  }
};

__device__ inline void poly8_roots(const double* c, double* re, double* im) {
  Hqr8 h;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) h.T[i][j] = 0.0;
  for (int j = 0; j < 8; ++j) h.T[0][j] = -c[j + 1] / c[0];
  for (int i = 1; i < 8; ++i) h.T[i][i - 1] = 1.0;
  h.reduce();
  h.eigenvalues(re, im);
}

// Fixed generic affine change of variables (stands in for the rand()-driven one, re3q3.h:41-42);
// identical constants in oracle/ppsfm_oracle.cc.
__device__ __constant__ double kVarChangeA[3][4] = {
    {-0.45264637943155561, -0.88862107060359552, 0.073917846740986226, 0.30304576336566319},
    {0.19122225569950785, -0.015767801546650473, 0.98142010645776834, -0.5050762722761053},
    {-0.8709450637742292, 0.45837099527970276, 0.1770613638646179, 0.80812203564176865}};

// Fixed substitute for setRandom() in the degenerate translation-block branch
// (absolute_pose.cc:128-134); identical constants in the oracle.
__device__ __constant__ double kMixA[3][3] = {{0.680375, -0.211234, 0.566198},
                                              {0.596880, 0.823295, -0.604897},
                                              {-0.329554, 0.536459, -0.444451}};

__device__ __constant__ int kRe3q3Cols[3][7] = {{0, 1, 2, 6, 7, 8, 9},
                                                {3, 1, 4, 7, 6, 8, 9},
                                                {5, 4, 2, 8, 7, 6, 9}};

// Core of re3q3 once the elimination variable is chosen (no change of variables).
__device__ int re3q3_core(const double coeffs[3][10], int elim_var, double solutions[3][8]) {
  double A[3][3], P[3][7];
  for (int k = 0; k < 3; ++k) {
    if (elim_var == 1) {
      A[k][0] = coeffs[k][3]; A[k][1] = coeffs[k][5]; A[k][2] = coeffs[k][4];
    } else if (elim_var == 2) {
      A[k][0] = coeffs[k][0]; A[k][1] = coeffs[k][5]; A[k][2] = coeffs[k][2];
    } else {
      A[k][0] = coeffs[k][3]; A[k][1] = coeffs[k][0]; A[k][2] = coeffs[k][1];
    }
    for (int j = 0; j < 7; ++j) P[k][j] = coeffs[k][kRe3q3Cols[elim_var - 1][j]];
  }
  solve_partial_piv3<7>(A, P);  // P = -A.lu().solve(P)
  for (int k = 0; k < 3; ++k)
    for (int j = 0; j < 7; ++j) P[k][j] = -P[k][j];

  double a[33], c[9];
  re3q3_resultant(P, a, c);  // a11 ... a313, c(0) ... c(8)

  double re[8], im[8];
  poly8_roots(c, re, im);

  int root_cnt = 0;
  for (int i = 0; i < 8; ++i) {
    if (fabs(im[i]) > 1e-8) continue;
    const double xs1 = re[i];
    solutions[0][root_cnt] = xs1;
    re3q3_backsubstitute(a, xs1, &solutions[1][root_cnt], &solutions[2][root_cnt]);
    ++root_cnt;
  }
  if (elim_var == 2) {
    for (int s = 0; s < root_cnt; ++s) swapd(solutions[0][s], solutions[1][s]);
  } else if (elim_var == 3) {
    for (int s = 0; s < root_cnt; ++s) swapd(solutions[0][s], solutions[2][s]);
  }
  return root_cnt;
}

// Picks the elimination variable (re3q3.h:19-37); returns max |det| through *det_out.
__device__ int re3q3_pick(const double coeffs[3][10], double* det_out) {
  double Ax[3][3], Ay[3][3], Az[3][3];
  for (int k = 0; k < 3; ++k) {
    Ax[k][0] = coeffs[k][3]; Ax[k][1] = coeffs[k][5]; Ax[k][2] = coeffs[k][4];
    Ay[k][0] = coeffs[k][0]; Ay[k][1] = coeffs[k][5]; Ay[k][2] = coeffs[k][2];
    Az[k][0] = coeffs[k][3]; Az[k][1] = coeffs[k][0]; Az[k][2] = coeffs[k][1];
  }
  const double detx = fabs(det3(Ax));
  const double dety = fabs(det3(Ay));
  const double detz = fabs(det3(Az));
  int elim_var = 1;
  double det = detx;
  if (det < dety) { det = dety; elim_var = 2; }
  if (det < detz) { det = detz; elim_var = 3; }
  *det_out = det;
  return elim_var;
}

__device__ int re3q3(double coeffs[3][10], double solutions[3][8]) {
  double det;
  int elim_var = re3q3_pick(coeffs, &det);
  if (det < 1e-10) {
    // affine change of variables v = A v' + a, Q' = G^T Q G  (re3q3.h:39-64)
    double G[4][4];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) G[i][j] = kVarChangeA[i][j];
    G[3][0] = 0.0; G[3][1] = 0.0; G[3][2] = 0.0; G[3][3] = 1.0;
    double c2[3][10];
    for (int k = 0; k < 3; ++k) {
      const double* c = coeffs[k];
      double Q[4][4];
      Q[0][0] = c[0];       Q[0][1] = 0.5 * c[1]; Q[0][2] = 0.5 * c[2]; Q[0][3] = 0.5 * c[6];
      Q[1][0] = 0.5 * c[1]; Q[1][1] = c[3];       Q[1][2] = 0.5 * c[4]; Q[1][3] = 0.5 * c[7];
      Q[2][0] = 0.5 * c[2]; Q[2][1] = 0.5 * c[4]; Q[2][2] = c[5];       Q[2][3] = 0.5 * c[8];
      Q[3][0] = 0.5 * c[6]; Q[3][1] = 0.5 * c[7]; Q[3][2] = 0.5 * c[8]; Q[3][3] = c[9];
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
    elim_var = re3q3_pick(c2, &det);
    const int n = re3q3_core(c2, elim_var, solutions);
    for (int s = 0; s < n; ++s) {
      const double x = solutions[0][s], y = solutions[1][s], z = solutions[2][s];
      for (int i = 0; i < 3; ++i)
        solutions[i][s] = kVarChangeA[i][0] * x + kVarChangeA[i][1] * y + kVarChangeA[i][2] * z +
                          kVarChangeA[i][3];
    }
    return n;
  }
  return re3q3_core(coeffs, elim_var, solutions);
}

// P6LEstimator::Estimate.  lines/points: 6 x 3; models: up to 8 x 12 (col-major 3x4).
__device__ int p6l_estimate(const double lines[6][3], const bool all_aligned,
                            const double points[6][3], double models[8][12]) {
  if (all_aligned) return 0;
  double tt[3][9], Rc[3][9];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k)
      for (int j = 0; j < 3; ++j) {
        tt[i][3 * k + j] = points[i][k] * lines[i][j];
        Rc[i][3 * k + j] = points[i + 3][k] * lines[i + 3][j];
      }
  double B[3][3], L1[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      B[r][c] = lines[c][r];
      L1[r][c] = lines[c + 3][r];
    }
  const double det_tt = fabs(det3(B));
  if (det_tt < 1e-10) {
    double tt2[3][9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 9; ++j) {
        const double s = kMixA[i][0] * Rc[0][j] + kMixA[i][1] * Rc[1][j] + kMixA[i][2] * Rc[2][j];
        tt2[i][j] = tt[i][j] + s;
      }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 9; ++j) tt[i][j] = tt2[i][j];
    double B2[3][3];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        const double s = L1[r][0] * kMixA[c][0] + L1[r][1] * kMixA[c][1] + L1[r][2] * kMixA[c][2];
        B2[r][c] = B[r][c] + s;
      }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) B[r][c] = B2[r][c];
  }
  double Bt[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Bt[r][c] = B[c][r];
  solve_partial_piv3<9>(Bt, tt);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 9; ++j) {
      const double s = L1[0][i] * tt[0][j] + L1[1][i] * tt[1][j] + L1[2][i] * tt[2][j];
      Rc[i][j] = Rc[i][j] - s;
    }
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
  const int n_sols = re3q3(coeffs, sols);
  for (int s = 0; s < n_sols; ++s) {
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
    for (int i = 0; i < 3; ++i) {
      double acc = (-tt[i][0]) * m[0];
      for (int j = 1; j < 9; ++j) acc = acc + (-tt[i][j]) * m[j];
      m[9 + i] = acc;
    }
  }
  return n_sols;
}

}  // namespace dev
}  // namespace ppsfm
