// p6l_octet.cuh — the P6L minimal solver with EIGHT lanes per hypothesis.
//
// Same arithmetic as p6l_device.cuh (P6LEstimator::Estimate, src/estimators/absolute_pose.cc:79-162
// + re3q3, lib/re3q3/re3q3/re3q3.h:16-200), operation for operation, hence bit-identical models;
// what changes is who executes it.  The thread-per-hypothesis kernel is one dependent instruction
// chain of ~59 k instructions per hypothesis, most of it the Francis QR sweeps over the 8 x 8
// companion matrix held in a local-memory frame; at the head of a RANSAC call (first wave, or a
// mapper-sized call) the GPU is nearly empty and that chain IS the latency of the call.  Here an
// "octet" of 8 lanes owns a hypothesis:
//   * the matrix lives in shared memory (8 x 9 doubles per octet: conflict-free by row and by
//     column), no local-memory frame;
//   * the scalar control of the QR iteration (deflation tests, shifts, Householder vectors) is
//     evaluated redundantly by the 8 lanes — free in SIMT, and it keeps the octet converged;
//   * a reflection from the left updates one COLUMN per lane, from the right one ROW per lane:
//     every element is produced by the same three-term expression as in the serial code;
//   * after the iteration every lane owns one eigenvalue: back-substitution, Cayley transform and
//     translation of the up-to-8 solutions run side by side, and a ballot compacts the real roots
//     in index order.
// Octets of a warp follow different control paths (iteration counts differ), so every exchange
// uses __syncwarp / __ballot_sync with the octet's own 8-lane mask.
#pragma once
#include "p6l_device.cuh"

namespace ppsfm {
namespace dev {

struct Octet {
  double* T;      // shared memory, 8 x kLd doubles
  int sub;        // lane within the octet, 0..7
  unsigned mask;  // the octet's lanes within the warp
  static constexpr int kLd = 9;
  __device__ __forceinline__ double& at(int i, int j) const { return T[i * kLd + j]; }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

// Hqr8::reduce of p6l_device.cuh, cooperatively.  Returns false if the iteration limit is hit.
// (The member mask is a run-time value: a variant templated on the octet's position in its warp
// — constant masks, single-instruction __syncwarp — was measured TWICE as slow, 0.25 against
// 0.12 ms: the four octets of a warp then run four copies of the code and can never issue
// together, whereas in this shared-code form they are converged most of the time.)
__device__ inline bool hqr8_reduce_octet(const Octet& o) {
  const int sub = o.sub;
  const unsigned oct_shift = (threadIdx.x & 31u) & ~7u;  // first lane of the octet in its warp
  // scale = max |T| (the comparator of the serial code ignores NaNs: any order gives the same)
  double scale = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double v = fabs(o.at(sub, j));
    scale = (scale < v) ? v : scale;
  }
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) {
    const double other = __shfl_xor_sync(o.mask, scale, s);
    scale = (scale < other) ? other : scale;
  }
  if (!(scale > 0.0) || !isfinite(scale)) return isfinite(scale);
#pragma unroll
  for (int j = 0; j < 8; ++j) o.at(sub, j) = o.at(sub, j) / scale;
  o.sync();
  // norm != 0 of the serial code: a sum of magnitudes over the Hessenberg part is non-zero (or
  // NaN) iff one of its terms is
  bool nz = false;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (sub < ((j + 2 < 8) ? (j + 2) : 8)) nz = nz || (fabs(o.at(sub, j)) != 0.0);
  const bool norm_nonzero = (__ballot_sync(o.mask, nz) & o.mask) != 0u;

  const int max_iters = 40 * 8;
  int iu = 7, iter = 0, total_iter = 0;
  double exshift = 0.0;
  const double eps = DBL_EPSILON;
  bool ok = true;
  if (norm_nonzero) {
    while (iu >= 0) {
      // findSmallSubdiagEntry: the serial loop stops at the largest i <= iu whose sub-diagonal
      // entry is negligible; lane i tests entry i, a ballot picks the largest
      int il;
      {
        bool small = false;
        if (sub >= 1 && sub <= iu) {
          const double s = fabs(o.at(sub - 1, sub - 1)) + fabs(o.at(sub, sub));
          small = fabs(o.at(sub, sub - 1)) <= eps * s;
        }
        const unsigned bits = (__ballot_sync(o.mask, small) & o.mask) >> oct_shift;
        il = bits ? 31 - __clz((int)bits) : 0;
        o.sync();  // the search's reads are ordered before the writes of the branch taken below
      }
      if (il == iu) {
        if (sub == 0) {
          o.at(iu, iu) = o.at(iu, iu) + exshift;
          if (iu > 0) o.at(iu, iu - 1) = 0.0;
        }
        o.sync();
        --iu;
        iter = 0;
      } else if (il == iu - 1) {
        // 2 x 2 block in registers, written back by one lane
        double a00 = o.at(iu - 1, iu - 1), a01 = o.at(iu - 1, iu);
        double a10 = o.at(iu, iu - 1), a11 = o.at(iu, iu);
        const double p = 0.5 * (a00 - a11);
        const double q = p * p + a10 * a01;
        a11 = a11 + exshift;
        a00 = a00 + exshift;
        if (q >= 0.0) {
          const double z = sqrt(fabs(q));
          const double a = (p >= 0.0) ? (p + z) : (p - z);
          const double b = a10;
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
          // rows (iu-1, iu) <- G^T rows, columns iu-1..iu
          const double r00 = c * a00 - s * a10, r10 = s * a00 + c * a10;
          const double r01 = c * a01 - s * a11, r11 = s * a01 + c * a11;
          // columns (iu-1, iu) <- columns G, rows iu-1..iu
          a00 = c * r00 - s * r01;
          a01 = s * r00 + c * r01;
          a10 = c * r10 - s * r11;
          a11 = s * r10 + c * r11;
          a10 = 0.0;
        }
        o.sync();  // every lane has read the block
        if (sub == 0) {
          o.at(iu - 1, iu - 1) = a00;
          o.at(iu - 1, iu) = a01;
          o.at(iu, iu - 1) = a10;
          o.at(iu, iu) = a11;
          if (iu > 1) o.at(iu - 1, iu - 2) = 0.0;
        }
        o.sync();
        iu -= 2;
        iter = 0;
      } else {
        double sh0 = o.at(iu, iu);
        double sh1 = o.at(iu - 1, iu - 1);
        double sh2 = o.at(iu, iu - 1) * o.at(iu - 1, iu);
        if (iter == 10) {
          exshift = exshift + sh0;
          o.sync();  // every lane has read its shifts
          if (sub <= iu) o.at(sub, sub) = o.at(sub, sub) - sh0;
          o.sync();
          const double s = fabs(o.at(iu, iu - 1)) + fabs(o.at(iu - 1, iu - 2));
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
            o.sync();
            if (sub <= iu) o.at(sub, sub) = o.at(sub, sub) - s;
            o.sync();
            sh0 = sh1 = sh2 = 0.964;
          }
        }
        ++iter;
        ++total_iter;
        if (total_iter > max_iters) {
          ok = false;
          break;
        }
        // initFrancisQRStep: the serial search walks im = iu-2 ... il and stops at the first im
        // that passes the test (im == il always does); every candidate only reads the matrix, so
        // lane m evaluates candidate m and the largest passing one wins
        int im;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        {
          bool pass = false;
          if (sub >= il && sub <= iu - 2) {
            const double Tmm = o.at(sub, sub);
            const double r = sh0 - Tmm;
            const double s = sh1 - Tmm;
            v0 = (r * s - sh2) / o.at(sub + 1, sub) + o.at(sub, sub + 1);
            v1 = o.at(sub + 1, sub + 1) - Tmm - r - s;
            v2 = o.at(sub + 2, sub + 1);
            if (sub == il) {
              pass = true;
            } else {
              const double lhs = o.at(sub, sub - 1) * (fabs(v1) + fabs(v2));
              const double rhs =
                  v0 * (fabs(o.at(sub - 1, sub - 1)) + fabs(Tmm) + fabs(o.at(sub + 1, sub + 1)));
              pass = fabs(lhs) < eps * rhs;
            }
          }
          const unsigned bits = (__ballot_sync(o.mask, pass) & o.mask) >> oct_shift;
          im = 31 - __clz((int)bits);  // bit il is always set
          const int src = (int)oct_shift + im;
          v0 = __shfl_sync(o.mask, v0, src);
          v1 = __shfl_sync(o.mask, v1, src);
          v2 = __shfl_sync(o.mask, v2, src);
        }
        o.sync();  // all reads of the search are done before the sweep writes
        for (int k = im; k <= iu - 2; ++k) {
          const bool first = (k == im);
          double w0, w1, w2;
          if (first) {
            w0 = v0; w1 = v1; w2 = v2;
          } else {
            w0 = o.at(k, k - 1);
            w1 = o.at(k + 1, k - 1);
            w2 = o.at(k + 2, k - 1);
          }
          // makeHouseholder on (w0, w1, w2)
          double tail_sq = w1 * w1;
          tail_sq = tail_sq + w2 * w2;
          double tau, beta, e0, e1;
          if (tail_sq <= DBL_MIN) {
            tau = 0.0; beta = w0; e0 = 0.0; e1 = 0.0;
          } else {
            double b = sqrt(w0 * w0 + tail_sq);
            if (w0 >= 0.0) b = -b;
            e0 = w1 / (w0 - b);
            e1 = w2 / (w0 - b);
            tau = (b - w0) / b;
            beta = b;
          }
          if (beta != 0.0) {
            const double sub_diag = o.at(k, k - (k > 0 ? 1 : 0));
            o.sync();  // w and sub_diag are read; column k-1 may now be overwritten
            if (sub == 0) {
              if (first && k > il) o.at(k, k - 1) = -sub_diag;
              else if (!first) o.at(k, k - 1) = beta;
            }
            if (tau != 0.0) {
              // from the left: rows k..k+2, one column (k..iu) per lane
              if (sub >= k && sub <= iu) {
                const double t0 = o.at(k, sub), t1 = o.at(k + 1, sub), t2 = o.at(k + 2, sub);
                double tmp = e0 * t1;
                tmp = tmp + e1 * t2;
                tmp = tmp + t0;
                o.at(k, sub) = t0 - tau * tmp;
                o.at(k + 1, sub) = t1 - (tau * e0) * tmp;
                o.at(k + 2, sub) = t2 - (tau * e1) * tmp;
              }
              o.sync();
              // from the right: columns k..k+2, one row (il..min(iu, k+3)) per lane
              const int r_hi = (iu < k + 3) ? iu : (k + 3);
              if (sub >= il && sub <= r_hi) {
                const double t0 = o.at(sub, k), t1 = o.at(sub, k + 1), t2 = o.at(sub, k + 2);
                double tmp = t1 * e0;
                tmp = tmp + t2 * e1;
                tmp = tmp + t0;
                o.at(sub, k) = t0 - tau * tmp;
                o.at(sub, k + 1) = t1 - (tau * tmp) * e0;
                o.at(sub, k + 2) = t2 - (tau * tmp) * e1;
              }
            }
            o.sync();
          }
        }
        {
          const double w0 = o.at(iu - 1, iu - 2), w1 = o.at(iu, iu - 2);
          const double tail_sq = w1 * w1;
          double tau, beta, e0;
          if (tail_sq <= DBL_MIN) {
            tau = 0.0; beta = w0; e0 = 0.0;
          } else {
            double b = sqrt(w0 * w0 + tail_sq);
            if (w0 >= 0.0) b = -b;
            e0 = w1 / (w0 - b);
            tau = (b - w0) / b;
            beta = b;
          }
          if (beta != 0.0) {
            o.sync();
            if (sub == 0) o.at(iu - 1, iu - 2) = beta;
            if (tau != 0.0) {
              if (sub >= iu - 1 && sub <= iu) {
                const double t0 = o.at(iu - 1, sub), t1 = o.at(iu, sub);
                double tmp = e0 * t1;
                tmp = tmp + t0;
                o.at(iu - 1, sub) = t0 - tau * tmp;
                o.at(iu, sub) = t1 - (tau * e0) * tmp;
              }
              o.sync();
              if (sub >= il && sub <= iu) {
                const double t0 = o.at(sub, iu - 1), t1 = o.at(sub, iu);
                double tmp = t1 * e0;
                tmp = tmp + t0;
                o.at(sub, iu - 1) = t0 - tau * tmp;
                o.at(sub, iu) = t1 - (tau * tmp) * e0;
              }
            }
            o.sync();
          }
        }
        o.sync();  // (a skipped reflection leaves its reads unordered against the zeroing below)
        if (sub >= im + 2 && sub <= iu) {
          o.at(sub, sub - 2) = 0.0;
          if (sub > im + 2) o.at(sub, sub - 3) = 0.0;
        }
        o.sync();
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) o.at(sub, j) = o.at(sub, j) * scale;
  o.sync();
  return ok;
}

// Hqr8::eigenvalues for the lane's own index: (re, im) of eigenvalue `sub`.
__device__ inline void hqr8_eigenvalue_octet(const Octet& o, double* re, double* im) {
  // the serial walk from the top decides which rows head a 2 x 2 block
  int i = 0, top = -1;  // top: head of the block that contains row `sub` (or sub itself)
  while (i < 8) {
    if (i == 7 || o.at(i + 1, i) == 0.0) {
      if (i == o.sub) top = -1;
      ++i;
    } else {
      if (i == o.sub || i + 1 == o.sub) top = i;
      i += 2;
    }
  }
  if (top < 0) {
    *re = o.at(o.sub, o.sub);
    *im = 0.0;
    return;
  }
  const double p = 0.5 * (o.at(top, top) - o.at(top + 1, top + 1));
  double t0 = o.at(top + 1, top);
  double t1 = o.at(top, top + 1);
  const double at0 = fabs(t0), at1 = fabs(t1), ap = fabs(p);
  const double m01 = (at0 < at1) ? at1 : at0;
  const double maxval = (ap < m01) ? m01 : ap;
  t0 = t0 / maxval;
  t1 = t1 / maxval;
  const double p0 = p / maxval;
  const double z = maxval * sqrt(fabs(p0 * p0 + t0 * t1));
  *re = o.at(top + 1, top + 1) + p;
  *im = (o.sub == top) ? z : -z;
}

// re3q3_core with the eigenvalue step on the octet; every lane returns with ITS solution
// (x, y, z) and *keep (its root is real); the return value is the number of real roots and
// *pos the lane's index among them (roots keep the order of the Schur diagonal).
__device__ inline int re3q3_core_octet(const Octet& o, const double coeffs[3][10], int elim_var,
                                       double sol[3], bool* keep, int* pos) {
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
  solve_partial_piv3<7>(A, P);
  for (int k = 0; k < 3; ++k)
    for (int j = 0; j < 7; ++j) P[k][j] = -P[k][j];
  double a[33], c[9];
  re3q3_resultant(P, a, c);

  // companion matrix (re3q3.h:152-160): lane `sub` fills row `sub`
  o.sync();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double v = 0.0;
    if (o.sub == 0) v = -c[j + 1] / c[0];
    else if (j == o.sub - 1) v = 1.0;
    o.at(o.sub, j) = v;
  }
  o.sync();
  hqr8_reduce_octet(o);
  double re, im;
  hqr8_eigenvalue_octet(o, &re, &im);
  const bool real_root = !(fabs(im) > 1e-8);
  const unsigned bal = __ballot_sync(o.mask, real_root) & o.mask;
  const unsigned lane = threadIdx.x & 31;
  *keep = real_root;
  *pos = __popc(bal & ((1u << lane) - 1u));
  sol[0] = re;
  re3q3_backsubstitute(a, re, &sol[1], &sol[2]);
  if (elim_var == 2) swapd(sol[0], sol[1]);
  else if (elim_var == 3) swapd(sol[0], sol[2]);
  return __popc(bal);
}

__device__ inline int re3q3_octet(const Octet& o, double coeffs[3][10], double sol[3], bool* keep,
                                  int* pos) {
  double det;
  int elim_var = re3q3_pick(coeffs, &det);
  if (det < 1e-10) {
    // affine change of variables v = A v' + a, Q' = G^T Q G  (re3q3.h:39-64), as in re3q3()
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
    const int n = re3q3_core_octet(o, c2, elim_var, sol, keep, pos);
    const double x = sol[0], y = sol[1], z = sol[2];
    for (int i = 0; i < 3; ++i)
      sol[i] = kVarChangeA[i][0] * x + kVarChangeA[i][1] * y + kVarChangeA[i][2] * z +
               kVarChangeA[i][3];
    return n;
  }
  return re3q3_core_octet(o, coeffs, elim_var, sol, keep, pos);
}

// P6LEstimator::Estimate on an octet: every lane evaluates the set-up redundantly; lanes whose
// root is real write model `*pos` to models_out (8 x 12).  Returns the number of models.
__device__ inline int p6l_estimate_octet(const Octet& o, const double lines[6][3],
                                         const bool all_aligned, const double points[6][3],
                                         double* __restrict__ models_out) {
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
  double sol[3];
  bool keep;
  int pos;
  const int n_sols = re3q3_octet(o, coeffs, sol, &keep, &pos);
  if (keep) {
    const double c0 = sol[0], c1 = sol[1], c2 = sol[2];
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
    double m[12];
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < 3; ++r) m[3 * c + r] = R[r][c] / nrm;
    for (int i = 0; i < 3; ++i) {
      double acc = (-tt[i][0]) * m[0];
      for (int j = 1; j < 9; ++j) acc = acc + (-tt[i][j]) * m[j];
      m[9 + i] = acc;
    }
    double* out = models_out + 12 * pos;
#pragma unroll
    for (int j = 0; j < 12; ++j) out[j] = m[j];
  }
  return n_sols;
}

}  // namespace dev
}  // namespace ppsfm
