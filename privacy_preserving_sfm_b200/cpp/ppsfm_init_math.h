// ppsfm_init_math.h — the per-track arithmetic of the four-view initialisation, written once for
// the host estimators (ppsfm_init.h) and for the GPU scoring kernels (csrc/init_kernels.cu): plain
// arrays, no allocation, `PPSFM_HD` = __host__ __device__ under nvcc.  Both sides execute the same
// IEEE operations in the same order (the CUDA file is built with --fmad=false, the host build has
// no FMA contraction), so a model scored on the GPU gets bit for bit the host's score.
//
//   three-view triangulation of a 2-D point     src/init/sfm2d.cc:196-215
//   FourView2dEstimator::EvaluateModelOnPoint   src/init/sfm2d.cc:302-319
//   four-view triangulation of a 3-D point      src/init/initializer.cc (FourViewTriangulate)
//   PlanarOffsetEstimator::EvaluateModelOnPoint src/init/initializer.cc:310-333
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define PPSFM_HD __host__ __device__
#else
#define PPSFM_HD
#endif

namespace ppsfm {
namespace init {
namespace hd {

// min |A x - b| for a fixed M x N system (row-major, N <= 4) by Householder QR with column
// pivoting; directions whose pivot is negligible get a zero coefficient.  Operation for operation
// la::qr_solve of ppsfm_init.h (tests/test_init.py compares the two on random systems).
// A and b are destroyed.
template <int M, int N>
PPSFM_HD inline void qr_solve_fixed(double* A, double* b, double* x) {
  int perm[4] = {0, 1, 2, 3};
  int rank = 0;
  double first_pivot = 0.0;
  const int steps = M < N ? M : N;
  for (int k = 0; k < steps; ++k) {
    int best = k;
    double best_norm = -1.0;
    for (int j = k; j < N; ++j) {
      double s = 0;
      for (int r = k; r < M; ++r) s += A[r * N + j] * A[r * N + j];
      if (s > best_norm) {
        best_norm = s;
        best = j;
      }
    }
    if (best != k) {
      for (int r = 0; r < M; ++r) {
        const double t = A[r * N + k];
        A[r * N + k] = A[r * N + best];
        A[r * N + best] = t;
      }
      const int t = perm[k];
      perm[k] = perm[best];
      perm[best] = t;
    }
    const double nrm = sqrt(best_norm);
    if (k == 0) first_pivot = nrm;
    if (nrm <= 1e-14 * first_pivot || nrm == 0.0) break;
    ++rank;
    const double akk = A[k * N + k];
    const double alpha = akk >= 0 ? -nrm : nrm;
    double v[M];
    v[0] = akk - alpha;
    for (int r = k + 1; r < M; ++r) v[r - k] = A[r * N + k];
    double vtv = 0;
    for (int r = 0; r < M - k; ++r) vtv += v[r] * v[r];
    if (vtv > 0) {
      for (int j = k; j < N; ++j) {
        double d = 0;
        for (int r = k; r < M; ++r) d += v[r - k] * A[r * N + j];
        d = 2.0 * d / vtv;
        for (int r = k; r < M; ++r) A[r * N + j] -= d * v[r - k];
      }
      double d = 0;
      for (int r = k; r < M; ++r) d += v[r - k] * b[r];
      d = 2.0 * d / vtv;
      for (int r = k; r < M; ++r) b[r] -= d * v[r - k];
    }
  }
  double y[4] = {0, 0, 0, 0};
  for (int k = rank - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < rank; ++j) s -= A[k * N + j] * y[j];
    y[k] = s / A[k * N + k];
  }
  for (int j = 0; j < N; ++j) x[j] = 0.0;
  for (int k = 0; k < rank; ++k) x[perm[k]] = y[k];
}

// 2-D point seen by three 1-D cameras.  P: three 2x3 cameras, row-major [m00 m01 m02 m10 m11 m12];
// x: the three unit observations.  sfm2d.cc:196-215.
PPSFM_HD inline void triangulate2d_point(const double* P0, const double* P1, const double* P2,
                                         const double* x0, const double* x1, const double* x2,
                                         double* X) {
  const double* P[3] = {P0, P1, P2};
  const double* x[3] = {x0, x1, x2};
  double A[6], b[3];
  for (int v = 0; v < 3; ++v) {
    A[2 * v] = x[v][0] * P[v][3] - x[v][1] * P[v][0];
    A[2 * v + 1] = x[v][0] * P[v][4] - x[v][1] * P[v][1];
    b[v] = x[v][1] * P[v][2] - x[v][0] * P[v][5];
  }
  qr_solve_fixed<3, 2>(A, b, X);
}

// FourView2dEstimator::EvaluateModelOnPoint (sfm2d.cc:302-319): max over the four views of the
// 1-D reprojection error; 1e6 for a point behind one of the cameras.
// cams: four 2x3 cameras (24 doubles); x: the four unit observations of the point.
PPSFM_HD inline double fourview2d_error(const double* cams, const double* x0, const double* x1,
                                        const double* x2, const double* x3, const double* X) {
  const double* x[4] = {x0, x1, x2, x3};
  double z[4][2];
  for (int v = 0; v < 4; ++v) {
    const double* P = cams + 6 * v;
    z[v][0] = P[0] * X[0] + P[1] * X[1] + P[2];
    z[v][1] = P[3] * X[0] + P[4] * X[1] + P[5];
  }
  if (z[0][1] < 0 || z[1][1] < 0 || z[2][1] < 0 || z[3][1] < 0) return 1000000.0;
  double err = 0;
  for (int v = 0; v < 4; ++v) {
    const double e = fabs(x[v][0] / x[v][1] - z[v][0] / z[v][1]);
    err = (err < e) ? e : err;  // std::max(err, e)
  }
  return err;
}

// 3-D point on four lifted lines.  cams: four 3x4 cameras (48 doubles, row-major); l: the four
// lines (a, b, c).  initializer.cc FourViewTriangulate.
PPSFM_HD inline void triangulate3d_point(const double* cams, const double* l0, const double* l1,
                                         const double* l2, const double* l3, double* X) {
  const double* l[4] = {l0, l1, l2, l3};
  double A[12], b[4];
  for (int j = 0; j < 4; ++j) {
    const double* P = cams + 12 * j;
    for (int c = 0; c < 3; ++c) A[3 * j + c] = l[j][0] * P[c] + l[j][1] * P[4 + c] + l[j][2] * P[8 + c];
    b[j] = -(l[j][0] * P[3] + l[j][1] * P[7] + l[j][2] * P[11]);
  }
  qr_solve_fixed<4, 3>(A, b, X);
}

// PlanarOffsetEstimator::EvaluateModelOnPoint (initializer.cc:310-333): max over the views of the
// point-to-line distance in the normalised image plane; 1e5 for a point behind a camera.
PPSFM_HD inline double planar_offset_error(const double* cams, const double* l0, const double* l1,
                                           const double* l2, const double* l3, const double* X) {
  const double* l[4] = {l0, l1, l2, l3};
  double z[4][3];
  for (int v = 0; v < 4; ++v) {
    const double* P = cams + 12 * v;
    for (int r = 0; r < 3; ++r)
      z[v][r] = P[4 * r] * X[0] + P[4 * r + 1] * X[1] + P[4 * r + 2] * X[2] + P[4 * r + 3];
  }
  if (z[0][2] < 0 || z[1][2] < 0 || z[2][2] < 0 || z[3][2] < 0) return 100000.0;
  double err = 0;
  for (int v = 0; v < 4; ++v) {
    const double d = (l[v][0] * z[v][0] / z[v][2] + l[v][1] * z[v][1] / z[v][2] + l[v][2]) /
                     sqrt(l[v][0] * l[v][0] + l[v][1] * l[v][1]);
    const double e = fabs(d);
    err = (err < e) ? e : err;
  }
  return err;
}

}  // namespace hd

// Scores many candidate models against all tracks at once (the GPU implementation lives in
// csrc/init_kernels.cu).  A score is the MSAC sum  sum_i min(error_i, threshold)  accumulated in
// track order, i.e. exactly what LocallyOptimizedMSAC::Score computes on the host;
// threshold_first selects std::min(threshold, error) over std::min(error, threshold) (the two
// call sites of the reference differ; the results only differ for a NaN error).
class BatchScorer {
 public:
  virtual ~BatchScorer() {}
  // FourView2dEstimator: cams = num_models x 24 doubles (four 2x3 cameras each); every track is
  // triangulated from the first three views and evaluated in all four.
  virtual bool ScoreFourView2d(const double* cams, int num_models, double threshold,
                               bool threshold_first, double* scores) const = 0;
  // PlanarOffsetEstimator: cams = num_models x 48 doubles (four 3x4 cameras each).
  virtual bool ScorePlanarOffset(const double* cams, int num_models, double threshold,
                                 bool threshold_first, double* scores) const = 0;
};

// Creates the scorers of one initialisation run (observations are handed over once).
class BatchScorerFactory {
 public:
  virtual ~BatchScorerFactory() {}
  virtual const BatchScorer* FourView2d(const double* const* x /* 4 x (n x 2) */, int n) = 0;
  virtual const BatchScorer* PlanarOffset(const double* const* lines /* 4 x (n x 3) */, int n) = 0;
};

}  // namespace init
}  // namespace ppsfm
