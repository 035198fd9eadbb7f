// ba_oracle.cc — CPU oracle of the line-reprojection bundle adjustment (see ba_oracle.h).
// TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED at the Ceres boundary (Ceres is not in-tree).
//
// Follows, by file:line of colmap/privacy_preserving_sfm:
//   residual functors            src/base/cost_functions.h:46-191
//   camera models                src/base/camera_models.h:615-904
//   problem assembly, gauge      src/optim/bundle_adjustment.cc:326-542
//   loss selection               src/optim/bundle_adjustment.cc:55-70
//   pose refinement              src/estimators/pose.cc:96-213
// and Ceres' public algorithm description for the trust-region LM.

#include "ba_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#include "camera_models_ext.h"

namespace {

// ---------------------------------------------------------------------------------------------
// Forward-mode dual numbers (what ceres::Jet<double, N> provides to AutoDiffCostFunction).
// ---------------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  explicit Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double x, int k) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> Jet<N> operator-(const Jet<N>& x) {
  Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
template <int N> Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) {
  Jet<N> r; const double inv = 1.0 / y.a; r.a = x.a * inv;
  for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
template <int N> Jet<N>& operator+=(Jet<N>& x, const Jet<N>& y) { x = x + y; return x; }
template <int N> Jet<N>& operator/=(Jet<N>& x, const Jet<N>& y) { x = x / y; return x; }

template <int N> Jet<N> sqrt(const Jet<N>& x) {
  Jet<N> r; r.a = std::sqrt(x.a); const double d = 1.0 / (2.0 * r.a);
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> Jet<N> atan(const Jet<N>& x) {
  Jet<N> r; r.a = std::atan(x.a); const double d = 1.0 / (1.0 + x.a * x.a);
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> Jet<N> tan(const Jet<N>& x) {
  Jet<N> r; r.a = std::tan(x.a); const double d = 1.0 + r.a * r.a;
  for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <typename T> struct Lift { static T C(double x) { return T(x); } };
template <> struct Lift<double> { static double C(double x) { return x; } };

// ceres::UnitQuaternionRotatePoint (ceres/rotation.h), q = (w, x, y, z), assumes |q| = 1.
template <typename T>
void UnitQuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  const T two = Lift<T>::C(2.0);
  const T t2 = q[0] * q[1];
  const T t3 = q[0] * q[2];
  const T t4 = q[0] * q[3];
  const T t5 = -(q[1] * q[1]);
  const T t6 = q[1] * q[2];
  const T t7 = q[1] * q[3];
  const T t8 = -(q[2] * q[2]);
  const T t9 = q[2] * q[3];
  const T t1 = -(q[3] * q[3]);
  result[0] = two * ((t8 + t1) * pt[0] + (t6 - t4) * pt[1] + (t3 + t7) * pt[2]) + pt[0];
  result[1] = two * ((t4 + t6) * pt[0] + (t5 + t1) * pt[1] + (t9 - t2) * pt[2]) + pt[1];
  result[2] = two * ((t7 - t3) * pt[0] + (t2 + t9) * pt[1] + (t5 + t8) * pt[2]) + pt[2];
}

// CameraModel::WorldToImage<T> (src/base/camera_models.h).  The intrinsics `p` are either plain
// doubles (constant camera: refine_* = false, the defaults of src/optim/bundle_adjustment.h:57-63)
// or jets like everything else (camera_params is the fourth parameter block of the functor,
// cost_functions.h:56-58).
template <typename T> inline T AsT(const T& x) { return x; }
template <typename T, typename = typename std::enable_if<!std::is_same<T, double>::value>::type>
inline T AsT(double x) { return Lift<T>::C(x); }
template <typename T, typename PT>
bool WorldToImage(int model, const PT* p, const T u, const T v, T* x, T* y) {
  auto P = [&](int k) { return AsT<T>(p[k]); };
  auto C = [](double c) { return Lift<T>::C(c); };
  switch (model) {
    case 0: {  // SIMPLE_PINHOLE f, cx, cy                      (:615-627)
      *x = P(0) * u + P(1);
      *y = P(0) * v + P(2);
      return true;
    }
    case 1: {  // PINHOLE fx, fy, cx, cy                        (:664-676)
      *x = P(0) * u + P(2);
      *y = P(1) * v + P(3);
      return true;
    }
    case 2: {  // SIMPLE_RADIAL f, cx, cy, k                    (:715-757)
      const T u2 = u * u, v2 = v * v, r2 = u2 + v2;
      const T radial = P(3) * r2;
      const T xx = u + u * radial, yy = v + v * radial;
      *x = P(0) * xx + P(1);
      *y = P(0) * yy + P(2);
      return true;
    }
    case 3: {  // RADIAL f, cx, cy, k1, k2                      (:784-829)
      const T u2 = u * u, v2 = v * v, r2 = u2 + v2;
      const T radial = P(3) * r2 + P(4) * r2 * r2;
      const T xx = u + u * radial, yy = v + v * radial;
      *x = P(0) * xx + P(1);
      *y = P(0) * yy + P(2);
      return true;
    }
    case 4: {  // OPENCV fx, fy, cx, cy, k1, k2, p1, p2          (:854-904)
      const T u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2;
      const T radial = P(4) * r2 + P(5) * r2 * r2;
      const T du = u * radial + C(2.0) * P(6) * uv + P(7) * (r2 + C(2.0) * u2);
      const T dv = v * radial + C(2.0) * P(7) * uv + P(6) * (r2 + C(2.0) * v2);
      const T xx = u + du, yy = v + dv;
      *x = P(0) * xx + P(2);
      *y = P(1) * yy + P(3);
      return true;
    }
  }
  // OPENCV_FISHEYE, FULL_OPENCV, FOV, SIMPLE_RADIAL_FISHEYE, RADIAL_FISHEYE, THIN_PRISM_FISHEYE
  return orc_cam::WorldToImageExt(model, p, u, v, x, y, C, P);
}

// BundleAdjustmentLineCostFunction::operator() (src/base/cost_functions.h:62-100).
template <typename T, typename PT>
bool LineCost(int model, const PT* cam, const double* line, const T* qvec, const T* tvec,
              const T* point3D, T* residuals) {
  T projection[3];
  UnitQuaternionRotatePoint(qvec, point3D, projection);
  projection[0] += tvec[0];
  projection[1] += tvec[1];
  projection[2] += tvec[2];
  projection[0] /= projection[2];
  projection[1] /= projection[2];
  const T a = Lift<T>::C(line[0]), b = Lift<T>::C(line[1]), c = Lift<T>::C(line[2]);
  const T alpha = a * projection[0] + b * projection[1] + c;
  T line_point[2];
  line_point[0] = projection[0] - alpha * a;
  line_point[1] = projection[1] - alpha * b;
  T im_projection[2], im_line_point[2];
  if (!WorldToImage(model, cam, projection[0], projection[1], &im_projection[0],
                    &im_projection[1]))
    return false;
  WorldToImage(model, cam, line_point[0], line_point[1], &im_line_point[0], &im_line_point[1]);
  residuals[0] = im_projection[0] - im_line_point[0];
  residuals[1] = im_projection[1] - im_line_point[1];
  return true;
}

// AutoDiff of the block (2; 4, 3, 3): residual + row-major Jacobians.
void LineCostAutoDiff(int model, const double* cam, const double* line, const double* q,
                      const double* t, const double* X, double* r, double* Jq, double* Jt,
                      double* JX) {
  typedef Jet<10> J;
  J jq[4], jt[3], jX[3], res[2];
  for (int i = 0; i < 4; ++i) jq[i] = J(q[i], i);
  for (int i = 0; i < 3; ++i) jt[i] = J(t[i], 4 + i);
  for (int i = 0; i < 3; ++i) jX[i] = J(X[i], 7 + i);
  LineCost<J>(model, cam, line, jq, jt, jX, res);
  for (int k = 0; k < 2; ++k) {
    r[k] = res[k].a;
    for (int i = 0; i < 4; ++i) Jq[4 * k + i] = res[k].v[i];
    for (int i = 0; i < 3; ++i) Jt[3 * k + i] = res[k].v[4 + i];
    for (int i = 0; i < 3; ++i) JX[3 * k + i] = res[k].v[7 + i];
  }
}

// AutoDiff of the block (2; 4, 3, 3, kNumParams): additionally the 2 x 8 Jacobian with respect to
// camera_params (columns >= NumParams stay zero).
constexpr int kIntrWidth = 12;  // widest model (FULL_OPENCV, THIN_PRISM_FISHEYE)
const int kNumParams[11] = {3, 4, 4, 5, 8, 8, 12, 5, 4, 5, 12};
void LineCostAutoDiffIntr(int model, const double* cam, const double* line, const double* q,
                          const double* t, const double* X, double* r, double* Jq, double* Jt,
                          double* JX, double* Jcam) {
  typedef Jet<10 + kIntrWidth> J;
  J jq[4], jt[3], jX[3], jc[kIntrWidth], res[2];
  for (int i = 0; i < 4; ++i) jq[i] = J(q[i], i);
  for (int i = 0; i < 3; ++i) jt[i] = J(t[i], 4 + i);
  for (int i = 0; i < 3; ++i) jX[i] = J(X[i], 7 + i);
  for (int i = 0; i < kIntrWidth; ++i)
    jc[i] = i < kNumParams[model] ? J(cam[i], 10 + i) : J(0.0);
  LineCost<J, J>(model, jc, line, jq, jt, jX, res);
  for (int k = 0; k < 2; ++k) {
    r[k] = res[k].a;
    for (int i = 0; i < 4; ++i) Jq[4 * k + i] = res[k].v[i];
    for (int i = 0; i < 3; ++i) Jt[3 * k + i] = res[k].v[4 + i];
    for (int i = 0; i < 3; ++i) JX[3 * k + i] = res[k].v[7 + i];
    for (int i = 0; i < kIntrWidth; ++i) Jcam[kIntrWidth * k + i] = res[k].v[10 + i];
  }
}

// ceres::QuaternionParameterization::ComputeJacobian (4x3, row-major)
void QuaternionPlusJacobian(const double* x, double* j) {
  j[0] = -x[1]; j[1] = -x[2]; j[2] = -x[3];
  j[3] = x[0];  j[4] = x[3];  j[5] = -x[2];
  j[6] = -x[3]; j[7] = x[0];  j[8] = x[1];
  j[9] = x[2];  j[10] = -x[1]; j[11] = x[0];
}

// ceres::QuaternionParameterization::Plus
void QuaternionPlus(const double* x, const double* delta, double* out) {
  const double norm_delta =
      std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (norm_delta > 0.0) {
    const double s = std::sin(norm_delta) / norm_delta;
    const double qd[4] = {std::cos(norm_delta), s * delta[0], s * delta[1], s * delta[2]};
    // QuaternionProduct(q_delta, x)
    out[0] = qd[0] * x[0] - qd[1] * x[1] - qd[2] * x[2] - qd[3] * x[3];
    out[1] = qd[0] * x[1] + qd[1] * x[0] + qd[2] * x[3] - qd[3] * x[2];
    out[2] = qd[0] * x[2] - qd[1] * x[3] + qd[2] * x[0] + qd[3] * x[1];
    out[3] = qd[0] * x[3] + qd[1] * x[2] - qd[2] * x[1] + qd[3] * x[0];
  } else {
    for (int i = 0; i < 4; ++i) out[i] = x[i];
  }
}

void LineCostTangent(int model, const double* cam, const double* line, const double* q,
                     const double* t, const double* X, double* r, double* Jc, double* JX,
                     double* Jcam = nullptr) {
  double Jq[8], Jt[6], pj[12];
  if (Jcam) LineCostAutoDiffIntr(model, cam, line, q, t, X, r, Jq, Jt, JX, Jcam);
  else LineCostAutoDiff(model, cam, line, q, t, X, r, Jq, Jt, JX);
  QuaternionPlusJacobian(q, pj);
  for (int k = 0; k < 2; ++k) {
    for (int c = 0; c < 3; ++c) {
      double s = 0.0;
      for (int i = 0; i < 4; ++i) s += Jq[4 * k + i] * pj[3 * i + c];
      Jc[6 * k + c] = s;
      Jc[6 * k + 3 + c] = Jt[3 * k + c];
    }
  }
}

// ceres loss functions: rho[0] = rho(s), rho[1] = rho'(s), rho[2] = rho''(s)
void EvaluateLoss(int type, double a, double s, double rho[3]) {
  if (type == 0) {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    return;
  }
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  if (type == 1) {  // SoftLOneLoss
    const double tmp = std::sqrt(sum);
    rho[0] = 2.0 * b * (tmp - 1.0);
    rho[1] = std::max(std::numeric_limits<double>::min(), 1.0 / tmp);
    rho[2] = -(c * rho[1]) / (2.0 * sum);
  } else {  // CauchyLoss
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
  }
}

// ceres Corrector: scales the residual and corrects the Jacobian rows of one block (2 x ncols).
void ApplyCorrector(double sq_norm, const double rho[3], double* r, double* J, int ncols) {
  const double sqrt_rho1 = std::sqrt(rho[1]);
  double residual_scaling, alpha_sq_norm;
  if (sq_norm == 0.0 || rho[2] <= 0.0) {
    residual_scaling = sqrt_rho1;
    alpha_sq_norm = 0.0;
  } else {
    const double D = 1.0 + 2.0 * sq_norm * rho[2] / rho[1];
    const double alpha = 1.0 - std::sqrt(D);
    residual_scaling = sqrt_rho1 / (1 - alpha);
    alpha_sq_norm = alpha / sq_norm;
  }
  if (alpha_sq_norm == 0.0) {
    for (int i = 0; i < 2 * ncols; ++i) J[i] *= sqrt_rho1;
  } else {
    for (int c = 0; c < ncols; ++c) {
      const double rtj = r[0] * J[c] + r[1] * J[ncols + c];
      J[c] = sqrt_rho1 * (J[c] - alpha_sq_norm * r[0] * rtj);
      J[ncols + c] = sqrt_rho1 * (J[ncols + c] - alpha_sq_norm * r[1] * rtj);
    }
  }
  r[0] *= residual_scaling;
  r[1] *= residual_scaling;
}

template <typename F>
void ParallelFor(int num_threads, int64_t n, F fn) {
  if (num_threads <= 1 || n < 2) {
    fn(0, 0, n);
    return;
  }
  std::vector<std::thread> th;
  const int64_t chunk = (n + num_threads - 1) / num_threads;
  for (int t = 0; t < num_threads; ++t) {
    const int64_t lo = t * chunk, hi = std::min<int64_t>(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([=]() { fn(t, lo, hi); });
  }
  for (auto& x : th) x.join();
}

// In-place lower Cholesky of the dense symmetric n x n matrix (row-major, lower part used).
// Blocked right-looking; the trailing update is threaded.  Returns false if not SPD.
bool CholeskyLower(std::vector<double>& A, int n, int num_threads) {
  const int NB = 64;
  for (int k0 = 0; k0 < n; k0 += NB) {
    const int kb = std::min(NB, n - k0);
    for (int j = k0; j < k0 + kb; ++j) {  // diagonal block
      double d = A[(size_t)j * n + j];
      for (int p = k0; p < j; ++p) d -= A[(size_t)j * n + p] * A[(size_t)j * n + p];
      if (!(d > 0.0)) return false;
      d = std::sqrt(d);
      A[(size_t)j * n + j] = d;
      for (int i = j + 1; i < k0 + kb; ++i) {
        double s = A[(size_t)i * n + j];
        for (int p = k0; p < j; ++p) s -= A[(size_t)i * n + p] * A[(size_t)j * n + p];
        A[(size_t)i * n + j] = s / d;
      }
    }
    const int r0 = k0 + kb;
    if (r0 >= n) break;
    ParallelFor(num_threads, n - r0, [&](int, int64_t lo, int64_t hi) {  // panel solve
      for (int64_t ii = lo; ii < hi; ++ii) {
        double* row = &A[(size_t)(r0 + ii) * n];
        for (int j = k0; j < k0 + kb; ++j) {
          double s = row[j];
          const double* lj = &A[(size_t)j * n];
          for (int p = k0; p < j; ++p) s -= row[p] * lj[p];
          row[j] = s / lj[j];
        }
      }
    });
    ParallelFor(num_threads, n - r0, [&](int, int64_t lo, int64_t hi) {  // trailing update
      for (int64_t ii = lo; ii < hi; ++ii) {
        const int i = r0 + (int)ii;
        double* row = &A[(size_t)i * n];
        for (int j = r0; j <= i; ++j) {
          const double* rj = &A[(size_t)j * n];
          double s = 0.0;
          for (int p = k0; p < k0 + kb; ++p) s += row[p] * rj[p];
          row[j] -= s;
        }
      }
    });
  }
  return true;
}

void CholeskySolve(const std::vector<double>& L, int n, std::vector<double>& b) {
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int p = 0; p < i; ++p) s -= L[(size_t)i * n + p] * b[p];
    b[i] = s / L[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int p = i + 1; p < n; ++p) s -= L[(size_t)p * n + i] * b[p];
    b[i] = s / L[(size_t)i * n + i];
  }
}

bool Invert3(const double* V, double* inv) {  // symmetric 3x3, row-major full
  const double a = V[0], b = V[1], c = V[2], d = V[4], e = V[5], f = V[8];
  const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
  const double det = a * c00 + b * c01 + c * c02;
  if (!(std::fabs(det) > 0.0)) return false;
  const double id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
  inv[3] = inv[1];   inv[4] = (a * f - c * c) * id; inv[5] = (b * c - a * e) * id;
  inv[6] = inv[2];   inv[7] = inv[5]; inv[8] = (a * d - b * b) * id;
  return true;
}

struct Solver {
  const orc_ba_problem& pb;
  const orc_ba_options& opt;
  int C, P;
  int64_t O;
  std::vector<double> q, t, X;            // current state
  std::vector<int> cam_block;             // image -> reduced block index or -1 (constant pose)
  std::vector<uint8_t> cam_mask;          // 6 bits: active tangent dims
  std::vector<uint8_t> pt_var;
  std::vector<int64_t> obs;               // kept observation ids, point-major
  std::vector<int64_t> pt_start;          // P + 1
  int nblocks = 0;
  // per kept observation (scaled, loss-corrected)
  std::vector<double> r, Jc, Jp;          // 2, 12, 6 per obs
  std::vector<double> cam_scale, pt_scale;  // 6 per block, 3 per point
  int threads;
  // intrinsics (ParameterizeCameras, bundle_adjustment.cc:490-528): cameras with at least one
  // variable parameter get a reduced block of kIntrWidth columns after the pose blocks
  std::vector<double> params;             // current Camera::Params(), 12 per camera
  std::vector<int> intr_block;            // camera -> block index or -1 (constant intrinsics)
  std::vector<uint16_t> intr_mask;        // per camera: bit k = parameter k is variable
  int nintr = 0;
  std::vector<double> Ji;                 // 2 x kIntrWidth per obs (only if nintr > 0)
  std::vector<double> intr_scale;         // kIntrWidth per intrinsics block

  Solver(const orc_ba_problem& p, const orc_ba_options& o) : pb(p), opt(o) {
    C = p.num_images; P = p.num_points; O = p.num_obs;
    threads = o.num_threads > 0 ? o.num_threads
                                : (int)std::max(1u, std::thread::hardware_concurrency());
    q.assign(p.qvecs, p.qvecs + 4 * (size_t)C);
    t.assign(p.tvecs, p.tvecs + 3 * (size_t)C);
    X.assign(p.points, p.points + 3 * (size_t)P);
    // image.NormalizeQvec() (bundle_adjustment.cc:355) for images with a variable pose
    cam_block.assign(C, -1);
    cam_mask.assign(C, 0);
    pt_var.assign(P, 1);
    if (p.point_const) for (int i = 0; i < P; ++i) pt_var[i] = p.point_const[i] ? 0 : 1;
    // ParameterizeCameras: constant unless a refine_* flag is set and the camera is not in
    // config.ConstantCameras(); the groups that are not refined stay constant
    // (SubsetParameterization).  Parameter groups per model: camera_models.h:597-846.
    params.assign(p.camera_params, p.camera_params + 12 * (size_t)p.num_cameras);
    intr_block.assign(p.num_cameras, -1);
    intr_mask.assign(p.num_cameras, 0);
    std::vector<uint16_t> intr_candidate(p.num_cameras, 0);
    for (int c = 0; c < p.num_cameras; ++c) {
      static const uint16_t kFocal[11] = {0x1, 0x3, 0x1, 0x1, 0x3, 0x3, 0x3, 0x3, 0x1, 0x1, 0x3};
      static const uint16_t kPP[11] = {0x6, 0xc, 0x6, 0x6, 0xc, 0xc, 0xc, 0xc, 0x6, 0x6, 0xc};
      static const uint16_t kExtra[11] = {0x0, 0x0, 0x8, 0x18, 0xf0, 0xf0, 0xff0, 0x10, 0x8, 0x18,
                                          0xff0};
      const int m = p.camera_model[c];
      uint16_t mask = 0;
      if (o.refine_focal_length) mask |= kFocal[m];
      if (o.refine_principal_point) mask |= kPP[m];
      if (o.refine_extra_params) mask |= kExtra[m];
      if (p.camera_const && p.camera_const[c]) mask = 0;
      intr_candidate[c] = mask;
    }
    // keep observations that touch at least one variable block (Ceres drops the rest)
    std::vector<int64_t> cnt(P + 1, 0);
    std::vector<uint8_t> cam_used(C, 0);
    for (int64_t o2 = 0; o2 < O; ++o2) {
      const int ci = p.obs_image[o2], pi = p.obs_point[o2];
      const bool cam_const = p.pose_flags && (p.pose_flags[ci] & 1);
      if (cam_const && !pt_var[pi] && !intr_candidate[p.image_camera[ci]]) continue;
      cnt[pi + 1]++;
      cam_used[ci] = 1;
    }
    pt_start.assign(P + 1, 0);
    for (int i = 0; i < P; ++i) pt_start[i + 1] = pt_start[i] + cnt[i + 1];
    obs.resize(pt_start[P]);
    std::vector<int64_t> fill(pt_start.begin(), pt_start.end() - 1);
    for (int64_t o2 = 0; o2 < O; ++o2) {
      const int ci = p.obs_image[o2], pi = p.obs_point[o2];
      const bool cam_const = p.pose_flags && (p.pose_flags[ci] & 1);
      if (cam_const && !pt_var[pi] && !intr_candidate[p.image_camera[ci]]) continue;
      obs[fill[pi]++] = o2;
    }
    for (int i = 0; i < C; ++i) {  // cameras of images that are part of the problem
      const int c = p.image_camera[i];
      if (cam_used[i] && intr_candidate[c] && intr_block[c] < 0) {
        intr_block[c] = nintr++;
        intr_mask[c] = intr_candidate[c];
      }
    }
    for (int i = 0; i < C; ++i) {
      const uint8_t f = p.pose_flags ? p.pose_flags[i] : 0;
      if ((f & 1) || !cam_used[i]) continue;
      cam_block[i] = nblocks++;
      uint8_t m = 0x07;  // rotation always free
      for (int k = 0; k < 3; ++k)
        if (!(f & (2 << k))) m |= (uint8_t)(8 << k);
      cam_mask[i] = m;
      double nrm = 0;
      for (int k = 0; k < 4; ++k) nrm += q[4 * i + k] * q[4 * i + k];
      nrm = std::sqrt(nrm);
      if (nrm > 0) for (int k = 0; k < 4; ++k) q[4 * i + k] /= nrm;
    }
    // points without any kept observation are not part of the problem
    for (int i = 0; i < P; ++i) if (pt_start[i + 1] == pt_start[i]) pt_var[i] = 0;
    const size_t K = obs.size();
    r.resize(2 * K); Jc.resize(12 * K); Jp.resize(6 * K);
    if (nintr > 0) Ji.assign(2 * kIntrWidth * K, 0.0);
    intr_scale.assign(kIntrWidth * (size_t)nintr, 1.0);
    cam_scale.assign(6 * (size_t)nblocks, 1.0);
    pt_scale.assign(3 * (size_t)P, 1.0);
  }

  const double* CamParams(int img, const std::vector<double>& prm) const {
    return prm.data() + 12 * (size_t)pb.image_camera[img];
  }
  int CamModel(int img) const { return pb.camera_model[pb.image_camera[img]]; }

  // residuals (+ Jacobians) at (q_, t_, X_); returns cost = 0.5 sum rho(s)
  double Evaluate(const std::vector<double>& q_, const std::vector<double>& t_,
                  const std::vector<double>& X_, bool jac) {
    return Evaluate(q_, t_, X_, params, jac);
  }
  double Evaluate(const std::vector<double>& q_, const std::vector<double>& t_,
                  const std::vector<double>& X_, const std::vector<double>& prm, bool jac) {
    std::vector<double> partial(threads, 0.0);
    ParallelFor(threads, P, [&](int tid, int64_t lo, int64_t hi) {
      double cost = 0.0;
      for (int64_t pi = lo; pi < hi; ++pi) {
        for (int64_t k = pt_start[pi]; k < pt_start[pi + 1]; ++k) {
          const int64_t o2 = obs[k];
          const int ci = pb.obs_image[o2];
          double rr[2], jc[12], jp[6], ji[2 * kIntrWidth];
          const int ib = intr_block[pb.image_camera[ci]];
          if (jac) {
            LineCostTangent(CamModel(ci), CamParams(ci, prm), pb.obs_line + 3 * o2, &q_[4 * ci],
                            &t_[3 * ci], &X_[3 * pi], rr, jc, jp, ib >= 0 ? ji : nullptr);
          } else {
            LineCost<double, double>(CamModel(ci), CamParams(ci, prm), pb.obs_line + 3 * o2,
                                     &q_[4 * ci], &t_[3 * ci], &X_[3 * pi], rr);
          }
          const double s = rr[0] * rr[0] + rr[1] * rr[1];
          double rho[3];
          EvaluateLoss(opt.loss_type, opt.loss_scale, s, rho);
          cost += 0.5 * rho[0];
          if (!jac) continue;
          // corrector on the block [jc | jp | ji] (2 x 17; ji = 0 for a constant camera)
          constexpr int kW = 9 + kIntrWidth;
          double blk[2 * kW];
          for (int row = 0; row < 2; ++row) {
            for (int c = 0; c < 6; ++c) blk[kW * row + c] = jc[6 * row + c];
            for (int c = 0; c < 3; ++c) blk[kW * row + 6 + c] = jp[3 * row + c];
            for (int c = 0; c < kIntrWidth; ++c)
              blk[kW * row + 9 + c] = ib >= 0 ? ji[kIntrWidth * row + c] : 0.0;
          }
          ApplyCorrector(s, rho, rr, blk, kW);
          const int b = cam_block[ci];
          for (int row = 0; row < 2; ++row) {
            for (int c = 0; c < 6; ++c) {
              const bool on = b >= 0 && ((cam_mask[ci] >> c) & 1);
              Jc[12 * k + 6 * row + c] = on ? blk[kW * row + c] * cam_scale[6 * b + c] : 0.0;
            }
            for (int c = 0; c < 3; ++c)
              Jp[6 * k + 3 * row + c] =
                  pt_var[pi] ? blk[kW * row + 6 + c] * pt_scale[3 * pi + c] : 0.0;
            if (ib >= 0)
              for (int c = 0; c < kIntrWidth; ++c) {
                const bool on = (intr_mask[pb.image_camera[ci]] >> c) & 1;
                Ji[2 * kIntrWidth * k + kIntrWidth * row + c] =
                    on ? blk[kW * row + 9 + c] * intr_scale[kIntrWidth * ib + c] : 0.0;
              }
          }
          r[2 * k] = rr[0];
          r[2 * k + 1] = rr[1];
        }
      }
      partial[tid] = cost;
    });
    double cost = 0.0;
    for (double c : partial) cost += c;
    return cost;
  }

  // diag(J^T J) per camera block (6) and per point (3), and gradient J^T r
  // intrinsics part of J^T J and J^T r (only with nintr > 0): Uii 8x8 per intrinsics block,
  // Uic 8x6 per pose block (coupling with the block of the image's own camera), gi
  std::vector<double> Uii, Uic, gi;
  void Normal(std::vector<double>& U, std::vector<double>& gc, std::vector<double>& V,
              std::vector<double>& gp) {
    if (nintr > 0) {
      constexpr int W = kIntrWidth;
      Uii.assign(W * W * (size_t)nintr, 0.0);
      Uic.assign(W * 6 * (size_t)nblocks, 0.0);
      gi.assign(W * (size_t)nintr, 0.0);
      for (size_t k = 0; k < obs.size(); ++k) {
        const int img = pb.obs_image[obs[k]];
        const int ib = intr_block[pb.image_camera[img]], b = cam_block[img];
        if (ib < 0) continue;
        const double* ji = &Ji[2 * W * k];
        const double* jc = &Jc[12 * k];
        for (int a = 0; a < W; ++a) {
          for (int c = 0; c < W; ++c) Uii[W * W * (size_t)ib + W * a + c] += ji[a] * ji[c] + ji[W + a] * ji[W + c];
          gi[W * (size_t)ib + a] += ji[a] * r[2 * k] + ji[W + a] * r[2 * k + 1];
          if (b >= 0)
            for (int c = 0; c < 6; ++c) Uic[W * 6 * (size_t)b + 6 * a + c] += ji[a] * jc[c] + ji[W + a] * jc[6 + c];
        }
      }
    }
    U.assign(36 * (size_t)nblocks, 0.0);
    gc.assign(6 * (size_t)nblocks, 0.0);
    V.assign(9 * (size_t)P, 0.0);
    gp.assign(3 * (size_t)P, 0.0);
    for (int pi = 0; pi < P; ++pi) {
      for (int64_t k = pt_start[pi]; k < pt_start[pi + 1]; ++k) {
        const int b = cam_block[pb.obs_image[obs[k]]];
        const double* jc = &Jc[12 * k];
        const double* jp = &Jp[6 * k];
        const double* rr = &r[2 * k];
        if (b >= 0) {
          for (int a = 0; a < 6; ++a) {
            for (int c = 0; c < 6; ++c)
              U[36 * (size_t)b + 6 * a + c] += jc[a] * jc[c] + jc[6 + a] * jc[6 + c];
            gc[6 * (size_t)b + a] += jc[a] * rr[0] + jc[6 + a] * rr[1];
          }
        }
        for (int a = 0; a < 3; ++a) {
          for (int c = 0; c < 3; ++c)
            V[9 * (size_t)pi + 3 * a + c] += jp[a] * jp[c] + jp[3 + a] * jp[3 + c];
          gp[3 * (size_t)pi + a] += jp[a] * rr[0] + jp[3 + a] * rr[1];
        }
      }
    }
  }
};

int SolveBA(const orc_ba_problem& pb, const orc_ba_options& opt, orc_ba_summary* sum) {
  const auto t_start = std::chrono::steady_clock::now();
  std::memset(sum, 0, sizeof(*sum));
  Solver S(pb, opt);
  sum->num_residuals = 2 * pb.num_obs;
  sum->num_residuals_reduced = 2 * (int64_t)S.obs.size();
  if (pb.num_obs == 0) return 0;  // bundle_adjustment.cc:269-271
  constexpr int IW = kIntrWidth;
  const int nb = S.nblocks, ni = S.nintr, n = 6 * nb + IW * ni, P = S.P;
  const int ioff = 6 * nb;  // first reduced index of the intrinsics blocks
  int eff = 0;
  for (int i = 0; i < S.C; ++i) eff += __builtin_popcount(S.cam_mask[i]);
  for (int i = 0; i < P; ++i) eff += S.pt_var[i] ? 3 : 0;
  for (int c = 0; c < pb.num_cameras; ++c)
    if (S.intr_block[c] >= 0) eff += __builtin_popcount(S.intr_mask[c]);
  sum->num_effective_parameters_reduced = eff;

  auto secs = [](std::chrono::steady_clock::time_point a) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
  };
  auto write_back = [&]() {
    std::memcpy(pb.qvecs, S.q.data(), sizeof(double) * S.q.size());
    std::memcpy(pb.tvecs, S.t.data(), sizeof(double) * S.t.size());
    std::memcpy(pb.points, S.X.data(), sizeof(double) * S.X.size());
    if (ni > 0) std::memcpy(pb.camera_params, S.params.data(), sizeof(double) * S.params.size());
  };
  if (S.obs.empty() || eff == 0) {
    sum->termination_type = 0;
    write_back();
    sum->total_time_s = secs(t_start);
    return 1;
  }

  std::vector<double> U, gc, V, gp;
  auto tj = std::chrono::steady_clock::now();
  double cost = S.Evaluate(S.q, S.t, S.X, true);
  S.Normal(U, gc, V, gp);
  if (opt.jacobi_scaling) {
    for (int b = 0; b < nb; ++b)
      for (int a = 0; a < 6; ++a)
        S.cam_scale[6 * b + a] = 1.0 / (1.0 + std::sqrt(U[36 * (size_t)b + 7 * a]));
    for (int p = 0; p < P; ++p)
      for (int a = 0; a < 3; ++a)
        S.pt_scale[3 * p + a] = 1.0 / (1.0 + std::sqrt(V[9 * (size_t)p + 4 * a]));
    for (int ib = 0; ib < ni; ++ib)
      for (int a = 0; a < IW; ++a)
        S.intr_scale[IW * ib + a] = 1.0 / (1.0 + std::sqrt(S.Uii[IW * IW * (size_t)ib + (IW + 1) * a]));
    cost = S.Evaluate(S.q, S.t, S.X, true);
    S.Normal(U, gc, V, gp);
  }
  sum->jacobian_time_s += secs(tj);
  sum->initial_cost = cost;

  // gradient max-norm of the UNSCALED problem (g = g_scaled / scale): max |x - Plus(x, -g)|
  auto gradient_max_norm = [&]() {
    double m = 0.0;
    for (int i = 0; i < S.C; ++i) {
      const int b = S.cam_block[i];
      if (b < 0) continue;
      double g[6];
      for (int a = 0; a < 6; ++a) g[a] = gc[6 * b + a] / S.cam_scale[6 * b + a];  // J_s = J diag(s)
      const double nd[3] = {-g[0], -g[1], -g[2]};
      double qp[4];
      QuaternionPlus(&S.q[4 * i], nd, qp);
      for (int k = 0; k < 4; ++k) m = std::max(m, std::fabs(qp[k] - S.q[4 * i + k]));
      for (int a = 3; a < 6; ++a) m = std::max(m, std::fabs(g[a]));
    }
    for (int p = 0; p < P; ++p)
      for (int a = 0; a < 3; ++a) m = std::max(m, std::fabs(gp[3 * p + a] / S.pt_scale[3 * p + a]));
    for (int ib = 0; ib < ni; ++ib)  // SubsetParameterization: Plus(x, d) = x + d
      for (int a = 0; a < IW; ++a) m = std::max(m, std::fabs(S.gi[IW * ib + a] / S.intr_scale[IW * ib + a]));
    return m;
  };

  int tl = 0;
  auto trace = [&](double c, double radius, int acc) {
    if (tl < ORC_BA_MAX_TRACE) {
      sum->trace_cost[tl] = c;
      sum->trace_radius[tl] = radius;
      sum->trace_accepted[tl] = acc;
      ++tl;
    }
  };
  double radius = opt.initial_trust_region_radius;
  double decrease_factor = 2.0;
  trace(cost, radius, 1);
  double gmax = gradient_max_norm();
  sum->termination_type = 1;
  int invalid = 0;
  bool done = gmax <= opt.gradient_tolerance;
  if (done) sum->termination_type = 0;

  std::vector<double> Smat, rhs, Vinv(9 * (size_t)P), dp(3 * (size_t)P), dc;
  std::vector<double> qn, tn, Xn;
  for (int iter = 0; !done && iter < opt.max_num_iterations; ++iter) {
    auto tl0 = std::chrono::steady_clock::now();
    // --- LM-damped normal equations, Schur complement on the cameras
    Smat.assign((size_t)n * n, 0.0);
    rhs.assign(n, 0.0);
    for (int b = 0; b < nb; ++b) {
      for (int a = 0; a < 6; ++a) {
        for (int c = 0; c < 6; ++c) Smat[(size_t)(6 * b + a) * n + 6 * b + c] = U[36 * (size_t)b + 6 * a + c];
        const double d = std::min(std::max(U[36 * (size_t)b + 7 * a], opt.min_lm_diagonal),
                                  opt.max_lm_diagonal);
        Smat[(size_t)(6 * b + a) * n + 6 * b + a] += d / radius;
        rhs[6 * b + a] = -gc[6 * b + a];
      }
    }
    for (int ib = 0; ib < ni; ++ib)
      for (int a = 0; a < IW; ++a) {
        const int ra = ioff + IW * ib + a;
        for (int c = 0; c < IW; ++c) Smat[(size_t)ra * n + ioff + IW * ib + c] = S.Uii[IW * IW * (size_t)ib + IW * a + c];
        const double d = std::min(std::max(S.Uii[IW * IW * (size_t)ib + (IW + 1) * a], opt.min_lm_diagonal),
                                  opt.max_lm_diagonal);
        Smat[(size_t)ra * n + ra] += d / radius;
        rhs[ra] = -S.gi[IW * ib + a];
      }
    if (ni > 0)
      for (int i = 0; i < S.C; ++i) {
        const int b = S.cam_block[i], ib = S.intr_block[pb.image_camera[i]];
        if (b < 0 || ib < 0) continue;
        for (int a = 0; a < IW; ++a)
          for (int c = 0; c < 6; ++c)
            Smat[(size_t)(ioff + IW * ib + a) * n + 6 * b + c] = S.Uic[IW * 6 * (size_t)b + 6 * a + c];
      }
    for (int p = 0; p < P; ++p) {
      double Vd[9];
      for (int k = 0; k < 9; ++k) Vd[k] = V[9 * (size_t)p + k];
      for (int a = 0; a < 3; ++a) {
        const double d = std::min(std::max(V[9 * (size_t)p + 4 * a], opt.min_lm_diagonal),
                                  opt.max_lm_diagonal);
        Vd[4 * a] += d / radius;
      }
      if (!S.pt_var[p] || !Invert3(Vd, &Vinv[9 * (size_t)p]))
        for (int k = 0; k < 9; ++k) Vinv[9 * (size_t)p + k] = 0.0;
    }
    {
      const int nlocks = 256;
      std::vector<std::mutex> locks(nlocks);
      std::mutex intr_lock;
      ParallelFor(S.threads, P, [&](int, int64_t lo, int64_t hi) {
        std::vector<double> Wb, Zb, Wi, Zi;
        std::vector<int> blk, iblk;
        for (int64_t p = lo; p < hi; ++p) {
          if (!S.pt_var[p]) continue;
          const int64_t k0 = S.pt_start[p], k1 = S.pt_start[p + 1];
          const int m = (int)(k1 - k0);
          Wb.assign(18 * (size_t)m, 0.0);
          Zb.assign(18 * (size_t)m, 0.0);
          blk.assign(m, -1);
          if (ni > 0) {
            Wi.assign(3 * IW * (size_t)m, 0.0);
            Zi.assign(3 * IW * (size_t)m, 0.0);
            iblk.assign(m, -1);
          }
          const double* vi = &Vinv[9 * (size_t)p];
          double vg[3];
          for (int a = 0; a < 3; ++a)
            vg[a] = vi[3 * a] * gp[3 * p] + vi[3 * a + 1] * gp[3 * p + 1] + vi[3 * a + 2] * gp[3 * p + 2];
          for (int e = 0; e < m; ++e) {
            const int64_t k = k0 + e;
            const int b = S.cam_block[pb.obs_image[S.obs[k]]];
            blk[e] = b;
            if (ni > 0) {
              const int ib = S.intr_block[pb.image_camera[pb.obs_image[S.obs[k]]]];
              iblk[e] = ib;
              if (ib >= 0) {
                const double* ji = &S.Ji[2 * IW * k];
                const double* jp = &S.Jp[6 * k];
                for (int a = 0; a < IW; ++a)
                  for (int c = 0; c < 3; ++c) {
                    Wi[3 * IW * (size_t)e + 3 * a + c] = ji[a] * jp[c] + ji[IW + a] * jp[3 + c];
                  }
                for (int a = 0; a < IW; ++a)
                  for (int c = 0; c < 3; ++c) {
                    const double* w = &Wi[3 * IW * (size_t)e + 3 * a];
                    Zi[3 * IW * (size_t)e + 3 * a + c] = w[0] * vi[c] + w[1] * vi[3 + c] + w[2] * vi[6 + c];
                  }
              }
            }
            if (b < 0) continue;
            const double* jc = &S.Jc[12 * k];
            const double* jp = &S.Jp[6 * k];
            double* W = &Wb[18 * (size_t)e];
            for (int a = 0; a < 6; ++a)
              for (int c = 0; c < 3; ++c) W[3 * a + c] = jc[a] * jp[c] + jc[6 + a] * jp[3 + c];
            double* Z = &Zb[18 * (size_t)e];
            for (int a = 0; a < 6; ++a)
              for (int c = 0; c < 3; ++c)
                Z[3 * a + c] = W[3 * a] * vi[c] + W[3 * a + 1] * vi[3 + c] + W[3 * a + 2] * vi[6 + c];
          }
          if (ni > 0) {  // rows of the intrinsics blocks: (intr, pose) and (intr, intr) products
            std::lock_guard<std::mutex> g(intr_lock);
            for (int e = 0; e < m; ++e) {
              const int ie = iblk[e];
              if (ie < 0) continue;
              const double* Z = &Zi[3 * IW * (size_t)e];
              const double* We = &Wi[3 * IW * (size_t)e];
              for (int a = 0; a < IW; ++a)
                rhs[ioff + IW * ie + a] += We[3 * a] * vg[0] + We[3 * a + 1] * vg[1] + We[3 * a + 2] * vg[2];
              for (int f = 0; f < m; ++f) {
                if (blk[f] >= 0) {
                  const double* W2 = &Wb[18 * (size_t)f];
                  for (int a = 0; a < IW; ++a)
                    for (int c = 0; c < 6; ++c)
                      Smat[(size_t)(ioff + IW * ie + a) * n + 6 * blk[f] + c] -=
                          Z[3 * a] * W2[3 * c] + Z[3 * a + 1] * W2[3 * c + 1] + Z[3 * a + 2] * W2[3 * c + 2];
                }
                if (iblk[f] >= 0 && iblk[f] <= ie) {
                  const double* W2 = &Wi[3 * IW * (size_t)f];
                  for (int a = 0; a < IW; ++a)
                    for (int c = 0; c < IW; ++c)
                      Smat[(size_t)(ioff + IW * ie + a) * n + ioff + IW * iblk[f] + c] -=
                          Z[3 * a] * W2[3 * c] + Z[3 * a + 1] * W2[3 * c + 1] + Z[3 * a + 2] * W2[3 * c + 2];
                }
              }
            }
          }
          for (int e = 0; e < m; ++e) {
            const int bi = blk[e];
            if (bi < 0) continue;
            std::lock_guard<std::mutex> g(locks[bi % nlocks]);
            const double* Z = &Zb[18 * (size_t)e];
            for (int a = 0; a < 6; ++a)
              rhs[6 * bi + a] += Wb[18 * (size_t)e + 3 * a] * vg[0] + Wb[18 * (size_t)e + 3 * a + 1] * vg[1] +
                                 Wb[18 * (size_t)e + 3 * a + 2] * vg[2];
            for (int f = 0; f < m; ++f) {
              const int bj = blk[f];
              if (bj < 0 || bj > bi) continue;  // lower triangle only
              const double* W2 = &Wb[18 * (size_t)f];
              for (int a = 0; a < 6; ++a)
                for (int c = 0; c < 6; ++c)
                  Smat[(size_t)(6 * bi + a) * n + 6 * bj + c] -=
                      Z[3 * a] * W2[3 * c] + Z[3 * a + 1] * W2[3 * c + 1] + Z[3 * a + 2] * W2[3 * c + 2];
            }
          }
        }
      });
    }
    // masked tangent dims: identity rows
    for (int i = 0; i < S.C; ++i) {
      const int b = S.cam_block[i];
      if (b < 0) continue;
      for (int a = 0; a < 6; ++a)
        if (!((S.cam_mask[i] >> a) & 1)) {
          Smat[(size_t)(6 * b + a) * n + 6 * b + a] = 1.0;
          rhs[6 * b + a] = 0.0;
        }
    }
    for (int c = 0; c < pb.num_cameras; ++c) {  // constant parameters of a variable camera
      const int ib = S.intr_block[c];
      if (ib < 0) continue;
      for (int a = 0; a < IW; ++a)
        if (!((S.intr_mask[c] >> a) & 1)) {
          Smat[(size_t)(ioff + IW * ib + a) * n + ioff + IW * ib + a] = 1.0;
          rhs[ioff + IW * ib + a] = 0.0;
        }
    }
    bool ok = n == 0 || CholeskyLower(Smat, n, S.threads);
    dc = rhs;
    if (ok && n > 0) CholeskySolve(Smat, n, dc);
    // back-substitution: dp = -Vinv (gp + W^T dc)
    double model_cost_change = 0.0;
    if (ok) {
      for (int p = 0; p < P; ++p) {
        double acc[3] = {gp[3 * p], gp[3 * p + 1], gp[3 * p + 2]};
        for (int64_t k = S.pt_start[p]; k < S.pt_start[p + 1]; ++k) {
          const int b = S.cam_block[pb.obs_image[S.obs[k]]];
          const int ib = ni > 0 ? S.intr_block[pb.image_camera[pb.obs_image[S.obs[k]]]] : -1;
          if (b < 0 && ib < 0) continue;
          const double* jc = &S.Jc[12 * k];
          const double* jp = &S.Jp[6 * k];
          double u0 = 0, u1 = 0;
          if (b >= 0)
            for (int a = 0; a < 6; ++a) { u0 += jc[a] * dc[6 * b + a]; u1 += jc[6 + a] * dc[6 * b + a]; }
          if (ib >= 0)
            for (int a = 0; a < IW; ++a) {
              u0 += S.Ji[2 * IW * k + a] * dc[ioff + IW * ib + a];
              u1 += S.Ji[2 * IW * k + IW + a] * dc[ioff + IW * ib + a];
            }
          for (int c = 0; c < 3; ++c) acc[c] += jp[c] * u0 + jp[3 + c] * u1;
        }
        const double* vi = &Vinv[9 * (size_t)p];
        for (int a = 0; a < 3; ++a)
          dp[3 * p + a] = -(vi[3 * a] * acc[0] + vi[3 * a + 1] * acc[1] + vi[3 * a + 2] * acc[2]);
      }
      // model_cost_change = -sum (J d) . (r + J d / 2)
      for (int p = 0; p < P; ++p)
        for (int64_t k = S.pt_start[p]; k < S.pt_start[p + 1]; ++k) {
          const int b = S.cam_block[pb.obs_image[S.obs[k]]];
          const double* jc = &S.Jc[12 * k];
          const double* jp = &S.Jp[6 * k];
          double m0 = 0, m1 = 0;
          if (b >= 0) for (int a = 0; a < 6; ++a) { m0 += jc[a] * dc[6 * b + a]; m1 += jc[6 + a] * dc[6 * b + a]; }
          if (ni > 0) {
            const int ib = S.intr_block[pb.image_camera[pb.obs_image[S.obs[k]]]];
            if (ib >= 0)
              for (int a = 0; a < IW; ++a) {
                m0 += S.Ji[2 * IW * k + a] * dc[ioff + IW * ib + a];
                m1 += S.Ji[2 * IW * k + IW + a] * dc[ioff + IW * ib + a];
              }
          }
          for (int c = 0; c < 3; ++c) { m0 += jp[c] * dp[3 * p + c]; m1 += jp[3 + c] * dp[3 * p + c]; }
          model_cost_change -= m0 * (S.r[2 * k] + m0 / 2.0) + m1 * (S.r[2 * k + 1] + m1 / 2.0);
        }
    }
    sum->linear_solver_time_s += secs(tl0);

    // --- step evaluation, in the order of ceres::internal::TrustRegionMinimizer::Minimize:
    // invalid step? -> candidate cost -> parameter tolerance -> function tolerance -> accept/reject
    bool accepted = false;
    double cost_new = cost;
    if (!ok || !(model_cost_change > 0.0)) {
      ++sum->num_unsuccessful_steps;
      if (++invalid >= opt.max_num_consecutive_invalid_steps) {
        sum->termination_type = 2;
        trace(cost, radius, 0);
        break;
      }
      radius = radius / decrease_factor;  // LevenbergMarquardtStrategy::StepIsInvalid
      decrease_factor *= 2.0;
      trace(cost, radius, 0);
      continue;
    }
    invalid = 0;
    qn = S.q; tn = S.t; Xn = S.X;
    double step_sq = 0.0, x_sq = 0.0;
    for (int i = 0; i < S.C; ++i) {
      const int b = S.cam_block[i];
      if (b < 0) continue;
      double d[6];
      for (int a = 0; a < 6; ++a) d[a] = dc[6 * b + a] * S.cam_scale[6 * b + a];
      QuaternionPlus(&S.q[4 * i], d, &qn[4 * i]);
      for (int a = 0; a < 3; ++a) tn[3 * i + a] = S.t[3 * i + a] + d[3 + a];
      for (int k = 0; k < 4; ++k) {
        step_sq += (qn[4 * i + k] - S.q[4 * i + k]) * (qn[4 * i + k] - S.q[4 * i + k]);
        x_sq += S.q[4 * i + k] * S.q[4 * i + k];
      }
      for (int k = 0; k < 3; ++k) {
        step_sq += d[3 + k] * d[3 + k];
        x_sq += S.t[3 * i + k] * S.t[3 * i + k];
      }
    }
    for (int p = 0; p < P; ++p) {
      if (!S.pt_var[p]) continue;
      for (int a = 0; a < 3; ++a) {
        const double d = dp[3 * p + a] * S.pt_scale[3 * p + a];
        Xn[3 * p + a] = S.X[3 * p + a] + d;
        step_sq += d * d;
        x_sq += S.X[3 * p + a] * S.X[3 * p + a];
      }
    }
    std::vector<double> prm_n = S.params;
    for (int c = 0; c < pb.num_cameras; ++c) {
      const int ib = S.intr_block[c];
      if (ib < 0) continue;
      for (int a = 0; a < kNumParams[pb.camera_model[c]]; ++a) {
        const double d = ((S.intr_mask[c] >> a) & 1) ? dc[ioff + IW * ib + a] * S.intr_scale[IW * ib + a] : 0.0;
        prm_n[12 * (size_t)c + a] = S.params[12 * (size_t)c + a] + d;
        step_sq += d * d;
        x_sq += S.params[12 * (size_t)c + a] * S.params[12 * (size_t)c + a];
      }
    }
    cost_new = S.Evaluate(qn, tn, Xn, prm_n, false);
    if (std::sqrt(step_sq) <= opt.parameter_tolerance * (std::sqrt(x_sq) + opt.parameter_tolerance)) {
      sum->termination_type = 0;  // parameter tolerance: the candidate is not applied
      break;
    }
    const double cost_change = cost - cost_new;
    if (std::fabs(cost_change) <= opt.function_tolerance * cost) {
      sum->termination_type = 0;  // function tolerance: the candidate is not applied
      break;
    }
    const double relative_decrease = cost_change / model_cost_change;
    accepted = relative_decrease > opt.min_relative_decrease;
    if (accepted) {
      const double tmp = 2.0 * relative_decrease - 1.0;
      radius = radius / std::max(1.0 / 3.0, 1.0 - tmp * tmp * tmp);
      radius = std::min(opt.max_trust_region_radius, radius);
      decrease_factor = 2.0;
      S.q.swap(qn); S.t.swap(tn); S.X.swap(Xn);
      S.params.swap(prm_n);
      tj = std::chrono::steady_clock::now();
      cost = S.Evaluate(S.q, S.t, S.X, true);
      S.Normal(U, gc, V, gp);
      sum->jacobian_time_s += secs(tj);
      ++sum->num_successful_steps;
      trace(cost, radius, 1);
      gmax = gradient_max_norm();
      if (gmax <= opt.gradient_tolerance) { sum->termination_type = 0; done = true; }
    } else {
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      ++sum->num_unsuccessful_steps;
      trace(cost, radius, 0);
      if (radius < opt.min_trust_region_radius) { sum->termination_type = 0; done = true; }
    }
  }
  sum->final_cost = cost;
  sum->final_gradient_max_norm = gmax;
  sum->trace_len = tl;
  write_back();
  sum->total_time_s = secs(t_start);
  return 1;
}

}  // namespace

extern "C" {

void orc_ba_options_default(orc_ba_options* o) {
  // BundleAdjustmentOptions() (src/optim/bundle_adjustment.h:80-93) + ceres::Solver::Options
  o->loss_type = 0;
  o->loss_scale = 1.0;
  o->max_num_iterations = 100;
  o->function_tolerance = 0.0;
  o->gradient_tolerance = 0.0;
  o->parameter_tolerance = 0.0;
  o->max_num_consecutive_invalid_steps = 10;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->jacobi_scaling = 1;
  o->num_threads = -1;
  o->refine_focal_length = 0;  // intrinsics constant (the pose-refinement / test default; the
  o->refine_principal_point = 0;  // mapper sets them per its options)
  o->refine_extra_params = 0;
}

int orc_ba_solve(const orc_ba_problem* problem, const orc_ba_options* options,
                 orc_ba_summary* summary) {
  return SolveBA(*problem, *options, summary);
}

void orc_line_cost(int model, const double* cam, const double* line, const double* q,
                   const double* t, const double* X, double* r, double* jq, double* jt,
                   double* jX) {
  LineCostAutoDiff(model, cam, line, q, t, X, r, jq, jt, jX);
}

void orc_line_cost_intr(int model, const double* cam, const double* line, const double* q,
                        const double* t, const double* X, double* r, double* jq, double* jt,
                        double* jX, double* jcam) {
  LineCostAutoDiffIntr(model, cam, line, q, t, X, r, jq, jt, jX, jcam);
}

void orc_line_cost_tangent(int model, const double* cam, const double* line, const double* q,
                           const double* t, const double* X, double* r, double* jc, double* jX) {
  LineCostTangent(model, cam, line, q, t, X, r, jc, jX);
}

double orc_ba_cost(const orc_ba_problem* problem, const orc_ba_options* options) {
  Solver S(*problem, *options);
  return S.Evaluate(S.q, S.t, S.X, false);
}

void orc_quaternion_plus(const double* q, const double* delta, double* out) {
  QuaternionPlus(q, delta, out);
}

int orc_refine_absolute_pose(const double* lines, const double* points, const uint8_t* mask,
                             size_t n, int model, const double* cam, double gradient_tolerance,
                             int max_num_iterations, double loss_scale, double* qvec,
                             double* tvec, orc_ba_summary* summary) {
  // src/estimators/pose.cc:96-213: one residual block per inlier, Cauchy loss, points constant,
  // intrinsics constant (refine_focal_length = refine_extra_params = false), quaternion
  // parameterisation; ceres::Solver::Options defaults except gradient_tolerance,
  // max_num_iterations, DENSE_QR.
  std::vector<int32_t> oi, op;
  std::vector<double> ol, pts;
  for (size_t i = 0; i < n; ++i) {
    if (!mask[i]) continue;
    oi.push_back(0);
    op.push_back((int32_t)(pts.size() / 3));
    for (int k = 0; k < 3; ++k) { ol.push_back(lines[3 * i + k]); pts.push_back(points[3 * i + k]); }
  }
  const int np = (int)(pts.size() / 3);
  std::vector<uint8_t> pc(np, 1);
  uint8_t flags = 0;
  int32_t icam = 0;
  double params[12] = {0};
  const int nparams[11] = {3, 4, 4, 5, 8, 8, 12, 5, 4, 5, 12};
  for (int k = 0; k < nparams[model]; ++k) params[k] = cam[k];
  // *qvec = NormalizeQuaternion(*qvec) (pose.cc:143)
  double nrm = std::sqrt(qvec[0] * qvec[0] + qvec[1] * qvec[1] + qvec[2] * qvec[2] + qvec[3] * qvec[3]);
  if (np > 0 && nrm > 0) for (int k = 0; k < 4; ++k) qvec[k] /= nrm;
  orc_ba_problem pb;
  pb.num_images = 1; pb.qvecs = qvec; pb.tvecs = tvec; pb.pose_flags = &flags; pb.image_camera = &icam;
  pb.num_cameras = 1; pb.camera_model = &model; pb.camera_params = params;
  pb.num_points = np; pb.points = pts.data(); pb.point_const = pc.data();
  pb.num_obs = np; pb.obs_image = oi.data(); pb.obs_point = op.data(); pb.obs_line = ol.data();
  pb.camera_const = nullptr;
  orc_ba_options o;
  orc_ba_options_default(&o);
  o.loss_type = 2;
  o.loss_scale = loss_scale;
  o.gradient_tolerance = gradient_tolerance;
  o.max_num_iterations = max_num_iterations;
  o.function_tolerance = 1e-6;   // ceres::Solver::Options defaults (pose.cc:187-190 sets only
  o.parameter_tolerance = 1e-8;  // gradient_tolerance, max_num_iterations, linear_solver_type)
  o.num_threads = 1;
  orc_ba_summary local;
  orc_ba_summary* s = summary ? summary : &local;
  if (np == 0) {  // empty problem: Ceres returns CONVERGENCE immediately; usable
    std::memset(s, 0, sizeof(*s));
    return 1;
  }
  SolveBA(pb, o, s);
  return s->termination_type != 2;  // Summary::IsSolutionUsable()
}

}  // extern "C"
