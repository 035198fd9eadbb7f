// Stand-in for <ceres/ceres.h> + <ceres/rotation.h> — TEST INFRASTRUCTURE ONLY (oracle/build_ref.sh).
// Ceres is absent in this image (and unpinned upstream).  The reference's cost functors
// (src/base/cost_functions.h) and camera models (src/base/camera_models.h) need four things from
// it: the scalar functions ceres::sqrt / atan / tan / sin / cos, the dual-number type their
// templates are instantiated with by AutoDiffCostFunction (ceres::Jet), UnitQuaternionRotatePoint,
// and the CostFunction / AutoDiffCostFunction class names of their Create() factories.  These are
// restated here from Ceres' published sources (1.14 - 2.0; from memory, the library is not in
// the image): what tests/test_ref_cost.py pins is the reference's functor and camera-model TEXT,
// not Ceres.
#pragma once
#include <cmath>
#include <memory>

#include <glog/logging.h>  // (real Ceres headers pull glog in; cost_functions.h relies on it for CHECK_NEAR)

namespace ceres {

using std::abs;
using std::atan;
using std::cos;
using std::sin;
using std::sqrt;
using std::tan;

// ceres/jet.h: a + sum_i v[i] e_i, e_i e_j = 0
template <typename T, int N>
struct Jet {
  T a;
  T v[N];
  Jet() : a() {
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(const T& value) : a(value) {  // NOLINT (implicit like ceres::Jet)
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(int value) : a(value) {  // NOLINT: T(2) in the functors
    for (int i = 0; i < N; ++i) v[i] = T();
  }
  Jet(const T& value, int k) : a(value) {
    for (int i = 0; i < N; ++i) v[i] = T();
    v[k] = T(1);
  }
  Jet& operator+=(const Jet& y) { return *this = *this + y; }
  Jet& operator-=(const Jet& y) { return *this = *this - y; }
  Jet& operator*=(const Jet& y) { return *this = *this * y; }
  Jet& operator/=(const Jet& y) { return *this = *this / y; }
};

template <typename T, int N>
Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
template <typename T, int N>
Jet<T, N> operator-(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = -f.a;
  for (int i = 0; i < N; ++i) r.v[i] = -f.v[i];
  return r;
}
template <typename T, int N>
Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r;
  r.a = f.a + g.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] + g.v[i];
  return r;
}
template <typename T, int N>
Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r;
  r.a = f.a - g.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] - g.v[i];
  return r;
}
template <typename T, int N>
Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> r;
  r.a = f.a * g.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.a * g.v[i] + f.v[i] * g.a;
  return r;
}
template <typename T, int N>
Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
  // ceres/jet.h: one reciprocal, b/c = b * c^-1 and (db - (b/c) dc) * c^-1
  Jet<T, N> r;
  const T g_a_inverse = T(1.0) / g.a;
  r.a = f.a * g_a_inverse;
  for (int i = 0; i < N; ++i) r.v[i] = (f.v[i] - r.a * g.v[i]) * g_a_inverse;
  return r;
}
// mixed with scalars
template <typename T, int N>
Jet<T, N> operator+(const Jet<T, N>& f, T s) { return f + Jet<T, N>(s); }
template <typename T, int N>
Jet<T, N> operator+(T s, const Jet<T, N>& f) { return Jet<T, N>(s) + f; }
template <typename T, int N>
Jet<T, N> operator-(const Jet<T, N>& f, T s) { return f - Jet<T, N>(s); }
template <typename T, int N>
Jet<T, N> operator-(T s, const Jet<T, N>& f) { return Jet<T, N>(s) - f; }
template <typename T, int N>
Jet<T, N> operator*(const Jet<T, N>& f, T s) { return f * Jet<T, N>(s); }
template <typename T, int N>
Jet<T, N> operator*(T s, const Jet<T, N>& f) { return Jet<T, N>(s) * f; }
template <typename T, int N>
Jet<T, N> operator/(const Jet<T, N>& f, T s) { return f / Jet<T, N>(s); }
template <typename T, int N>
Jet<T, N> operator/(T s, const Jet<T, N>& f) { return Jet<T, N>(s) / f; }

#define MINICERES_CMP(op)                                                                 \
  template <typename T, int N>                                                            \
  bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; }         \
  template <typename T, int N>                                                            \
  bool operator op(const Jet<T, N>& f, T s) { return f.a op s; }                          \
  template <typename T, int N>                                                            \
  bool operator op(T s, const Jet<T, N>& f) { return s op f.a; }
MINICERES_CMP(<)
MINICERES_CMP(<=)
MINICERES_CMP(>)
MINICERES_CMP(>=)
MINICERES_CMP(==)
MINICERES_CMP(!=)
#undef MINICERES_CMP

template <typename T, int N>
Jet<T, N> abs(const Jet<T, N>& f) { return f.a < T(0.0) ? -f : f; }
template <typename T, int N>
Jet<T, N> sqrt(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::sqrt(f.a);
  const T two_a_inverse = T(1.0) / (T(2.0) * r.a);
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * two_a_inverse;
  return r;
}
template <typename T, int N>
Jet<T, N> atan(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::atan(f.a);
  const T tmp = T(1.0) / (T(1.0) + f.a * f.a);
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * tmp;
  return r;
}
template <typename T, int N>
Jet<T, N> tan(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::tan(f.a);
  const T tmp = T(1.0) + r.a * r.a;
  for (int i = 0; i < N; ++i) r.v[i] = f.v[i] * tmp;
  return r;
}
template <typename T, int N>
Jet<T, N> sin(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::sin(f.a);
  const T c = std::cos(f.a);
  for (int i = 0; i < N; ++i) r.v[i] = c * f.v[i];
  return r;
}
template <typename T, int N>
Jet<T, N> cos(const Jet<T, N>& f) {
  Jet<T, N> r;
  r.a = std::cos(f.a);
  const T s = -std::sin(f.a);
  for (int i = 0; i < N; ++i) r.v[i] = s * f.v[i];
  return r;
}

// ceres/rotation.h (1.14 - 2.0): rotation by a UNIT quaternion (w, x, y, z)
template <typename T>
inline void UnitQuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  const T t2 = q[0] * q[1];
  const T t3 = q[0] * q[2];
  const T t4 = q[0] * q[3];
  const T t5 = -q[1] * q[1];
  const T t6 = q[1] * q[2];
  const T t7 = q[1] * q[3];
  const T t8 = -q[2] * q[2];
  const T t9 = q[2] * q[3];
  const T t1 = -q[3] * q[3];
  result[0] = T(2) * ((t8 + t1) * pt[0] + (t6 - t4) * pt[1] + (t3 + t7) * pt[2]) + pt[0];
  result[1] = T(2) * ((t4 + t6) * pt[0] + (t5 + t1) * pt[1] + (t9 - t2) * pt[2]) + pt[1];
  result[2] = T(2) * ((t7 - t3) * pt[0] + (t2 + t9) * pt[1] + (t5 + t8) * pt[2]) + pt[2];
}

// the class names of the functors' Create() factories (never evaluated through here)
class CostFunction {
 public:
  virtual ~CostFunction() {}
};
template <typename Functor, int kNumResiduals, int... Ns>
class AutoDiffCostFunction : public CostFunction {
  std::unique_ptr<Functor> functor_;

 public:
  explicit AutoDiffCostFunction(Functor* functor) : functor_(functor) {}
  const Functor& functor() const { return *functor_; }
};

}  // namespace ceres
