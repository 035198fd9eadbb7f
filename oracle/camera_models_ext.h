// camera_models_ext.h — CameraModel::WorldToImage<T> of the six remaining COLMAP camera models
// (src/base/camera_models.h: OPENCV_FISHEYE :929-986, FULL_OPENCV :1024-1080, FOV :1103-1166,
// SIMPLE_RADIAL_FISHEYE :1238-1290, RADIAL_FISHEYE :1318-1366, THIN_PRISM_FISHEYE :1405-1481),
// restated expression by expression.  TEST INFRASTRUCTURE (oracle).  T is double or a jet; PT the
// type of the parameters (double constants, or jets when intrinsics are refined).  The including
// file provides AsT<T>(PT), Lift<T>::C(double) and sqrt / atan / tan for T.
#pragma once
#include <cmath>
#include <limits>

namespace orc_cam {

using std::atan;
using std::sqrt;
using std::tan;

template <typename T> inline double Value(const T& x) { return x.a; }
inline double Value(double x) { return x; }

// theta_d / r scaling shared by the fisheye models: (du, dv) = (u, v) * thetad / r - (u, v)
template <typename T, typename F>
inline void FisheyeDistortion(const T u, const T v, F thetad_of_theta, T* du, T* dv, T zero) {
  const T r = sqrt(u * u + v * v);
  if (Value(r) > std::numeric_limits<double>::epsilon()) {
    const T theta = atan(r);
    const T thetad = thetad_of_theta(theta);
    *du = u * thetad / r - u;
    *dv = v * thetad / r - v;
  } else {
    *du = zero;
    *dv = zero;
  }
}

template <typename T, typename PT, typename LiftC, typename LiftP>
bool WorldToImageExt(int model, const PT* p, const T u, const T v, T* x, T* y, LiftC C, LiftP P) {
  switch (model) {
    case 5: {  // OPENCV_FISHEYE fx, fy, cx, cy, k1, k2, k3, k4
      T du, dv;
      FisheyeDistortion(u, v, [&](const T& theta) {
        const T theta2 = theta * theta;
        const T theta4 = theta2 * theta2;
        const T theta6 = theta4 * theta2;
        const T theta8 = theta4 * theta4;
        return theta * (C(1.0) + P(4) * theta2 + P(5) * theta4 + P(6) * theta6 + P(7) * theta8);
      }, &du, &dv, C(0.0));
      *x = u + du;
      *y = v + dv;
      *x = P(0) * *x + P(2);
      *y = P(1) * *y + P(3);
      return true;
    }
    case 6: {  // FULL_OPENCV fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, k5, k6
      const T u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2, r4 = r2 * r2, r6 = r4 * r2;
      const T radial = (C(1.0) + P(4) * r2 + P(5) * r4 + P(8) * r6) /
                       (C(1.0) + P(9) * r2 + P(10) * r4 + P(11) * r6);
      const T du = u * radial + C(2.0) * P(6) * uv + P(7) * (r2 + C(2.0) * u2) - u;
      const T dv = v * radial + C(2.0) * P(7) * uv + P(6) * (r2 + C(2.0) * v2) - v;
      *x = u + du;
      *y = v + dv;
      *x = P(0) * *x + P(2);
      *y = P(1) * *y + P(3);
      return true;
    }
    case 7: {  // FOV fx, fy, cx, cy, omega
      const T omega = P(4);
      const double kEpsilon = 1e-4;
      const T radius2 = u * u + v * v;
      const T omega2 = omega * omega;
      T factor;
      if (Value(omega2) < kEpsilon) {
        factor = (omega2 * radius2) / C(3.0) - omega2 / C(12.0) + C(1.0);
      } else if (Value(radius2) < kEpsilon) {
        const T tan_half_omega = tan(omega / C(2.0));
        factor = (C(-2.0) * tan_half_omega *
                  (C(4.0) * radius2 * tan_half_omega * tan_half_omega - C(3.0))) /
                 (C(3.0) * omega);
      } else {
        const T radius = sqrt(radius2);
        const T numerator = atan(radius * C(2.0) * tan(omega / C(2.0)));
        factor = numerator / (radius * omega);
      }
      *x = u * factor;
      *y = v * factor;
      *x = P(0) * *x + P(2);
      *y = P(1) * *y + P(3);
      return true;
    }
    case 8:    // SIMPLE_RADIAL_FISHEYE f, cx, cy, k
    case 9: {  // RADIAL_FISHEYE f, cx, cy, k1, k2
      T du, dv;
      FisheyeDistortion(u, v, [&](const T& theta) {
        const T theta2 = theta * theta;
        if (model == 8) return theta * (C(1.0) + P(3) * theta2);
        const T theta4 = theta2 * theta2;
        return theta * (C(1.0) + P(3) * theta2 + P(4) * theta4);
      }, &du, &dv, C(0.0));
      *x = u + du;
      *y = v + dv;
      *x = P(0) * *x + P(1);
      *y = P(0) * *y + P(2);
      return true;
    }
    case 10: {  // THIN_PRISM_FISHEYE fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, sx1, sy1
      const T r = sqrt(u * u + v * v);
      T uu, vv;
      if (Value(r) > std::numeric_limits<double>::epsilon()) {
        const T theta = atan(r);
        uu = theta * u / r;
        vv = theta * v / r;
      } else {
        uu = u;
        vv = v;
      }
      const T u2 = uu * uu, uv = uu * vv, v2 = vv * vv, r2 = u2 + v2, r4 = r2 * r2, r6 = r4 * r2,
              r8 = r6 * r2;
      const T radial = P(4) * r2 + P(5) * r4 + P(8) * r6 + P(9) * r8;
      const T du = uu * radial + C(2.0) * P(6) * uv + P(7) * (r2 + C(2.0) * u2) + P(10) * r2;
      const T dv = vv * radial + C(2.0) * P(7) * uv + P(6) * (r2 + C(2.0) * v2) + P(11) * r2;
      *x = uu + du;
      *y = vv + dv;
      *x = P(0) * *x + P(2);
      *y = P(1) * *y + P(3);
      return true;
    }
  }
  return false;
}

}  // namespace orc_cam
