// camera_models.cuh — CameraModel::WorldToImage of the supported COLMAP camera models
// (src/base/camera_models.h:615-904) as a device function, shared by the bundle adjustment and the
// observation filters.
#pragma once
#include <cfloat>

namespace ppsfm {

// ------------------------------------------------------------------------------------------
// The six remaining COLMAP models (src/base/camera_models.h: OPENCV_FISHEYE :929-986,
// FULL_OPENCV :1024-1080, FOV :1103-1166, SIMPLE_RADIAL_FISHEYE :1238-1290, RADIAL_FISHEYE
// :1318-1366, THIN_PRISM_FISHEYE :1405-1481).  Each model is written ONCE, expression by
// expression as the reference, over a scalar type S that is either double or a two-direction dual
// number (value, d/du, d/dv): the 2x2 Jacobian the bundle adjustment needs is then the forward-mode
// derivative of exactly the expressions the reference differentiates with ceres::Jet.  Not
// inlined: the pinhole / radial / OpenCV path of the callers keeps its register budget.
// ------------------------------------------------------------------------------------------
struct Dual2 {
  double v, a, b;  // value, d/du, d/dv
};
__device__ __forceinline__ Dual2 operator+(Dual2 x, Dual2 y) { return {x.v + y.v, x.a + y.a, x.b + y.b}; }
__device__ __forceinline__ Dual2 operator-(Dual2 x, Dual2 y) { return {x.v - y.v, x.a - y.a, x.b - y.b}; }
__device__ __forceinline__ Dual2 operator*(Dual2 x, Dual2 y) {
  return {x.v * y.v, x.v * y.a + x.a * y.v, x.v * y.b + x.b * y.v};
}
__device__ __forceinline__ Dual2 operator/(Dual2 x, Dual2 y) {
  const double inv = 1.0 / y.v, q = x.v * inv;
  return {q, (x.a - q * y.a) * inv, (x.b - q * y.b) * inv};
}
__device__ __forceinline__ Dual2 cm_sqrt(Dual2 x) {
  const double r = sqrt(x.v), d = 1.0 / (2.0 * r);
  return {r, x.a * d, x.b * d};
}
__device__ __forceinline__ Dual2 cm_atan(Dual2 x) {
  const double d = 1.0 / (1.0 + x.v * x.v);
  return {atan(x.v), x.a * d, x.b * d};
}
__device__ __forceinline__ double cm_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double cm_atan(double x) { return atan(x); }
__device__ __forceinline__ double cm_value(double x) { return x; }
__device__ __forceinline__ double cm_value(Dual2 x) { return x.v; }
template <typename S> __device__ __forceinline__ S cm_const(double c);
template <> __device__ __forceinline__ double cm_const<double>(double c) { return c; }
template <> __device__ __forceinline__ Dual2 cm_const<Dual2>(double c) { return {c, 0.0, 0.0}; }

template <typename S, typename F>
__device__ __forceinline__ void cm_fisheye(S u, S v, F thetad_of_theta, S& du, S& dv) {
  const S r = cm_sqrt(u * u + v * v);
  if (cm_value(r) > DBL_EPSILON) {
    const S theta = cm_atan(r);
    const S thetad = thetad_of_theta(theta);
    du = u * thetad / r - u;
    dv = v * thetad / r - v;
  } else {
    du = cm_const<S>(0.0);
    dv = cm_const<S>(0.0);
  }
}

template <typename S>
__device__ void world_to_image_ext_t(int model, const double* __restrict__ p, S u, S v, S& x, S& y) {
  auto C = [](double c) { return cm_const<S>(c); };
  auto P = [&](int k) { return cm_const<S>(p[k]); };
  switch (model) {
    case 5: {  // OPENCV_FISHEYE fx, fy, cx, cy, k1, k2, k3, k4
      S du, dv;
      cm_fisheye(u, v, [&](S theta) {
        const S theta2 = theta * theta;
        const S theta4 = theta2 * theta2;
        const S theta6 = theta4 * theta2;
        const S theta8 = theta4 * theta4;
        return theta * (C(1.0) + P(4) * theta2 + P(5) * theta4 + P(6) * theta6 + P(7) * theta8);
      }, du, dv);
      x = u + du;
      y = v + dv;
      x = P(0) * x + P(2);
      y = P(1) * y + P(3);
      break;
    }
    case 6: {  // FULL_OPENCV fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, k5, k6
      const S u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2, r4 = r2 * r2, r6 = r4 * r2;
      const S radial = (C(1.0) + P(4) * r2 + P(5) * r4 + P(8) * r6) /
                       (C(1.0) + P(9) * r2 + P(10) * r4 + P(11) * r6);
      const S du = u * radial + C(2.0) * P(6) * uv + P(7) * (r2 + C(2.0) * u2) - u;
      const S dv = v * radial + C(2.0) * P(7) * uv + P(6) * (r2 + C(2.0) * v2) - v;
      x = u + du;
      y = v + dv;
      x = P(0) * x + P(2);
      y = P(1) * y + P(3);
      break;
    }
    case 7: {  // FOV fx, fy, cx, cy, omega  (omega is a constant here: plain double branches)
      const double omega = p[4], omega2 = omega * omega;
      const S radius2 = u * u + v * v;
      S factor;
      if (omega2 < 1e-4) {
        factor = (C(omega2) * radius2) / C(3.0) - C(omega2 / 12.0) + C(1.0);
      } else if (cm_value(radius2) < 1e-4) {
        const double tan_half_omega = tan(omega / 2.0);
        factor = (C(-2.0 * tan_half_omega) *
                  (C(4.0) * radius2 * C(tan_half_omega) * C(tan_half_omega) - C(3.0))) /
                 C(3.0 * omega);
      } else {
        const S radius = cm_sqrt(radius2);
        const S numerator = cm_atan(radius * C(2.0) * C(tan(omega / 2.0)));
        factor = numerator / (radius * C(omega));
      }
      x = u * factor;
      y = v * factor;
      x = P(0) * x + P(2);
      y = P(1) * y + P(3);
      break;
    }
    case 8:    // SIMPLE_RADIAL_FISHEYE f, cx, cy, k
    case 9: {  // RADIAL_FISHEYE f, cx, cy, k1, k2
      S du, dv;
      cm_fisheye(u, v, [&](S theta) {
        const S theta2 = theta * theta;
        if (model == 8) return theta * (C(1.0) + P(3) * theta2);
        const S theta4 = theta2 * theta2;
        return theta * (C(1.0) + P(3) * theta2 + P(4) * theta4);
      }, du, dv);
      x = u + du;
      y = v + dv;
      x = P(0) * x + P(1);
      y = P(0) * y + P(2);
      break;
    }
    default: {  // 10: THIN_PRISM_FISHEYE fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, sx1, sy1
      const S r = cm_sqrt(u * u + v * v);
      S uu, vv;
      if (cm_value(r) > DBL_EPSILON) {
        const S theta = cm_atan(r);
        uu = theta * u / r;
        vv = theta * v / r;
      } else {
        uu = u;
        vv = v;
      }
      const S u2 = uu * uu, uv = uu * vv, v2 = vv * vv, r2 = u2 + v2, r4 = r2 * r2, r6 = r4 * r2,
              r8 = r6 * r2;
      const S radial = P(4) * r2 + P(5) * r4 + P(8) * r6 + P(9) * r8;
      const S du = uu * radial + C(2.0) * P(6) * uv + P(7) * (r2 + C(2.0) * u2) + P(10) * r2;
      const S dv = vv * radial + C(2.0) * P(7) * uv + P(6) * (r2 + C(2.0) * v2) + P(11) * r2;
      x = uu + du;
      y = vv + dv;
      x = P(0) * x + P(2);
      y = P(1) * y + P(3);
      break;
    }
  }
}

// Result by value: the caller's own variables never have their address taken, so its fast path
// (pinhole / radial / OpenCV, inlined) keeps them in registers.
struct WorldToImageResult {
  double x, y, xu, xv, yu, yv;
};
template <bool JAC>
__device__ __noinline__ WorldToImageResult world_to_image_ext(int model,
                                                              const double* __restrict__ p,
                                                              double u, double v) {
  WorldToImageResult r;
  if (JAC) {
    Dual2 X, Y;
    world_to_image_ext_t<Dual2>(model, p, Dual2{u, 1.0, 0.0}, Dual2{v, 0.0, 1.0}, X, Y);
    r.x = X.v; r.xu = X.a; r.xv = X.b;
    r.y = Y.v; r.yu = Y.a; r.yv = Y.b;
  } else {
    double x, y;
    world_to_image_ext_t<double>(model, p, u, v, x, y);
    r.x = x; r.y = y;
    r.xu = r.xv = r.yu = r.yv = 0.0;
  }
  return r;
}

// CameraModel::WorldToImage (src/base/camera_models.h) and its 2x2 Jacobian d(x,y)/d(u,v).
// EXT = false: models 0..4 only (the caller has dealt with the others)
template <bool JAC, bool EXT = true>
__device__ __forceinline__ void world_to_image(int model, const double* __restrict__ p, double u,
                                               double v, double& x, double& y, double& xu,
                                               double& xv, double& yu, double& yv) {
  if (EXT && model >= 5) {  // fisheye / FOV / full-OpenCV / thin-prism: out of line
    const WorldToImageResult r = world_to_image_ext<JAC>(model, p, u, v);
    x = r.x; y = r.y;
    if (JAC) { xu = r.xu; xv = r.xv; yu = r.yu; yv = r.yv; }
    return;
  }
  switch (model) {
    case 0: {  // SIMPLE_PINHOLE f, cx, cy
      x = p[0] * u + p[1];
      y = p[0] * v + p[2];
      if (JAC) { xu = p[0]; xv = 0.0; yu = 0.0; yv = p[0]; }
      break;
    }
    case 1: {  // PINHOLE fx, fy, cx, cy
      x = p[0] * u + p[2];
      y = p[1] * v + p[3];
      if (JAC) { xu = p[0]; xv = 0.0; yu = 0.0; yv = p[1]; }
      break;
    }
    case 2:    // SIMPLE_RADIAL f, cx, cy, k
    case 3: {  // RADIAL f, cx, cy, k1, k2
      const double k1 = p[3], k2 = (model == 3) ? p[4] : 0.0;
      const double u2 = u * u, v2 = v * v, r2 = u2 + v2;
      const double radial = k1 * r2 + k2 * r2 * r2;
      x = p[0] * (u + u * radial) + p[1];
      y = p[0] * (v + v * radial) + p[2];
      if (JAC) {
        const double g = 2.0 * (k1 + 2.0 * k2 * r2);  // d radial / d(r2) * 2
        xu = p[0] * (1.0 + radial + u * u * g);
        xv = p[0] * (u * v * g);
        yu = xv;
        yv = p[0] * (1.0 + radial + v * v * g);
      }
      break;
    }
    default: {  // 4: OPENCV fx, fy, cx, cy, k1, k2, p1, p2
      const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
      const double u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2;
      const double radial = k1 * r2 + k2 * r2 * r2;
      const double du = u * radial + 2.0 * p1 * uv + p2 * (r2 + 2.0 * u2);
      const double dv = v * radial + 2.0 * p2 * uv + p1 * (r2 + 2.0 * v2);
      x = p[0] * (u + du) + p[2];
      y = p[1] * (v + dv) + p[3];
      if (JAC) {
        const double g = 2.0 * (k1 + 2.0 * k2 * r2);
        const double duu = radial + u2 * g + 2.0 * p1 * v + 6.0 * p2 * u;
        const double duv = uv * g + 2.0 * p1 * u + 2.0 * p2 * v;
        const double dvu = uv * g + 2.0 * p2 * v + 2.0 * p1 * u;
        const double dvv = radial + v2 * g + 2.0 * p2 * u + 6.0 * p1 * v;
        xu = p[0] * (1.0 + duu);
        xv = p[0] * duv;
        yu = p[1] * dvu;
        yv = p[1] * (1.0 + dvv);
      }
      break;
    }
  }
}


}  // namespace ppsfm
