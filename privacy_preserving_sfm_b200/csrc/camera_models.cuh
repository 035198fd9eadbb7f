// camera_models.cuh — CameraModel::WorldToImage of the supported COLMAP camera models
// (src/base/camera_models.h:615-904) as a device function, shared by the bundle adjustment and the
// observation filters.
#pragma once

namespace ppsfm {

// CameraModel::WorldToImage (src/base/camera_models.h) and its 2x2 Jacobian d(x,y)/d(u,v).
template <bool JAC>
__device__ __forceinline__ void world_to_image(int model, const double* __restrict__ p, double u,
                                               double v, double& x, double& y, double& xu,
                                               double& xv, double& yu, double& yv) {
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
