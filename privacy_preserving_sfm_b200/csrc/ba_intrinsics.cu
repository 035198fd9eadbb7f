// ba_intrinsics.cu — intrinsics refinement in the bundle adjustment (BundleAdjuster::
// ParameterizeCameras, src/optim/bundle_adjustment.cc:490-528; camera_params is the fourth
// parameter block of the line cost functors, src/base/cost_functions.h:56-58, 130-141).
//
// Every default of the reference keeps the intrinsics constant (bundle_adjustment.h:58-64,
// controllers/incremental_mapper.h:81-83), so this is the option path, kept OUT of the hot kernels:
// with no variable camera nothing here runs and the pose / point path is untouched.  A camera with
// at least one variable parameter gets a reduced block of kIntrW = 12 columns behind the pose
// blocks (reduced index 6 NB + 12 c + a); constant parameters of a variable camera
// (SubsetParameterization) keep a zero Jacobian column and an identity row.  Per LM iteration:
//   ba_intr_jacobian_kernel   thread / observation   J_i = d r / d params (2 x 12) by forward-mode
//                             dual numbers through the SAME model expressions the reference
//                             differentiates with ceres::Jet; loss-weighted, Jacobi-scaled, masked
//   ba_intr_normal_kernel     thread / observation   U_ii (12x12 per camera), U_ic (12x6 per pose
//                             block), g_i: warp-aggregated atomics into small arrays
//   ba_intr_rows_kernel       S rows of the intrinsics blocks: direct terms + LM damping
//   ba_intr_schur_kernel      thread / point         Y_c = sum_e (J_i,e^T J_p,e) L^-T per variable
//                             camera seen by the point; S[c][c'] -= Y_c Y_c'^T,
//                             S[c][pose b] -= Y_c Z_f^T, rhs_c += Y_c h: atomics into the
//                             (small) intrinsics rows of the reduced system
// plus the intrinsics terms of back-substitution, update, Jacobi scaling and gradient norm.
// At most kMaxVarCams variable cameras per problem (a point sees few distinct cameras).
#include <cfloat>

#include "ba_kernels.h"
#include "camera_models.cuh"
#include "common.h"

namespace ppsfm {

namespace {

constexpr int IW = kIntrW;

// ---- forward-mode dual number over the 12 camera parameters ----------------------------------
struct DualP {
  double v;
  double d[IW];
};
__device__ __forceinline__ DualP dp_const(double c) {
  DualP r;
  r.v = c;
#pragma unroll
  for (int i = 0; i < IW; ++i) r.d[i] = 0.0;
  return r;
}
__device__ __forceinline__ DualP operator+(const DualP& x, const DualP& y) {
  DualP r;
  r.v = x.v + y.v;
#pragma unroll
  for (int i = 0; i < IW; ++i) r.d[i] = x.d[i] + y.d[i];
  return r;
}
__device__ __forceinline__ DualP operator-(const DualP& x, const DualP& y) {
  DualP r;
  r.v = x.v - y.v;
#pragma unroll
  for (int i = 0; i < IW; ++i) r.d[i] = x.d[i] - y.d[i];
  return r;
}
__device__ __forceinline__ DualP operator*(const DualP& x, const DualP& y) {
  DualP r;
  r.v = x.v * y.v;
#pragma unroll
  for (int i = 0; i < IW; ++i) r.d[i] = x.v * y.d[i] + x.d[i] * y.v;
  return r;
}
__device__ __forceinline__ DualP operator/(const DualP& x, const DualP& y) {
  DualP r;
  const double inv = 1.0 / y.v;
  r.v = x.v * inv;
#pragma unroll
  for (int i = 0; i < IW; ++i) r.d[i] = (x.d[i] - r.v * y.d[i]) * inv;
  return r;
}
__device__ __forceinline__ DualP dp_atan(const DualP& x) {
  DualP r;
  r.v = atan(x.v);
  const double s = 1.0 / (1.0 + x.v * x.v);
#pragma unroll
  for (int i = 0; i < IW; ++i) r.d[i] = x.d[i] * s;
  return r;
}
__device__ __forceinline__ DualP dp_tan(const DualP& x) {
  DualP r;
  r.v = tan(x.v);
  const double s = 1.0 + r.v * r.v;
#pragma unroll
  for (int i = 0; i < IW; ++i) r.d[i] = x.d[i] * s;
  return r;
}

// CameraModel::WorldToImage (src/base/camera_models.h:615-1481) with the parameters as dual
// numbers; (u, v) are plain values here (constants with respect to the intrinsics).
__device__ __noinline__ void world_to_image_dparams(int model, const DualP* p, double ud, double vd,
                                                    DualP& x, DualP& y) {
  auto C = [](double c) { return dp_const(c); };
  const DualP u = C(ud), v = C(vd);
  auto fisheye = [&](auto thetad_of_theta, DualP& du, DualP& dv) {
    const double r = sqrt(ud * ud + vd * vd);
    if (r > DBL_EPSILON) {
      const DualP theta = C(atan(r));
      const DualP thetad = thetad_of_theta(theta);
      du = u * thetad / C(r) - u;
      dv = v * thetad / C(r) - v;
    } else {
      du = C(0.0);
      dv = C(0.0);
    }
  };
  switch (model) {
    case 0:  // SIMPLE_PINHOLE f, cx, cy
      x = p[0] * u + p[1];
      y = p[0] * v + p[2];
      break;
    case 1:  // PINHOLE fx, fy, cx, cy
      x = p[0] * u + p[2];
      y = p[1] * v + p[3];
      break;
    case 2:    // SIMPLE_RADIAL f, cx, cy, k
    case 3: {  // RADIAL f, cx, cy, k1, k2
      const double r2 = ud * ud + vd * vd;
      DualP radial = p[3] * C(r2);
      if (model == 3) radial = radial + p[4] * C(r2 * r2);
      x = p[0] * (u + u * radial) + p[1];
      y = p[0] * (v + v * radial) + p[2];
      break;
    }
    case 4: {  // OPENCV fx, fy, cx, cy, k1, k2, p1, p2
      const double u2 = ud * ud, uv = ud * vd, v2 = vd * vd, r2 = u2 + v2;
      const DualP radial = p[4] * C(r2) + p[5] * C(r2 * r2);
      const DualP du = u * radial + C(2.0 * uv) * p[6] + p[7] * C(r2 + 2.0 * u2);
      const DualP dv = v * radial + C(2.0 * uv) * p[7] + p[6] * C(r2 + 2.0 * v2);
      x = p[0] * (u + du) + p[2];
      y = p[1] * (v + dv) + p[3];
      break;
    }
    case 5: {  // OPENCV_FISHEYE fx, fy, cx, cy, k1, k2, k3, k4
      DualP du, dv;
      fisheye([&](const DualP& theta) {
        const DualP t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
        return theta * (C(1.0) + p[4] * t2 + p[5] * t4 + p[6] * t6 + p[7] * t8);
      }, du, dv);
      x = p[0] * (u + du) + p[2];
      y = p[1] * (v + dv) + p[3];
      break;
    }
    case 6: {  // FULL_OPENCV fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, k5, k6
      const double u2 = ud * ud, uv = ud * vd, v2 = vd * vd, r2 = u2 + v2, r4 = r2 * r2, r6 = r4 * r2;
      const DualP radial = (C(1.0) + p[4] * C(r2) + p[5] * C(r4) + p[8] * C(r6)) /
                           (C(1.0) + p[9] * C(r2) + p[10] * C(r4) + p[11] * C(r6));
      const DualP du = u * radial + C(2.0 * uv) * p[6] + p[7] * C(r2 + 2.0 * u2) - u;
      const DualP dv = v * radial + C(2.0 * uv) * p[7] + p[6] * C(r2 + 2.0 * v2) - v;
      x = p[0] * (u + du) + p[2];
      y = p[1] * (v + dv) + p[3];
      break;
    }
    case 7: {  // FOV fx, fy, cx, cy, omega
      const DualP omega = p[4];
      const double radius2 = ud * ud + vd * vd;
      const DualP omega2 = omega * omega;
      DualP factor;
      if (omega2.v < 1e-4) {
        factor = (omega2 * C(radius2)) / C(3.0) - omega2 / C(12.0) + C(1.0);
      } else if (radius2 < 1e-4) {
        const DualP tho = dp_tan(omega / C(2.0));
        factor = (C(-2.0) * tho * (C(4.0 * radius2) * tho * tho - C(3.0))) / (C(3.0) * omega);
      } else {
        const double radius = sqrt(radius2);
        const DualP numerator = dp_atan(C(radius * 2.0) * dp_tan(omega / C(2.0)));
        factor = numerator / (C(radius) * omega);
      }
      x = p[0] * (u * factor) + p[2];
      y = p[1] * (v * factor) + p[3];
      break;
    }
    case 8:    // SIMPLE_RADIAL_FISHEYE f, cx, cy, k
    case 9: {  // RADIAL_FISHEYE f, cx, cy, k1, k2
      DualP du, dv;
      fisheye([&](const DualP& theta) {
        const DualP t2 = theta * theta;
        DualP poly = C(1.0) + p[3] * t2;
        if (model == 9) poly = poly + p[4] * (t2 * t2);
        return theta * poly;
      }, du, dv);
      x = p[0] * (u + du) + p[1];
      y = p[0] * (v + dv) + p[2];
      break;
    }
    default: {  // 10: THIN_PRISM_FISHEYE fx, fy, cx, cy, k1, k2, p1, p2, k3, k4, sx1, sy1
      const double r = sqrt(ud * ud + vd * vd);
      double uu = ud, vv = vd;
      if (r > DBL_EPSILON) {
        const double theta = atan(r);
        uu = theta * ud / r;
        vv = theta * vd / r;
      }
      const double u2 = uu * uu, uv = uu * vv, v2 = vv * vv, r2 = u2 + v2, r4 = r2 * r2, r6 = r4 * r2,
                   r8 = r6 * r2;
      const DualP radial = p[4] * C(r2) + p[5] * C(r4) + p[8] * C(r6) + p[9] * C(r8);
      const DualP du = C(uu) * radial + C(2.0 * uv) * p[6] + p[7] * C(r2 + 2.0 * u2) + p[10] * C(r2);
      const DualP dv = C(vv) * radial + C(2.0 * uv) * p[7] + p[6] * C(r2 + 2.0 * v2) + p[11] * C(r2);
      x = p[0] * (C(uu) + du) + p[2];
      y = p[1] * (C(vv) + dv) + p[3];
      break;
    }
  }
}

__device__ __forceinline__ void loss_rho1(BaLoss loss, double s, double& rho1) {
  if (loss.type == 0) {
    rho1 = 1.0;
    return;
  }
  const double b = loss.scale * loss.scale, sum = 1.0 + s / b;
  rho1 = loss.type == 1 ? fmax(DBL_MIN, 1.0 / sqrt(sum)) : fmax(DBL_MIN, 1.0 / sum);
}

// ---- J_i ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
ba_intr_jacobian_kernel(BaDev d, const double* __restrict__ q, const double* __restrict__ t,
                        const double* __restrict__ X, BaLoss loss) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= d.K) return;
  const int ci = d.obs_cam[k], cam = d.img_cam[ci];
  const int ib = d.cam_intr_block[cam];
  double* out = d.Ji + (size_t)k * 2 * IW;
  if (ib < 0) {
#pragma unroll
    for (int i = 0; i < 2 * IW; ++i) out[i] = 0.0;
    return;
  }
  const int pi = d.obs_pt[k];
  const double a = d.obs_line[k], b = d.obs_line[d.K + k], c = d.obs_line[2 * d.K + k];
  const double qw = q[4 * (size_t)ci], qx = q[4 * (size_t)ci + 1], qy = q[4 * (size_t)ci + 2],
               qz = q[4 * (size_t)ci + 3];
  const double X0 = X[3 * (size_t)pi], X1 = X[3 * (size_t)pi + 1], X2 = X[3 * (size_t)pi + 2];
  // ceres::UnitQuaternionRotatePoint, as in ba_linearize_kernel
  const double t2 = qw * qx, t3 = qw * qy, t4 = qw * qz, t5 = -qx * qx, t6 = qx * qy;
  const double t7 = qx * qz, t8 = -qy * qy, t9 = qy * qz, t1 = -qz * qz;
  const double p0 = 2.0 * ((t8 + t1) * X0 + (t6 - t4) * X1 + (t3 + t7) * X2) + X0 + t[3 * (size_t)ci];
  const double p1 = 2.0 * ((t4 + t6) * X0 + (t5 + t1) * X1 + (t9 - t2) * X2) + X1 + t[3 * (size_t)ci + 1];
  const double p2 = 2.0 * ((t7 - t3) * X0 + (t2 + t9) * X1 + (t5 + t8) * X2) + X2 + t[3 * (size_t)ci + 2];
  const double iz = 1.0 / p2, u = p0 * iz, v = p1 * iz;
  const double alpha = a * u + b * v + c;
  const double lu = u - alpha * a, lv = v - alpha * b;
  const int model = d.cam_model[cam];
  DualP prm[IW];
#pragma unroll
  for (int i = 0; i < IW; ++i) {
    prm[i] = dp_const(d.cam_params[12 * (size_t)cam + i]);
    prm[i].d[i] = 1.0;
  }
  DualP x1, y1, x2, y2;
  world_to_image_dparams(model, prm, u, v, x1, y1);
  world_to_image_dparams(model, prm, lu, lv, x2, y2);
  const double r0 = x1.v - x2.v, r1 = y1.v - y2.v;
  double rho1;
  loss_rho1(loss, r0 * r0 + r1 * r1, rho1);
  const double sr = sqrt(rho1);
  const unsigned mask = d.intr_mask[ib];
#pragma unroll
  for (int i = 0; i < IW; ++i) {
    const double s = ((mask >> i) & 1u) ? sr * d.intr_scale[IW * (size_t)ib + i] : 0.0;
    out[i] = s * (x1.d[i] - x2.d[i]);
    out[IW + i] = s * (y1.d[i] - y2.d[i]);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums `v` over the CTA; result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* smem /* >= 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = (lane < nw) ? smem[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr),
            (unsigned long long)__double_as_longlong(v));
}

// ---- U_ii, U_ic, g_i -----------------------------------------------------------------------------
// Observations are point-major and cameras are few, so a warp holds very few distinct intrinsics
// blocks: for every distinct block (uniform loop) the warp sums each value and one lane adds it.
__global__ void __launch_bounds__(256) ba_intr_normal_kernel(BaDev d) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int ib = -1, b = -1;
  double ji[2 * IW], r0 = 0.0, r1 = 0.0;
#pragma unroll
  for (int i = 0; i < 2 * IW; ++i) ji[i] = 0.0;
  if (k < d.K) {
    const int ci = d.obs_cam[k];
    ib = d.cam_intr_block[d.img_cam[ci]];
    b = d.cam_block[ci];
    if (ib >= 0) {
#pragma unroll
      for (int i = 0; i < 2 * IW; ++i) ji[i] = d.Ji[(size_t)k * 2 * IW + i];
      r0 = d.J[ba_jidx(0, k)];
      r1 = d.J[ba_jidx(1, k)];
      if (b >= 0) {  // U_ic: the pose block differs from lane to lane -> plain atomics
        double jc[12];
#pragma unroll
        for (int a = 0; a < 12; ++a) jc[a] = d.J[ba_jidx(2 + a, k)];
#pragma unroll
        for (int a = 0; a < IW; ++a)
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const double v = ji[a] * jc[c] + ji[IW + a] * jc[6 + c];
            if (v != 0.0) atomicAdd(&d.Uic[(size_t)b * IW * 6 + 6 * a + c], v);
          }
      }
    }
  }
  unsigned remaining = 0xffffffffu;
  const unsigned mine = __match_any_sync(0xffffffffu, ib);
  while (remaining) {  // (uniform: every lane sees the same `remaining`)
    const int leader = __ffs(remaining) - 1;
    const unsigned gmask = __shfl_sync(0xffffffffu, mine, leader);
    const int gib = __shfl_sync(0xffffffffu, ib, leader);
    remaining &= ~gmask;
    if (gib < 0) continue;
    const bool in = (gmask >> lane) & 1u;
    double* uii = d.Uii + (size_t)gib * IW * IW;
    for (int a = 0; a < IW; ++a) {
      for (int c = 0; c <= a; ++c) {
        const double v = warp_sum(in ? ji[a] * ji[c] + ji[IW + a] * ji[IW + c] : 0.0);
        if (lane == leader && v != 0.0) {
          atomicAdd(&uii[IW * a + c], v);
          if (c != a) atomicAdd(&uii[IW * c + a], v);
        }
      }
      const double g = warp_sum(in ? ji[a] * r0 + ji[IW + a] * r1 : 0.0);
      if (lane == leader && g != 0.0) atomicAdd(&d.gi[(size_t)gib * IW + a], g);
    }
  }
}

__global__ void ba_intr_scales_kernel(BaDev d, int jacobi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= IW * d.NCv) return;
  const int ib = i / IW, a = i - IW * ib;
  d.intr_scale[i] = jacobi ? 1.0 / (1.0 + sqrt(d.Uii[(size_t)ib * IW * IW + (IW + 1) * a])) : 1.0;
}

// ---- S rows of the intrinsics blocks: direct terms ------------------------------------------------
__global__ void ba_intr_rows_kernel(BaDev d, double radius, double min_diag, double max_diag,
                                    int include_cam) {
  const int ioff = 6 * d.NB;
  const int row = blockIdx.x;  // 0 .. IW * NCv (last = right-hand-side entries)
  const int nrows = IW * d.NCv;
  if (row == nrows) {  // rhs entries of the intrinsics columns
    for (int j = threadIdx.x; j < nrows; j += blockDim.x) {
      const int ib = j / IW, a = j - IW * ib;
      const bool on = (d.intr_mask[ib] >> a) & 1u;
      d.S[(size_t)d.n * d.ld + ioff + j] = (on && include_cam) ? -d.gi[j] : 0.0;
    }
    return;
  }
  const int ib = row / IW, a = row - IW * ib;
  const bool on_a = (d.intr_mask[ib] >> a) & 1u;
  double* Srow = d.S + (size_t)(ioff + row) * d.ld;
  for (int j = threadIdx.x; j <= ioff + row; j += blockDim.x) {
    double v = 0.0;
    if (j < ioff) {  // coupling with pose block j / 6 (only the blocks of this camera's images)
      const int b = j / 6, c = j - 6 * b;
      if (include_cam && on_a && d.cam_intr_block[d.img_cam[d.block_img[b]]] == ib)
        v = d.Uic[(size_t)b * IW * 6 + 6 * a + c];
    } else {
      const int jb = (j - ioff) / IW, c = (j - ioff) - IW * jb;
      if (jb == ib) {
        const bool on_c = (d.intr_mask[ib] >> c) & 1u;
        if (on_a && on_c) {
          if (include_cam) {
            const double u = d.Uii[(size_t)ib * IW * IW + IW * a + c];
            v = u;
            if (a == c) v += fmin(fmax(u, min_diag), max_diag) / radius;
          }
        } else if (a == c && include_cam) {
          v = 1.0;  // constant parameter of a variable camera: identity row
        }
      }
    }
    Srow[j] = v;
  }
}

// ---- Schur terms of the intrinsics rows -----------------------------------------------------------
constexpr int kSlots = 8;  // distinct variable cameras among the views of one point
__global__ void __launch_bounds__(128) ba_intr_schur_kernel(BaDev d, int* __restrict__ overflow) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.P || !d.pt_var[p]) return;
  const int64_t k0 = d.pt_start[p], k1 = d.pt_start[p + 1];
  const double* lv = d.Lv + 9 * (size_t)p;
  const double m00 = lv[0], m10 = lv[1], m11 = lv[2], m20 = lv[3], m21 = lv[4], m22 = lv[5];
  const double h0 = lv[6], h1 = lv[7], h2 = lv[8];
  int slot_cam[kSlots];
  double Y[kSlots][IW][3];
  int ns = 0;
  for (int64_t k = k0; k < k1; ++k) {
    const int ib = d.cam_intr_block[d.img_cam[d.obs_cam[k]]];
    if (ib < 0) continue;
    int s = 0;
    while (s < ns && slot_cam[s] != ib) ++s;
    if (s == ns) {
      if (ns == kSlots) {
        *overflow = 1;
        return;
      }
      slot_cam[ns++] = ib;
      for (int a = 0; a < IW; ++a) Y[s][a][0] = Y[s][a][1] = Y[s][a][2] = 0.0;
    }
    double jp[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) jp[c] = d.J[ba_jidx(14 + c, k)];
    const double* ji = d.Ji + (size_t)k * 2 * IW;
    for (int a = 0; a < IW; ++a) {
      const double j0 = ji[a], j1 = ji[IW + a];
      const double w0 = j0 * jp[0] + j1 * jp[3], w1 = j0 * jp[1] + j1 * jp[4],
                   w2 = j0 * jp[2] + j1 * jp[5];
      Y[s][a][0] += w0 * m00;                          // Z = W M^T, as ba_zbuild_kernel
      Y[s][a][1] += w0 * m10 + w1 * m11;
      Y[s][a][2] += w0 * m20 + w1 * m21 + w2 * m22;
    }
  }
  if (ns == 0) return;
  const int ioff = 6 * d.NB;
  for (int s = 0; s < ns; ++s) {
    const int ra = ioff + IW * slot_cam[s];
    // right-hand side: + Y h
    for (int a = 0; a < IW; ++a) {
      const double v = Y[s][a][0] * h0 + Y[s][a][1] * h1 + Y[s][a][2] * h2;
      if (v != 0.0) atomicAdd(&d.S[(size_t)d.n * d.ld + ra + a], v);
    }
    // (intrinsics, intrinsics): - Y_c Y_c'^T, lower triangle
    for (int s2 = 0; s2 < ns; ++s2) {
      if (slot_cam[s2] > slot_cam[s]) continue;
      const int rc = ioff + IW * slot_cam[s2];
      for (int a = 0; a < IW; ++a)
        for (int c = 0; c < (s2 == s ? a + 1 : IW); ++c) {
          const double v = Y[s][a][0] * Y[s2][c][0] + Y[s][a][1] * Y[s2][c][1] + Y[s][a][2] * Y[s2][c][2];
          if (v != 0.0) atomicAdd(&d.S[(size_t)(ra + a) * d.ld + rc + c], -v);
        }
    }
  }
  // (intrinsics, pose): - Y_c Z_f^T for every view f of the point with a variable pose
  for (int64_t k = k0; k < k1; ++k) {
    const int b = d.cam_block[d.obs_cam[k]];
    if (b < 0) continue;
    const double* rec = d.Zrec + (size_t)k * 24;  // rows [Z_a0 Z_a1 Z_a2 z_a], a = 0..5
    for (int s = 0; s < ns; ++s) {
      const int ra = ioff + IW * slot_cam[s];
      for (int a = 0; a < IW; ++a)
        for (int c = 0; c < 6; ++c) {
          const double v = Y[s][a][0] * rec[4 * c] + Y[s][a][1] * rec[4 * c + 1] + Y[s][a][2] * rec[4 * c + 2];
          if (v != 0.0) atomicAdd(&d.S[(size_t)(ra + a) * d.ld + 6 * b + c], -v);
        }
    }
  }
}

// ---- update, gradient norm ------------------------------------------------------------------------
__global__ void ba_intr_update_kernel(BaDev d, double* __restrict__ cam_params_n,
                                      double* __restrict__ img_params_n, int count_norms) {
  __shared__ double red[32];
  double step_sq = 0.0, x_sq = 0.0;
  const int ioff = 6 * d.NB;
  for (int i = threadIdx.x; i < 12 * d.num_cameras; i += blockDim.x) {
    const int cam = i / 12, a = i - 12 * cam;
    const int ib = d.cam_intr_block[cam];
    const double x = d.cam_params[i];
    double dl = 0.0;
    if (ib >= 0) {
      if ((d.intr_mask[ib] >> a) & 1u) dl = d.dc[ioff + IW * ib + a] * d.intr_scale[IW * (size_t)ib + a];
      if (a < d.cam_nparams[cam]) {
        step_sq += dl * dl;
        x_sq += x * x;
      }
    }
    cam_params_n[i] = x + dl;
  }
  const double ts = block_sum(step_sq, red);
  const double tx = block_sum(x_sq, red);
  if (threadIdx.x == 0 && count_norms) {  // on top of the pose / point parts (same stream, earlier)
    d.scalars[kStepSq] += ts;
    d.scalars[kXSq] += tx;
  }
  __syncthreads();  // cam_params_n complete (one CTA)
  for (int i = threadIdx.x; i < 12 * d.C; i += blockDim.x) {
    const int img = i / 12, a = i - 12 * img;
    img_params_n[i] = cam_params_n[12 * (size_t)d.img_cam[img] + a];
  }
}

// u += J_i d_i (kept for the model-cost pass) and acc_p += J_p^T (J_i d_i); runs between
// ba_backsub_accum_kernel and ba_point_step_kernel
__global__ void __launch_bounds__(256) ba_intr_backsub_kernel(BaDev d) {
  const int64_t K = d.K;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int ib = d.cam_intr_block[d.img_cam[d.obs_cam[k]]];
  if (ib < 0) return;
  const double* ji = d.Ji + (size_t)k * 2 * IW;
  const double* di = d.dc + 6 * (size_t)d.NB + (size_t)IW * ib;
  double u0 = 0.0, u1 = 0.0;
#pragma unroll
  for (int a = 0; a < IW; ++a) {
    u0 += ji[a] * di[a];
    u1 += ji[IW + a] * di[a];
  }
  d.u[k] += u0;
  d.u[K + k] += u1;
  const int p = d.obs_pt[k];
  const int P = d.P;
#pragma unroll
  for (int c = 0; c < 3; ++c)
    atomicAdd(&d.dp[(size_t)c * P + p],
              d.J[ba_jidx(14 + c, k)] * u0 + d.J[ba_jidx(17 + c, k)] * u1);
}

__global__ void ba_intr_gradient_kernel(BaDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= IW * d.NCv) return;
  // SubsetParameterization: Plus(x, delta) = x + delta -> |x - Plus(x, -g)| = |g|
  const double g = fabs(d.gi[i] / d.intr_scale[i]);
  if (g > 0.0) atomic_max_nonneg(&d.scalars[kGradMax], g);
}

}  // namespace

// ================================================================================================
int launch_intr_jacobian(const BaDev& d, const double* q, const double* t, const double* X,
                         BaLoss loss, cudaStream_t s) {
  if (d.NCv == 0 || d.K == 0) return 0;
  ba_intr_jacobian_kernel<<<(unsigned)((d.K + 127) / 128), 128, 0, s>>>(d, q, t, X, loss);
  return 1;
}

int launch_intr_normal(const BaDev& d, cudaStream_t s) {
  if (d.NCv == 0) return 0;
  cudaMemsetAsync(d.Uii, 0, sizeof(double) * intr_normal_doubles(d.NB, d.NCv), s);  // Uii|Uic|gi
  if (d.K > 0) ba_intr_normal_kernel<<<(unsigned)((d.K + 255) / 256), 256, 0, s>>>(d);
  return 1;
}

int launch_intr_scales(const BaDev& d, bool jacobi, cudaStream_t s) {
  if (d.NCv == 0) return 0;
  ba_intr_scales_kernel<<<(IW * d.NCv + 127) / 128, 128, 0, s>>>(d, jacobi ? 1 : 0);
  return 1;
}

int launch_intr_reduced_rows(const BaDev& d, double radius, double min_diag, double max_diag,
                             bool include_camera_terms, int* overflow, cudaStream_t s) {
  if (d.NCv == 0) return 0;
  ba_intr_rows_kernel<<<IW * d.NCv + 1, 256, 0, s>>>(d, radius, min_diag, max_diag,
                                                     include_camera_terms ? 1 : 0);
  if (d.P > 0) ba_intr_schur_kernel<<<(d.P + 127) / 128, 128, 0, s>>>(d, overflow);
  return 2;
}

int launch_intr_backsub(const BaDev& d, cudaStream_t s) {
  if (d.NCv == 0 || d.K == 0) return 0;
  ba_intr_backsub_kernel<<<(unsigned)((d.K + 255) / 256), 256, 0, s>>>(d);
  return 1;
}

int launch_intr_update(const BaDev& d, double* cam_params_n, double* img_params_n,
                       bool count_norms, cudaStream_t s) {
  if (d.NCv == 0) return 0;
  ba_intr_update_kernel<<<1, 128, 0, s>>>(d, cam_params_n, img_params_n, count_norms ? 1 : 0);
  return 1;
}

int launch_intr_gradient(const BaDev& d, cudaStream_t s) {
  if (d.NCv == 0) return 0;
  ba_intr_gradient_kernel<<<(IW * d.NCv + 127) / 128, 128, 0, s>>>(d);
  return 1;
}

}  // namespace ppsfm
