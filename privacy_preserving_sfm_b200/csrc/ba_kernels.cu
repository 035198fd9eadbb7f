// ba_kernels.cu — sm_100a kernels of the line-reprojection bundle adjustment.
//
// Replaces what the reference delegates to Ceres (src/optim/bundle_adjustment.cc:306):
//   ba_linearize_kernel      residual of BundleAdjustmentLineCostFunction
//                            (src/base/cost_functions.h:62-100) with ANALYTIC 2x6 / 2x3 Jacobian
//                            blocks in the tangent space of ceres::QuaternionParameterization,
//                            robust-loss correction and Jacobi column scaling fused in
//   ba_point_normal_kernel   V_p = sum J_p^T J_p, g_p          (thread per observation, segmented)
//   ba_camera_normal_kernel  U_c = sum J_c^T J_c, g_c          (4 CTAs per camera, shuffles)
//   (reduced camera system: ba_schur.cu)
//   ba_backsub_kernel        dp = -V^-1 (g_p + W^T dc), model cost change, candidate points
//   ba_camera_update_kernel  candidate poses via QuaternionParameterization::Plus
// Observation data is stored SoA over the (point-major) observation index so that every warp
// access is a coalesced 256-byte line.
#include <cfloat>
#include <cstdint>

#include "ba_kernels.h"
#include "camera_models.cuh"
#include "common.h"

namespace ppsfm {

namespace {

constexpr int kThreads = 256;

// Linearisation storage (see BaDev::J): field f of observation k.
#define JR(row, k) d.J[ba_jidx((row), (k))]
#define JC(i, k) d.J[ba_jidx(2 + (i), (k))]
#define JP(i, k) d.J[ba_jidx(14 + (i), (k))]

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// J store: plain (write-back, stays in L2) or streaming (evict-first)
__device__ __forceinline__ void jstore(double* p, double v, int streaming) {
  if (streaming) __stcs(p, v); else *p = v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Sums `v` over the CTA (blockDim.x multiple of 32, <= 1024); result valid in thread 0.
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

// ceres loss functions (TrivialLoss, SoftLOneLoss, CauchyLoss): rho(s) and rho'(s).
// rho'' <= 0 for all three, so the Triggs corrector reduces to scaling by sqrt(rho').
__device__ __forceinline__ void eval_loss(BaLoss loss, double s, double& rho0, double& rho1) {
  if (loss.type == 0) {
    rho0 = s;
    rho1 = 1.0;
    return;
  }
  const double b = loss.scale * loss.scale, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  if (loss.type == 1) {
    const double tmp = sqrt(sum);
    rho0 = 2.0 * b * (tmp - 1.0);
    rho1 = fmax(DBL_MIN, 1.0 / tmp);
  } else {
    const double inv = 1.0 / sum;
    rho0 = b * log(sum);
    rho1 = fmax(DBL_MIN, inv);
  }
}

// ------------------------------------------------------------------------------------------
// Linearisation: one thread per observation.
// Algorithmic HBM traffic (materialised variant): read line 24 + cam 4 + pt 4 +
// point 24 B, write r 16 + J_c 96 + J_p 48 B = 216 B per observation.
// ------------------------------------------------------------------------------------------
// Per-image record staged in shared memory by the persistent variant (doubles):
//   [0..3] q, [4..6] t, [7..12] Jacobi scale x tangent mask (0 = fixed dim / constant pose).
// Odd stride: lanes that read different images mostly hit different banks.  Intrinsics are
// staged per CAMERA (8 doubles = OPENCV, the largest supported model) next to it.
constexpr int kCamRec = 13;
constexpr int kIntrRec = 13;  // 12 parameters + model id (stored as a double)

// EXT: the problem holds a camera of the fisheye / FOV / full-OpenCV / thin-prism family; those
// models are evaluated out of line, and the instance without them is the unchanged hot kernel.
template <bool JAC, bool SMEM, bool EXT = false>
__global__ void __launch_bounds__(kThreads, 2)  // ~100 live registers: 2 CTAs / SM, no spills
ba_linearize_kernel(BaDev d, const double* __restrict__ q, const double* __restrict__ t,
                    const double* __restrict__ X, BaLoss loss, double* __restrict__ partials,
                    int num_chunks, int store_mode) {
  __shared__ double red[32];
  // [C][kCamRec] image records, [num_cameras][kIntrRec] intrinsics, [C] camera index of an image
  extern __shared__ __align__(16) double cam_tab[];
  double* intr_tab = cam_tab + (size_t)d.C * kCamRec;
  int* tab_cam = reinterpret_cast<int*>(intr_tab + (size_t)d.num_cameras * kIntrRec);
  const int64_t K = d.K;
  // Software pipeline: the streaming inputs (image / point index, line) of this CTA's NEXT chunk
  // are loaded into registers while the current chunk is computed, and the chunk after that is
  // pulled into L2 — the kernel writes 3x what it reads, and demand reads that queue behind the
  // write stream in DRAM would otherwise stall every warp of the SM at the same time.
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int ci_n = 0, pi_n = 0;
  double a_n = 0.0, b_n = 0.0, c_n = 0.0;
  if (k < K) {
    ci_n = d.obs_cam[k]; pi_n = d.obs_pt[k];
    a_n = d.obs_line[k]; b_n = d.obs_line[K + k]; c_n = d.obs_line[2 * K + k];
  }
  if (SMEM) {
    for (int idx = threadIdx.x; idx < d.C * kCamRec; idx += kThreads) {
      const int ci = idx / kCamRec, f = idx - ci * kCamRec;
      double v;
      if (f < 4) v = q[4 * (size_t)ci + f];
      else if (f < 7) v = t[3 * (size_t)ci + (f - 4)];
      else {
        const int blk = d.cam_block[ci];
        const int j = f - 7;
        v = (JAC && blk >= 0 && ((d.cam_mask[ci] >> j) & 1)) ? d.cam_scale[6 * (size_t)blk + j] : 0.0;
      }
      cam_tab[idx] = v;
    }
    for (int idx = threadIdx.x; idx < d.num_cameras * kIntrRec; idx += kThreads) {
      const int cam = idx / kIntrRec, f = idx - cam * kIntrRec;
      intr_tab[idx] = (f < 12) ? d.cam_params[12 * (size_t)cam + f] : (double)d.cam_model[cam];
    }
    for (int ci = threadIdx.x; ci < d.C; ci += kThreads) tab_cam[ci] = d.img_cam[ci];
    __syncthreads();
  }
  double cost = 0.0;
  for (int chunk = blockIdx.x; chunk < num_chunks; chunk += gridDim.x) {
    const int64_t kc = k;
    const bool valid = kc < K;
    const int ci = ci_n, pi = pi_n;
    const double a = a_n, b = b_n, c = c_n;
    // gathers of the current observation (L2-resident tables), issued before the next chunk's
    // streaming loads so that they are first in the queue
    double X0 = 0.0, X1 = 0.0, X2 = 1.0, ps[3] = {0.0, 0.0, 0.0};
    if (valid) {
      X0 = X[3 * (size_t)pi]; X1 = X[3 * (size_t)pi + 1]; X2 = X[3 * (size_t)pi + 2];
      if (JAC) {  // independent loads (no load -> branch -> load chain), masked afterwards
        const uint8_t pvar = d.pt_var[pi];
#pragma unroll
        for (int j = 0; j < 3; ++j) ps[j] = d.pt_scale[3 * (size_t)pi + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) ps[j] = pvar ? ps[j] : 0.0;
      }
    }
    k = kc + stride;
    if (k < K) {
      ci_n = d.obs_cam[k]; pi_n = d.obs_pt[k];
      a_n = d.obs_line[k]; b_n = d.obs_line[K + k]; c_n = d.obs_line[2 * K + k];
      const int64_t kp = k + stride;
      if (kp < K) {
        prefetch_l2(d.obs_cam + kp);
        prefetch_l2(d.obs_pt + kp);
        prefetch_l2(d.obs_line + kp);
        prefetch_l2(d.obs_line + K + kp);
        prefetch_l2(d.obs_line + 2 * K + kp);
      }
    }
    if (!valid) continue;
    double qw, qx, qy, qz, tx, ty, tz, cs[6], prm[8];
    int model;
    if (SMEM) {
      const double* rec = cam_tab + (size_t)ci * kCamRec;
      qw = rec[0]; qx = rec[1]; qy = rec[2]; qz = rec[3];
      tx = rec[4]; ty = rec[5]; tz = rec[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) cs[j] = rec[7 + j];
      const double* irec = intr_tab + (size_t)tab_cam[ci] * kIntrRec;
#pragma unroll
      for (int j = 0; j < 8; ++j) prm[j] = irec[j];
      model = (int)irec[12];
    } else {
      const double4 qq = *reinterpret_cast<const double4*>(q + 4 * (size_t)ci);
      qw = qq.x; qx = qq.y; qy = qq.z; qz = qq.w;
      tx = t[3 * (size_t)ci]; ty = t[3 * (size_t)ci + 1]; tz = t[3 * (size_t)ci + 2];
      const int blk = d.cam_block[ci];
      const unsigned mask = (JAC && blk >= 0) ? d.cam_mask[ci] : 0u;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        cs[j] = ((mask >> j) & 1u) ? d.cam_scale[6 * (size_t)blk + j] : 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) prm[j] = d.img_params[12 * (size_t)ci + j];
      model = d.img_model[ci];
    }
    // ceres::UnitQuaternionRotatePoint
    const double t2 = qw * qx, t3 = qw * qy, t4 = qw * qz, t5 = -qx * qx, t6 = qx * qy;
    const double t7 = qx * qz, t8 = -qy * qy, t9 = qy * qz, t1 = -qz * qz;
    const double R00 = 2.0 * (t8 + t1), R01 = 2.0 * (t6 - t4), R02 = 2.0 * (t3 + t7);
    const double R10 = 2.0 * (t4 + t6), R11 = 2.0 * (t5 + t1), R12 = 2.0 * (t9 - t2);
    const double R20 = 2.0 * (t7 - t3), R21 = 2.0 * (t2 + t9), R22 = 2.0 * (t5 + t8);
    const double pr0 = R00 * X0 + R01 * X1 + R02 * X2 + X0;
    const double pr1 = R10 * X0 + R11 * X1 + R12 * X2 + X1;
    const double pr2 = R20 * X0 + R21 * X1 + R22 * X2 + X2;
    const double p0 = pr0 + tx, p1 = pr1 + ty, p2 = pr2 + tz;
    const double iz = 1.0 / p2;
    const double u = p0 * iz, v = p1 * iz;
    const double alpha = a * u + b * v + c;
    const double lu = u - alpha * a, lv = v - alpha * b;
    double x1, y1, x2, y2, d1xu = 0, d1xv = 0, d1yu = 0, d1yv = 0, d2xu = 0, d2xv = 0, d2yu = 0,
                           d2yv = 0;
    if (EXT && model >= 5) {  // fisheye / FOV / full-OpenCV / thin-prism: out of line, up to 12
                       // parameters read where they lie (the common models keep 8 in registers)
      const double* prm_all = SMEM ? intr_tab + (size_t)tab_cam[ci] * kIntrRec
                                   : d.img_params + 12 * (size_t)ci;
      const WorldToImageResult w1 = world_to_image_ext<JAC>(model, prm_all, u, v);
      const WorldToImageResult w2 = world_to_image_ext<JAC>(model, prm_all, lu, lv);
      x1 = w1.x; y1 = w1.y; x2 = w2.x; y2 = w2.y;
      if (JAC) {
        d1xu = w1.xu; d1xv = w1.xv; d1yu = w1.yu; d1yv = w1.yv;
        d2xu = w2.xu; d2xv = w2.xv; d2yu = w2.yu; d2yv = w2.yv;
      }
    } else {
      world_to_image<JAC, false>(model, prm, u, v, x1, y1, d1xu, d1xv, d1yu, d1yv);
      world_to_image<JAC, false>(model, prm, lu, lv, x2, y2, d2xu, d2xv, d2yu, d2yv);
    }
    const double r0 = x1 - x2, r1 = y1 - y2;
    const double sq = r0 * r0 + r1 * r1;
    double rho0, rho1;
    eval_loss(loss, sq, rho0, rho1);
    cost += 0.5 * rho0;
    if (JAC) {
      const double sr = sqrt(rho1);
      // d r / d(u,v) = D1 - D2 (I - n n^T)
      const double m00 = 1.0 - a * a, m01 = -a * b, m11 = 1.0 - b * b;
      const double E00 = d1xu - (d2xu * m00 + d2xv * m01), E01 = d1xv - (d2xu * m01 + d2xv * m11);
      const double E10 = d1yu - (d2yu * m00 + d2yv * m01), E11 = d1yv - (d2yu * m01 + d2yv * m11);
      // d r / d p  (p = R X + t), pre-multiplied by sqrt(rho')
      double G[2][3];
      G[0][0] = sr * E00 * iz; G[0][1] = sr * E01 * iz; G[0][2] = -(G[0][0] * u + G[0][1] * v);
      G[1][0] = sr * E10 * iz; G[1][1] = sr * E11 * iz; G[1][2] = -(G[1][0] * u + G[1][1] * v);
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        const double g0 = G[row][0], g1 = G[row][1], g2 = G[row][2];
        // rotation (left perturbation q_delta * q, angle 2|delta|): d p / d delta = -2 [R X]_x
        jstore(&JC((6 * row + 0), kc), cs[0] * 2.0 * (g2 * pr1 - g1 * pr2), store_mode);
        jstore(&JC((6 * row + 1), kc), cs[1] * 2.0 * (g0 * pr2 - g2 * pr0), store_mode);
        jstore(&JC((6 * row + 2), kc), cs[2] * 2.0 * (g1 * pr0 - g0 * pr1), store_mode);
        jstore(&JC((6 * row + 3), kc), cs[3] * g0, store_mode);
        jstore(&JC((6 * row + 4), kc), cs[4] * g1, store_mode);
        jstore(&JC((6 * row + 5), kc), cs[5] * g2, store_mode);
        // point: G R   (R = I + R..)
        jstore(&JP((3 * row + 0), kc), ps[0] * (g0 * (R00 + 1.0) + g1 * R10 + g2 * R20), store_mode);
        jstore(&JP((3 * row + 1), kc), ps[1] * (g0 * R01 + g1 * (R11 + 1.0) + g2 * R21), store_mode);
        jstore(&JP((3 * row + 2), kc), ps[2] * (g0 * R02 + g1 * R12 + g2 * (R22 + 1.0)), store_mode);
      }
      jstore(&JR(0, kc), sr * r0, store_mode);
      jstore(&JR(1, kc), sr * r1, store_mode);
    }
  }
  const double total = block_sum(cost, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

// Deterministic final reduction: channel ch sums partials[ch * stride .. + count).
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const double* __restrict__ partials, int count, int stride, int nch,
                       double* __restrict__ scalars, int first_scalar, int accumulate) {
  __shared__ double red[32];
  for (int ch = 0; ch < nch; ++ch) {
    double s = 0.0;
    for (int i = threadIdx.x; i < count; i += blockDim.x) s += partials[(size_t)ch * stride + i];
    const double tot = block_sum(s, red);
    if (threadIdx.x == 0) {
      if (accumulate) scalars[first_scalar + ch] += tot; else scalars[first_scalar + ch] = tot;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Normal-equation blocks.
// ------------------------------------------------------------------------------------------
// V_p = sum J_p^T J_p, g_p = sum J_p^T r.  One thread per observation (coalesced reads of the
// blocked-SoA linearisation); the observations of a point are contiguous, so the per-point sums
// are a segmented warp reduction.  Segments that lie inside one warp batch are stored directly;
// the (at most two per batch) that cross a batch boundary are added atomically into the zeroed
// arrays — two partials commute, so the result is reproducible for tracks of <= 32 observations.
__global__ void __launch_bounds__(kThreads) ba_point_normal_kernel(BaDev d) {
  const int64_t K = d.K;
  const int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int pid = -1;
  double v[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) v[i] = 0.0;
  if (k < K) {
    pid = d.obs_pt[k];
#pragma unroll
    for (int row = 0; row < 2; ++row) {
      const double j0 = JP((3 * row), k), j1 = JP((3 * row + 1), k);
      const double j2 = JP((3 * row + 2), k), rr = JR(row, k);
      v[0] += j0 * j0; v[1] += j0 * j1; v[2] += j0 * j2;
      v[3] += j1 * j1; v[4] += j1 * j2; v[5] += j2 * j2;
      v[6] += j0 * rr; v[7] += j1 * rr; v[8] += j2 * rr;
    }
  }
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int other = __shfl_down_sync(0xffffffffu, pid, off);
    const bool take = (lane + off < 32) && (other == pid);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      const double t = __shfl_down_sync(0xffffffffu, v[i], off);
      if (take) v[i] += t;
    }
  }
  const int prev = __shfl_up_sync(0xffffffffu, pid, 1);
  const bool head = pid >= 0 && (lane == 0 || prev != pid);
  if (!head) return;
  const int64_t batch_end = (k - lane) + 32;  // one past the last observation of the warp batch
  const bool whole = d.pt_start[pid] == k && d.pt_start[pid + 1] <= batch_end;
  const int P = d.P;
  if (whole) {
#pragma unroll
    for (int i = 0; i < 6; ++i) d.V[(size_t)i * P + pid] = v[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) d.gp[(size_t)i * P + pid] = v[6 + i];
  } else {
#pragma unroll
    for (int i = 0; i < 6; ++i) atomicAdd(&d.V[(size_t)i * P + pid], v[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) atomicAdd(&d.gp[(size_t)i * P + pid], v[6 + i]);
  }
}

// U_c = sum J_c^T J_c, g_c = sum J_c^T r: kCamSplit CTAs per camera block, each reduces its part
// of the camera's observation list (gather through cam_obs) into 27 partial sums; the partials are
// added in a fixed order by ba_camera_reduce_kernel.
constexpr int kCamSplit = 4;
__global__ void __launch_bounds__(128) ba_camera_normal_kernel(BaDev d) {
  __shared__ double red[32];
  const int b = blockIdx.x / kCamSplit, part = blockIdx.x % kCamSplit;
  double u[21], g[6];
#pragma unroll
  for (int i = 0; i < 21; ++i) u[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) g[i] = 0.0;
  const int64_t c0 = d.cam_start[b], c1 = d.cam_start[b + 1];
  const int64_t len = (c1 - c0 + kCamSplit - 1) / kCamSplit;
  const int64_t i0 = c0 + part * len, i1 = (i0 + len < c1) ? (i0 + len) : c1;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int64_t k = d.cam_obs[i];
#pragma unroll
    for (int row = 0; row < 2; ++row) {
      double j[6];
#pragma unroll
      for (int a = 0; a < 6; ++a) j[a] = JC((6 * row + a), k);
      const double rr = JR(row, k);
      int idx = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int c = a; c < 6; ++c) u[idx++] += j[a] * j[c];
        g[a] += j[a] * rr;
      }
    }
  }
  double* out = d.Upart + 27 * (size_t)blockIdx.x;
#pragma unroll
  for (int i = 0; i < 21; ++i) {
    const double tot = block_sum(u[i], red);
    if (threadIdx.x == 0) out[i] = tot;
  }
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    const double tg = block_sum(g[a], red);
    if (threadIdx.x == 0) out[21 + a] = tg;
  }
}

__global__ void ba_camera_reduce_kernel(BaDev d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = t / 27, i = t - 27 * b;
  if (b >= d.NB) return;
  double s = 0.0;
#pragma unroll
  for (int part = 0; part < kCamSplit; ++part) s += d.Upart[27 * ((size_t)b * kCamSplit + part) + i];
  if (i >= 21) {
    d.gc[6 * (size_t)b + (i - 21)] = s;
    return;
  }
  int a = 0, rem = i;  // upper-triangle index -> (a, c)
  while (rem >= 6 - a) {
    rem -= 6 - a;
    ++a;
  }
  const int c = a + rem;
  d.U[36 * (size_t)b + 6 * a + c] = s;
  d.U[36 * (size_t)b + 6 * c + a] = s;
}

__global__ void ba_jacobi_scales_kernel(BaDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * d.NB) {
    const int b = i / 6, a = i % 6;
    d.cam_scale[i] = 1.0 / (1.0 + sqrt(d.U[36 * (size_t)b + 7 * a]));
  }
  const int j = i - 6 * d.NB;
  if (j >= 0 && j < 3 * d.P) {
    const int p = j / 3, a = j % 3;
    const int vi = (a == 0) ? 0 : (a == 1 ? 3 : 5);
    d.pt_scale[j] = 1.0 / (1.0 + sqrt(d.V[(size_t)vi * d.P + p]));
  }
}

// ------------------------------------------------------------------------------------------
// Back-substitution, model cost change, candidate state.
// ------------------------------------------------------------------------------------------
// (1) one thread per observation: u = J_c dc (kept for the model-cost pass), and the point-side
//     sums s_p = sum_e J_p,e^T u_e (into d.dp, zeroed before).  The observations of a point are
//     contiguous, so this is the segmented warp reduction of ba_point_normal_kernel: segments
//     inside one warp batch are stored, the (at most two per batch) that cross a batch boundary
//     are added atomically to zero — two partials commute, so the sums, and with them the whole
//     solve, are reproducible run to run for tracks of <= 32 observations.
__global__ void __launch_bounds__(kThreads) ba_backsub_accum_kernel(BaDev d) {
  const int64_t K = d.K;
  const int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int pid = -1;
  double v[3] = {0.0, 0.0, 0.0};
  if (k < K) {
    pid = d.obs_pt[k];
    const int b = d.cam_block[d.obs_cam[k]];
    double u0 = 0.0, u1 = 0.0;
    if (b >= 0) {
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const double dca = d.dc[6 * b + a];
        u0 += JC(a, k) * dca;
        u1 += JC((6 + a), k) * dca;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = JP(c, k) * u0 + JP((3 + c), k) * u1;
    }
    d.u[k] = u0;
    d.u[K + k] = u1;
  }
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int other = __shfl_down_sync(0xffffffffu, pid, off);
    const bool take = (lane + off < 32) && (other == pid);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double t = __shfl_down_sync(0xffffffffu, v[i], off);
      if (take) v[i] += t;
    }
  }
  const int prev = __shfl_up_sync(0xffffffffu, pid, 1);
  const bool head = pid >= 0 && (lane == 0 || prev != pid);
  if (!head) return;
  const int64_t batch_end = (k - lane) + 32;
  const bool whole = d.pt_start[pid] == k && d.pt_start[pid + 1] <= batch_end;
  const int P = d.P;
  if (whole) {
#pragma unroll
    for (int c = 0; c < 3; ++c) d.dp[(size_t)c * P + pid] = v[c];
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(&d.dp[(size_t)c * P + pid], v[c]);
  }
}

// (2) one thread per point: dp = -V^-1 acc, candidate point, step / x norm partials.
__global__ void __launch_bounds__(kThreads)
ba_point_step_kernel(BaDev d, double* __restrict__ partials, int stride) {
  __shared__ double red[32];
  const int p = blockIdx.x * kThreads + threadIdx.x;
  const int P = d.P;
  double step_sq = 0.0, x_sq = 0.0;
  if (p < P) {
    // acc = g_p + sum_e J_p,e^T u_e
    const double acc0 = d.gp[p] + d.dp[p], acc1 = d.gp[P + p] + d.dp[P + p],
                 acc2 = d.gp[2 * P + p] + d.dp[2 * P + p];
    const double w00 = d.Vinv[p], w01 = d.Vinv[P + p], w02 = d.Vinv[2 * P + p];
    const double w11 = d.Vinv[3 * P + p], w12 = d.Vinv[4 * P + p], w22 = d.Vinv[5 * P + p];
    const double dp0 = -(w00 * acc0 + w01 * acc1 + w02 * acc2);
    const double dp1 = -(w01 * acc0 + w11 * acc1 + w12 * acc2);
    const double dp2 = -(w02 * acc0 + w12 * acc1 + w22 * acc2);
    d.dp[p] = dp0; d.dp[P + p] = dp1; d.dp[2 * P + p] = dp2;
    const double x0 = d.X[3 * (size_t)p], x1 = d.X[3 * (size_t)p + 1], x2 = d.X[3 * (size_t)p + 2];
    const double s0 = dp0 * d.pt_scale[3 * (size_t)p], s1 = dp1 * d.pt_scale[3 * (size_t)p + 1];
    const double s2 = dp2 * d.pt_scale[3 * (size_t)p + 2];
    d.Xn[3 * (size_t)p] = x0 + s0;
    d.Xn[3 * (size_t)p + 1] = x1 + s1;
    d.Xn[3 * (size_t)p + 2] = x2 + s2;
    if (d.pt_var[p]) {
      step_sq = s0 * s0 + s1 * s1 + s2 * s2;
      x_sq = x0 * x0 + x1 * x1 + x2 * x2;
    }
  }
  const double ts = block_sum(step_sq, red);
  const double tx = block_sum(x_sq, red);
  if (threadIdx.x == 0) {
    partials[stride + blockIdx.x] = ts;
    partials[2 * stride + blockIdx.x] = tx;
  }
}

// (3) one thread per observation: model_cost_change = -sum (J d) . (r + J d / 2)
__global__ void __launch_bounds__(kThreads)
ba_model_cost_kernel(BaDev d, double* __restrict__ partials) {
  __shared__ double red[32];
  const int64_t K = d.K;
  const int64_t k = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  double model = 0.0;
  if (k < K) {
    const int p = d.obs_pt[k];
    const int P = d.P;
    const double dp0 = d.dp[p], dp1 = d.dp[P + p], dp2 = d.dp[2 * P + p];
#pragma unroll
    for (int row = 0; row < 2; ++row) {
      const double m = d.u[row * K + k] + JP((3 * row), k) * dp0 +
                       JP((3 * row + 1), k) * dp1 + JP((3 * row + 2), k) * dp2;
      model -= m * (JR(row, k) + 0.5 * m);
    }
  }
  const double tm = block_sum(model, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = tm;
}

// ceres::QuaternionParameterization::Plus
__device__ __forceinline__ void quaternion_plus(const double* x, const double* delta, double* out) {
  const double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    const double s = sin(nd) / nd;
    const double q0 = cos(nd), q1 = s * delta[0], q2 = s * delta[1], q3 = s * delta[2];
    out[0] = q0 * x[0] - q1 * x[1] - q2 * x[2] - q3 * x[3];
    out[1] = q0 * x[1] + q1 * x[0] + q2 * x[3] - q3 * x[2];
    out[2] = q0 * x[2] - q1 * x[3] + q2 * x[0] + q3 * x[1];
    out[3] = q0 * x[3] + q1 * x[2] - q2 * x[1] + q3 * x[0];
  } else {
    out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3];
  }
}

// single CTA (cameras are few): candidate poses + camera part of step / x norms
__global__ void __launch_bounds__(kThreads) ba_camera_update_kernel(BaDev d, int count_norms) {
  __shared__ double red[32];
  double step_sq = 0.0, x_sq = 0.0;
  for (int i = threadIdx.x; i < d.C; i += kThreads) {
    const int b = d.cam_block[i];
    double qv[4], tv[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) qv[k] = d.q[4 * (size_t)i + k];
#pragma unroll
    for (int k = 0; k < 3; ++k) tv[k] = d.t[3 * (size_t)i + k];
    if (b < 0) {
#pragma unroll
      for (int k = 0; k < 4; ++k) d.qn[4 * (size_t)i + k] = qv[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) d.tn[3 * (size_t)i + k] = tv[k];
      continue;
    }
    double dl[6], qp[4];
#pragma unroll
    for (int a = 0; a < 6; ++a) dl[a] = d.dc[6 * b + a] * d.cam_scale[6 * (size_t)b + a];
    quaternion_plus(qv, dl, qp);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      d.qn[4 * (size_t)i + k] = qp[k];
      step_sq += (qp[k] - qv[k]) * (qp[k] - qv[k]);
      x_sq += qv[k] * qv[k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      d.tn[3 * (size_t)i + k] = tv[k] + dl[3 + k];
      step_sq += dl[3 + k] * dl[3 + k];
      x_sq += tv[k] * tv[k];
    }
  }
  const double ts = block_sum(step_sq, red);
  const double tx = block_sum(x_sq, red);
  if (threadIdx.x == 0) {  // replicated across ranks: only one rank contributes to the sums
    d.scalars[kStepSq] = count_norms ? ts : 0.0;
    d.scalars[kXSq] = count_norms ? tx : 0.0;
  }
}

__global__ void ba_axpby_kernel(double* __restrict__ out, const double* __restrict__ a,
                                const double* __restrict__ b, double beta, size_t n) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + beta * b[i];
}

// Lower triangle (rows 0..n-1, columns 0..r) plus the right-hand-side row n of the bordered
// reduced matrix <-> packed buffer: the multi-GPU all-reduce then moves half the bytes.
// Row r starts at r (r + 1) / 2; the rhs row (n entries) follows the triangle.
__global__ void __launch_bounds__(kThreads)
ba_pack_lower_kernel(const double* __restrict__ S, int n, int ld, double* __restrict__ packed) {
  const int r = blockIdx.x;  // 0..n (row n = rhs)
  const int len = r < n ? r + 1 : n;
  const size_t off = r < n ? (size_t)r * (r + 1) / 2 : (size_t)n * (n + 1) / 2;
  for (int c = threadIdx.x; c < len; c += kThreads) packed[off + c] = S[(size_t)r * ld + c];
}
__global__ void __launch_bounds__(kThreads)
ba_unpack_lower_kernel(double* __restrict__ S, int n, int ld, const double* __restrict__ packed) {
  const int r = blockIdx.x;
  const int len = r < n ? r + 1 : n;
  const size_t off = r < n ? (size_t)r * (r + 1) / 2 : (size_t)n * (n + 1) / 2;
  for (int c = threadIdx.x; c < len; c += kThreads) S[(size_t)r * ld + c] = packed[off + c];
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr),
            (unsigned long long)__double_as_longlong(v));
}

// gradient of the unscaled problem: g = g_scaled / scale (J_s = J diag(scale)); max |x - Plus(x, -g)|
__global__ void ba_gradient_max_kernel(BaDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (i < d.NB) {
    const int img = d.block_img[i];
    double g[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) g[a] = d.gc[6 * (size_t)i + a] / d.cam_scale[6 * (size_t)i + a];
    const double nd[3] = {-g[0], -g[1], -g[2]};
    double qv[4], qp[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) qv[k] = d.q[4 * (size_t)img + k];
    quaternion_plus(qv, nd, qp);
#pragma unroll
    for (int k = 0; k < 4; ++k) m = fmax(m, fabs(qp[k] - qv[k]));
#pragma unroll
    for (int a = 3; a < 6; ++a) m = fmax(m, fabs(g[a]));
  }
  const int p = i - d.NB;
  if (p >= 0 && p < d.P) {
#pragma unroll
    for (int a = 0; a < 3; ++a)
      m = fmax(m, fabs(d.gp[(size_t)a * d.P + p] / d.pt_scale[3 * (size_t)p + a]));
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, s));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomic_max_nonneg(&d.scalars[kGradMax], m);
}

}  // namespace

// ============================================================================================
int launch_linearize(const BaDev& d, const double* q, const double* t, const double* X,
                     bool jacobians, BaLoss loss, cudaStream_t s) {
  const int chunks = (int)((d.K + kThreads - 1) / kThreads);
  if (chunks == 0) {
    cudaMemsetAsync(d.scalars + kCost, 0, sizeof(double), s);
    return 0;
  }
  // Persistent CTAs with the per-image table in shared memory when it fits (2 CTAs / SM);
  // otherwise one CTA per chunk gathering from global memory.
  static PerDevice<> per_device;
  static const int store_mode = tune_int("PPSFM_BA_J_STREAM", 0);
  const auto& dev = per_device.get([](const DeviceFacts& f, int&) {
    cudaError_t e = cudaSuccess;
    auto optin = [&](auto kernel) {
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 f.max_smem_optin - 1024);
    };
    optin(ba_linearize_kernel<true, true, false>);
    optin(ba_linearize_kernel<false, true, false>);
    optin(ba_linearize_kernel<true, true, true>);
    optin(ba_linearize_kernel<false, true, true>);
    return e;
  });
  // (a failed attribute call surfaces as a launch error at the caller's next CUDA check)
  const int num_sms = dev.facts.num_sms, max_smem = dev.facts.max_smem_optin;
  const size_t tab_bytes = (size_t)d.C * (kCamRec * sizeof(double) + sizeof(int)) +
                           (size_t)d.num_cameras * kIntrRec * sizeof(double) + 16;
  const bool use_smem = tab_bytes <= (size_t)(max_smem - 2048) && chunks > 2 * num_sms;
  int blocks = chunks;
  if (use_smem) {
    int per_sm = (int)((size_t)(227 * 1024) / (tab_bytes + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);
    blocks = num_sms * per_sm;
    if (blocks > chunks) blocks = chunks;
  }
  if (blocks > d.num_partials) blocks = d.num_partials;
  auto launch = [&](auto kernel, size_t smem) {
    kernel<<<blocks, kThreads, smem, s>>>(d, q, t, X, loss, d.partials, chunks, store_mode);
  };
  const int variant = (jacobians ? 4 : 0) | (use_smem ? 2 : 0) | (d.has_ext_models ? 1 : 0);
  switch (variant) {
    case 7: launch(ba_linearize_kernel<true, true, true>, tab_bytes); break;
    case 6: launch(ba_linearize_kernel<true, true, false>, tab_bytes); break;
    case 5: launch(ba_linearize_kernel<true, false, true>, 0); break;
    case 4: launch(ba_linearize_kernel<true, false, false>, 0); break;
    case 3: launch(ba_linearize_kernel<false, true, true>, tab_bytes); break;
    case 2: launch(ba_linearize_kernel<false, true, false>, tab_bytes); break;
    case 1: launch(ba_linearize_kernel<false, false, true>, 0); break;
    default: launch(ba_linearize_kernel<false, false, false>, 0); break;
  }
  reduce_partials_kernel<<<1, 1024, 0, s>>>(d.partials, blocks, d.num_partials, 1, d.scalars,
                                            kCost, 0);
  return 2;
}

int launch_normal_equations(const BaDev& d, cudaStream_t s) {
  int n = 0;
  if (d.P > 0) {
    cudaMemsetAsync(d.V, 0, sizeof(double) * 6 * (size_t)d.P, s);
    cudaMemsetAsync(d.gp, 0, sizeof(double) * 3 * (size_t)d.P, s);
    if (d.K > 0) {
      ba_point_normal_kernel<<<(unsigned)((d.K + kThreads - 1) / kThreads), kThreads, 0, s>>>(d);
      ++n;
    }
  }
  if (d.NB > 0) {
    ba_camera_normal_kernel<<<d.NB * kCamSplit, 128, 0, s>>>(d);
    ba_camera_reduce_kernel<<<(27 * d.NB + 255) / 256, 256, 0, s>>>(d);
    n += 2;
  }
  return n;
}

int launch_jacobi_scales(const BaDev& d, cudaStream_t s) {
  const int total = 6 * d.NB + 3 * d.P;
  if (total == 0) return 0;
  ba_jacobi_scales_kernel<<<(total + 255) / 256, 256, 0, s>>>(d);
  return 1;
}

size_t packed_lower_doubles(int n) { return (size_t)n * (n + 1) / 2 + (size_t)n; }
void launch_pack_lower(const double* S, int n, int ld, double* packed, cudaStream_t s) {
  if (n > 0) ba_pack_lower_kernel<<<n + 1, kThreads, 0, s>>>(S, n, ld, packed);
}
void launch_unpack_lower(double* S, int n, int ld, const double* packed, cudaStream_t s) {
  if (n > 0) ba_unpack_lower_kernel<<<n + 1, kThreads, 0, s>>>(S, n, ld, packed);
}

// Xg[3 (p world + rank) + k] = X[3 p + k] - X0[3 p + k]: this rank's point displacements at the
// caller's indices (download of a sharded solve)
__global__ void ba_scatter_displacement_kernel(double* __restrict__ Xg, const double* __restrict__ X,
                                               const double* __restrict__ X0, int P, int world,
                                               int rank) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * P) return;
  const int p = i / 3, k = i - 3 * p;
  Xg[3 * ((size_t)p * world + rank) + k] = X[i] - X0[i];
}
void launch_scatter_displacement(double* Xg, const double* X, const double* X0, int P, int world,
                                 int rank, cudaStream_t s) {
  if (P > 0)
    ba_scatter_displacement_kernel<<<(3 * P + 255) / 256, 256, 0, s>>>(Xg, X, X0, P, world, rank);
}

void launch_axpby(double* out, const double* a, const double* b, double beta, size_t n,
                  cudaStream_t s) {
  if (n == 0) return;
  ba_axpby_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, a, b, beta, n);
}

int launch_backsubstitute_and_update(const BaDev& d, bool count_camera_norms, cudaStream_t s) {
  int n = 0;
  ba_camera_update_kernel<<<1, kThreads, 0, s>>>(d, count_camera_norms ? 1 : 0);
  ++n;
  const int pblocks = (d.P + kThreads - 1) / kThreads;
  const int oblocks = (int)((d.K + kThreads - 1) / kThreads);
  if (pblocks > 0 && oblocks > 0) {
    cudaMemsetAsync(d.dp, 0, sizeof(double) * 3 * (size_t)d.P, s);
    ba_backsub_accum_kernel<<<oblocks, kThreads, 0, s>>>(d);
    n += launch_intr_backsub(d, s);  // (variable intrinsics only)
    ba_point_step_kernel<<<pblocks, kThreads, 0, s>>>(d, d.partials, d.num_partials);
    // step / x norms accumulate on top of the camera part written by ba_camera_update_kernel
    reduce_partials_kernel<<<1, 1024, 0, s>>>(d.partials + d.num_partials, pblocks,
                                              d.num_partials, 2, d.scalars, kStepSq, 1);
    ba_model_cost_kernel<<<oblocks, kThreads, 0, s>>>(d, d.partials);
    reduce_partials_kernel<<<1, 1024, 0, s>>>(d.partials, oblocks, d.num_partials, 1, d.scalars,
                                              kModelChange, 0);
    n += 5;
  } else {
    cudaMemsetAsync(d.scalars + kModelChange, 0, sizeof(double), s);
  }
  return n;
}

int launch_gradient_max_norm(const BaDev& d, cudaStream_t s) {
  cudaMemsetAsync(d.scalars + kGradMax, 0, sizeof(double), s);
  const int total = d.NB + d.P;
  if (total == 0) return 0;
  ba_gradient_max_kernel<<<(total + 255) / 256, 256, 0, s>>>(d);
  return 1;
}

}  // namespace ppsfm
