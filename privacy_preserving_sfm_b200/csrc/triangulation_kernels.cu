// triangulation_kernels.cu — batched robust line triangulation.
//
// Replaces, for a whole batch of tracks at once,
//   EstimateTriangulation              src/estimators/triangulation.cc:118-149
//   TriangulationEstimator::Estimate   :55-95   (multi-view point from lines, cheirality,
//                                               minimum triangulation angle)
//   TriangulationEstimator::Residuals  :97-116  (squared angular / line-reprojection error)
//   TriangulateMultiViewPoint          src/base/triangulation.cc:41-57 (null vector of the n x 4
//                                               system rows l_i^T P_i)
//   LORANSAC<..., CombinationSampler>  src/optim/loransac.h:91-234,
//                                      src/optim/combination_sampler.cc:41-70
// In the mapper this runs once per new track (tens of thousands of tiny, independent problems per
// image, src/sfm/incremental_triangulator.cc:468-560): one thread per track executes the serial
// LORANSAC loop — samples are the 3-combinations of the track in lexicographic order, each model
// is scored on the whole track, the local optimisation re-estimates from the inliers.
// The n x 4 null vector is computed without storing the system: rows are folded into a 4 x 4
// triangular factor by Givens rotations (a streaming QR), whose SVD (one-sided Jacobi) has the
// same right singular vectors as the full system.
// The adaptive trial bound (RANSAC::ComputeNumTrials, src/optim/ransac.h:158-176, k = 3) comes
// from a host-computed table so that it is bit-identical to the CPU path.
#include <cfloat>
#include <cmath>
#include <vector>

#include "camera_models.cuh"
#include "common.h"

namespace ppsfm {

namespace {

constexpr int kTableN = 128;  // trial-bound table covers tracks up to this length

struct TriDev {
  int C, T, num_cameras;
  int64_t O;
  const double* proj;     // [C][12] row-major 3x4
  const double* centers;  // [C][3]
  const double *cam_params, *obs_line;
  const int *img_cam, *cam_model, *cam_w, *cam_h, *obs_image;
  const int64_t* track_start;
  const unsigned long long* trial_table;  // [kTableN + 1][kTableN + 1]
  double* xyz;
  uint8_t *success, *inlier_mask;
  unsigned* num_trials;
};

struct TriOpts {
  double min_tri_angle, max_residual, confidence, multiplier;
  unsigned long long min_num_trials, max_num_trials;
  int residual_type;  // 0 ANGULAR_ERROR, 1 REPROJECTION_ERROR
  int exhaustive_threshold;
};

__global__ void tri_pose_kernel(int C, const double* __restrict__ q, const double* __restrict__ t,
                                double* __restrict__ proj, double* __restrict__ centers) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  const double* qv = q + 4 * (size_t)i;
  const double n = sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2] + qv[3] * qv[3]);
  const double w = qv[0] / n, x = qv[1] / n, y = qv[2] / n, z = qv[3] / n;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  double R[9] = {1.0 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.0 - (txx + tzz),
                 tyz - twx, txz - twy, tyz + twx, 1.0 - (txx + tyy)};
  const double* tv = t + 3 * (size_t)i;
  double* P = proj + 12 * (size_t)i;
  for (int r = 0; r < 3; ++r) {
    P[4 * r] = R[3 * r]; P[4 * r + 1] = R[3 * r + 1]; P[4 * r + 2] = R[3 * r + 2];
    P[4 * r + 3] = tv[r];
  }
  for (int k = 0; k < 3; ++k)
    centers[3 * (size_t)i + k] = -(R[k] * tv[0] + R[3 + k] * tv[1] + R[6 + k] * tv[2]);
}

// streaming QR of the rows l^T P: R is 4x4 upper triangular (row-major)
__device__ __forceinline__ void fold_row(double R[16], const double* l, const double* P) {
  double a[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) a[c] = l[0] * P[c] + l[1] * P[4 + c] + l[2] * P[8 + c];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (a[c] == 0.0) continue;
    const double rr = hypot(R[5 * c], a[c]);
    const double cs = R[5 * c] / rr, sn = a[c] / rr;
#pragma unroll
    for (int j = c; j < 4; ++j) {
      const double t = cs * R[4 * c + j] + sn * a[j];
      a[j] = cs * a[j] - sn * R[4 * c + j];
      R[4 * c + j] = t;
    }
  }
}

// right singular vector of the smallest singular value of the 4x4 matrix W (one-sided Jacobi),
// returned de-homogenised
__device__ __forceinline__ void null_point(double W[16], double X[3]) {
  double V[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) V[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          alpha += W[4 * r + p] * W[4 * r + p];
          beta += W[4 * r + q] * W[4 * r + q];
          gamma += W[4 * r + p] * W[4 * r + q];
        }
        if (fabs(gamma) <= 1e-300 || fabs(gamma) <= 2.3e-16 * sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const double wp = W[4 * r + p], wq = W[4 * r + q];
          W[4 * r + p] = c * wp - s * wq;
          W[4 * r + q] = s * wp + c * wq;
          const double vp = V[4 * r + p], vq = V[4 * r + q];
          V[4 * r + p] = c * vp - s * vq;
          V[4 * r + q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  int best = 0;
  double best_norm = DBL_MAX;
  for (int j = 0; j < 4; ++j) {
    double s2 = 0;
    for (int r = 0; r < 4; ++r) s2 += W[4 * r + j] * W[4 * r + j];
    if (s2 < best_norm) {
      best_norm = s2;
      best = j;
    }
  }
  const double w = V[12 + best];
  X[0] = V[best] / w; X[1] = V[4 + best] / w; X[2] = V[8 + best] / w;
}

__device__ __forceinline__ double tri_angle(const double* c1, const double* c2, const double* X) {
  double b2 = 0, r1 = 0, r2 = 0;
  for (int k = 0; k < 3; ++k) {
    b2 += (c1[k] - c2[k]) * (c1[k] - c2[k]);
    r1 += (X[k] - c1[k]) * (X[k] - c1[k]);
    r2 += (X[k] - c2[k]) * (X[k] - c2[k]);
  }
  const double den = 2.0 * sqrt(r1 * r2);
  if (den == 0.0) return 0.0;
  const double angle = fabs(acos((r1 + r2 - b2) / den));
  return fmin(angle, M_PI - angle);
}

// TriangulationEstimator::Residuals for one observation
__device__ __forceinline__ double residual(const TriDev& d, const TriOpts& o, int64_t k,
                                           const double* X) {
  const int img = d.obs_image[k];
  const double* P = d.proj + 12 * (size_t)img;
  const double* l = d.obs_line + 3 * (size_t)k;
  const double r0 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3];
  const double r1 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7];
  const double r2 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11];
  const int cam = d.img_cam[img];
  const double* prm = d.cam_params + 12 * (size_t)cam;
  const int model = d.cam_model[cam];
  double x1, y1, j0, j1, j2, j3;
  if (o.residual_type == 0) {
    // CalculateNormalizedLineAngularError (src/base/projection.cc:241-260)
    if (r2 < 0) return DBL_MAX;
    world_to_image<false>(model, prm, r0 / r2, r1 / r2, x1, y1, j0, j1, j2, j3);
    if (x1 < 0 || x1 >= (double)d.cam_w[cam] || y1 < 0 || y1 >= (double)d.cam_h[cam]) return DBL_MAX;
    const double nl = sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
    const double nr = sqrt(r0 * r0 + r1 * r1 + r2 * r2);
    const double dot = (l[0] / nl) * (r0 / nr) + (l[1] / nl) * (r1 / nr) + (l[2] / nl) * (r2 / nr);
    const double e = fabs(M_PI_2 - acos(fabs(dot)));
    return e * e;
  }
  // CalculateSquaredLineReprojectionError (src/base/projection.cc:162-203)
  if (r2 < DBL_EPSILON) return DBL_MAX;
  const double inv = 1.0 / r2;
  const double u = inv * r0, v = inv * r1;
  const double alpha = l[0] * u + l[1] * v + l[2];
  const double lu = u - l[0] * alpha, lv = v - l[1] * alpha;
  world_to_image<false>(model, prm, u, v, x1, y1, j0, j1, j2, j3);
  if (!(x1 >= 0.0 && x1 < (double)d.cam_w[cam] && y1 >= 0.0 && y1 < (double)d.cam_h[cam])) return DBL_MAX;
  double x2, y2;
  world_to_image<false>(model, prm, lu, lv, x2, y2, j0, j1, j2, j3);
  return (x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2);
}

__device__ __forceinline__ bool positive_depth(const double* P, const double* X) {
  return P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] >= DBL_EPSILON;
}

// support of a model over the whole track (InlierSupportMeasurer::Evaluate, index order)
__device__ __forceinline__ void support_of(const TriDev& d, const TriOpts& o, int64_t k0, int64_t k1,
                                           const double* X, unsigned* num, double* sum) {
  unsigned n = 0;
  double s = 0;
  for (int64_t k = k0; k < k1; ++k) {
    const double r = residual(d, o, k, X);
    if (r <= o.max_residual) {
      ++n;
      s += r;
    }
  }
  *num = n;
  *sum = s;
}

__device__ __forceinline__ unsigned long long trials_needed(const TriDev& d, const TriOpts& o,
                                                            unsigned num_inliers, unsigned n) {
  if (n <= kTableN) return d.trial_table[(size_t)n * (kTableN + 1) + num_inliers];
  const double ratio = num_inliers / (double)n;
  const double nom = 1 - o.confidence;
  if (nom <= 0) return ~0ull;
  const double denom = 1 - ratio * ratio * ratio;
  if (denom <= 0) return 1;
  return (unsigned long long)ceil(log(nom) / log(denom) * o.multiplier);
}

__global__ void __launch_bounds__(128) tri_ransac_kernel(TriDev d, TriOpts o) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.T) return;
  const int64_t k0 = d.track_start[t], k1 = d.track_start[t + 1];
  const unsigned n = (unsigned)(k1 - k0);
  d.success[t] = 0;
  d.num_trials[t] = 0;
  for (int64_t k = k0; k < k1; ++k) d.inlier_mask[k] = 0;
  if (n < 3) return;  // EstimateTriangulation: point_data.size() < 3 -> false
  // NChooseK(n, 3) with the reference's integer recursion (src/util/math.cc:36-42)
  const unsigned long long c1 = n - 2, c2 = ((unsigned long long)(n - 1) * c1) / 2;
  const unsigned long long num_comb = ((unsigned long long)n * c2) / 3;
  unsigned long long min_trials = o.min_num_trials;
  if ((int)n <= o.exhaustive_threshold) min_trials = num_comb;
  const unsigned long long max_trials = o.max_num_trials < num_comb ? o.max_num_trials : num_comb;
  unsigned long long dyn_max = max_trials;
  unsigned best_num = 0;
  double best_sum = DBL_MAX, best[3] = {0, 0, 0};
  bool abort = false;
  unsigned a = 0, b = 1, c = 2;  // current 3-combination (lexicographic)
  unsigned long long trial = 0;
  for (; trial < max_trials; ++trial) {
    if (abort) {
      trial += 1;
      break;
    }
    const int64_t s[3] = {k0 + a, k0 + b, k0 + c};
    // advance (CombinationSampler::Sample wraps around after the last combination)
    if (c + 1 < n) ++c;
    else if (b + 2 < n) { ++b; c = b + 1; }
    else if (a + 3 < n) { ++a; b = a + 1; c = b + 1; }
    else { a = 0; b = 1; c = 2; }
    double R[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) R[i] = 0.0;
    for (int i = 0; i < 3; ++i)
      fold_row(R, d.obs_line + 3 * (size_t)s[i], d.proj + 12 * (size_t)d.obs_image[s[i]]);
    double X[3];
    null_point(R, X);
    bool ok = true;
    for (int i = 0; i < 3; ++i) ok = ok && positive_depth(d.proj + 12 * (size_t)d.obs_image[s[i]], X);
    if (ok) {
      ok = false;
      for (int i = 0; i < 3 && !ok; ++i)
        for (int j = 0; j < i; ++j)
          if (tri_angle(d.centers + 3 * (size_t)d.obs_image[s[i]],
                        d.centers + 3 * (size_t)d.obs_image[s[j]], X) >= o.min_tri_angle) {
            ok = true;
            break;
          }
    }
    if (!ok) continue;  // no model from this sample
    unsigned num;
    double sum;
    support_of(d, o, k0, k1, X, &num, &sum);
    if (num > best_num || (num == best_num && sum < best_sum)) {
      best_num = num; best_sum = sum;
      best[0] = X[0]; best[1] = X[1]; best[2] = X[2];
      if (num > 3) {  // local optimisation from the inliers
        // inlier set of the sample model: a bit mask for tracks up to 64 views, recomputed on the
        // fly for longer ones
        unsigned long long bits = 0;
        if (n <= 64)
          for (unsigned i = 0; i < n; ++i)
            if (residual(d, o, k0 + i, X) <= o.max_residual) bits |= 1ull << i;
        auto is_inlier = [&](int64_t k) {
          return n <= 64 ? ((bits >> (unsigned)(k - k0)) & 1ull) != 0
                         : residual(d, o, k, X) <= o.max_residual;
        };
        double RL[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) RL[i] = 0.0;
        for (int64_t k = k0; k < k1; ++k)
          if (is_inlier(k))
            fold_row(RL, d.obs_line + 3 * (size_t)k, d.proj + 12 * (size_t)d.obs_image[k]);
        double XL[3];
        null_point(RL, XL);
        bool okl = true;
        for (int64_t k = k0; k < k1 && okl; ++k)
          if (is_inlier(k)) okl = positive_depth(d.proj + 12 * (size_t)d.obs_image[k], XL);
        if (okl) {
          okl = false;
          for (int64_t i = k0; i < k1 && !okl; ++i) {
            if (!is_inlier(i)) continue;
            for (int64_t j = k0; j < i; ++j) {
              if (!is_inlier(j)) continue;
              if (tri_angle(d.centers + 3 * (size_t)d.obs_image[i],
                            d.centers + 3 * (size_t)d.obs_image[j], XL) >= o.min_tri_angle) {
                okl = true;
                break;
              }
            }
          }
        }
        if (okl) {
          unsigned numl;
          double suml;
          support_of(d, o, k0, k1, XL, &numl, &suml);
          if (numl > best_num || (numl == best_num && suml < best_sum)) {
            best_num = numl; best_sum = suml;
            best[0] = XL[0]; best[1] = XL[1]; best[2] = XL[2];
          }
        }
      }
      dyn_max = trials_needed(d, o, best_num, n);
    }
    if (trial >= dyn_max && trial >= min_trials) abort = true;  // (one model per sample)
  }
  d.num_trials[t] = (unsigned)trial;
  if (best_num < 3) return;
  d.success[t] = 1;
  d.xyz[3 * (size_t)t] = best[0]; d.xyz[3 * (size_t)t + 1] = best[1]; d.xyz[3 * (size_t)t + 2] = best[2];
  for (int64_t k = k0; k < k1; ++k) d.inlier_mask[k] = residual(d, o, k, best) <= o.max_residual ? 1 : 0;
}

size_t HostNumTrials(size_t num_inliers, size_t num_samples, double confidence, double multiplier) {
  const double inlier_ratio = num_inliers / static_cast<double>(num_samples);
  const double nom = 1 - confidence;
  if (nom <= 0) return std::numeric_limits<size_t>::max();
  const double denom = 1 - std::pow(inlier_ratio, 3);
  if (denom <= 0) return 1;
  return static_cast<size_t>(std::ceil(std::log(nom) / std::log(denom) * multiplier));
}

}  // namespace
}  // namespace ppsfm

extern "C" {

void ppsfm_triangulation_options_default(ppsfm_triangulation_options* o) {
  if (!o) return;
  // EstimateTriangulationOptions (src/estimators/triangulation.h:117-135) + RANSACOptions defaults
  o->min_tri_angle = 0.0;
  o->residual_type = 0;
  o->max_error = 0.0;
  o->min_inlier_ratio = 0.1;
  o->confidence = 0.99;
  o->dyn_num_trials_multiplier = 3.0;
  o->min_num_trials = 0;
  o->max_num_trials = 0xffffffffffffffffull;
  o->exhaustive_threshold = 0;
}

int ppsfm_estimate_triangulation_batch(ppsfm_ctx* ctx, const ppsfm_filter_problem* pb,
                                       const ppsfm_triangulation_options* opt, double* xyz,
                                       uint8_t* success, uint8_t* inlier_mask,
                                       uint32_t* num_trials) {
  using namespace ppsfm;
  if (!ctx || !pb || !opt || !xyz || !success || !inlier_mask) return PPSFM_ERR_INVALID;
  // RANSACOptions::Check (src/optim/ransac.h:68-75)
  if (!(opt->max_error > 0) || opt->min_inlier_ratio < 0 || opt->min_inlier_ratio > 1 ||
      opt->confidence < 0 || opt->confidence > 1 || opt->min_num_trials > opt->max_num_trials ||
      opt->min_tri_angle < 0)
    return fail(ctx, PPSFM_ERR_INVALID, "EstimateTriangulationOptions::Check failed");
  if (pb->num_images < 0 || pb->num_points < 0 || pb->num_obs < 0)
    return fail(ctx, PPSFM_ERR_INVALID, "negative size");
  for (int i = 0; i < pb->num_images; ++i) {
    const int cam = pb->image_camera[i];
    if (cam < 0 || cam >= pb->num_cameras || pb->camera_model[cam] < 0 || pb->camera_model[cam] > 10)
      return fail(ctx, PPSFM_ERR_INVALID, "image %d: missing camera or unsupported model", i);
  }
  const int T = pb->num_points;
  const int64_t O = pb->num_obs;
  if (pb->track_start[0] != 0 || pb->track_start[T] != O)
    return fail(ctx, PPSFM_ERR_INVALID, "track_start does not cover the observations");
  for (int64_t k = 0; k < O; ++k)
    if (pb->obs_image[k] < 0 || pb->obs_image[k] >= pb->num_images)
      return fail(ctx, PPSFM_ERR_INVALID, "observation %lld references a missing image", (long long)k);
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  // constructor cap on max_num_trials (src/optim/ransac.h:144-156)
  TriOpts o;
  o.min_tri_angle = opt->min_tri_angle;
  o.max_residual = opt->max_error * opt->max_error;
  o.confidence = opt->confidence;
  o.multiplier = opt->dyn_num_trials_multiplier;
  o.min_num_trials = opt->min_num_trials;
  const size_t kNumSamples = 100000;
  const size_t cap = HostNumTrials(static_cast<size_t>(opt->min_inlier_ratio * kNumSamples),
                                   kNumSamples, opt->confidence, opt->dyn_num_trials_multiplier);
  o.max_num_trials = std::min<unsigned long long>(opt->max_num_trials, cap);
  o.residual_type = opt->residual_type;
  o.exhaustive_threshold = opt->exhaustive_threshold;
  std::vector<unsigned long long> table((size_t)(kTableN + 1) * (kTableN + 1), 0);
  for (int n = 1; n <= kTableN; ++n)
    for (int k = 0; k <= n; ++k)
      table[(size_t)n * (kTableN + 1) + k] =
          HostNumTrials((size_t)k, (size_t)n, opt->confidence, opt->dyn_num_trials_multiplier);
  std::vector<void*> bufs;
  cudaError_t e = cudaSuccess;
  auto up = [&](const void* host, size_t bytes) -> void* {
    void* p = nullptr;
    if (e != cudaSuccess) return nullptr;
    e = cudaMallocAsync(&p, bytes ? bytes : 16, s);
    if (e != cudaSuccess) return nullptr;
    bufs.push_back(p);
    if (host && bytes) e = cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, s);
    return p;
  };
  const int C = pb->num_images;
  TriDev d;
  d.C = C; d.T = T; d.num_cameras = pb->num_cameras; d.O = O;
  const double* dq = (const double*)up(pb->qvecs, sizeof(double) * 4 * C);
  const double* dt = (const double*)up(pb->tvecs, sizeof(double) * 3 * C);
  double* proj = (double*)up(nullptr, sizeof(double) * 12 * C);
  double* centers = (double*)up(nullptr, sizeof(double) * 3 * C);
  d.proj = proj; d.centers = centers;
  d.cam_params = (const double*)up(pb->camera_params, sizeof(double) * 12 * pb->num_cameras);
  d.obs_line = (const double*)up(pb->obs_line, sizeof(double) * 3 * O);
  d.img_cam = (const int*)up(pb->image_camera, sizeof(int) * C);
  d.cam_model = (const int*)up(pb->camera_model, sizeof(int) * pb->num_cameras);
  d.cam_w = (const int*)up(pb->camera_width, sizeof(int) * pb->num_cameras);
  d.cam_h = (const int*)up(pb->camera_height, sizeof(int) * pb->num_cameras);
  d.obs_image = (const int*)up(pb->obs_image, sizeof(int) * O);
  d.track_start = (const int64_t*)up(pb->track_start, sizeof(int64_t) * ((size_t)T + 1));
  d.trial_table = (const unsigned long long*)up(table.data(), sizeof(unsigned long long) * table.size());
  d.xyz = (double*)up(nullptr, sizeof(double) * 3 * (size_t)T);
  d.success = (uint8_t*)up(nullptr, (size_t)T);
  d.inlier_mask = (uint8_t*)up(nullptr, (size_t)O);
  d.num_trials = (unsigned*)up(nullptr, sizeof(unsigned) * (size_t)T);
  int rc = PPSFM_OK;
  auto body = [&]() -> int {
    PPSFM_CUDA(ctx, e);
    if (C > 0) tri_pose_kernel<<<(C + 127) / 128, 128, 0, s>>>(C, dq, dt, proj, centers);
    if (T > 0) {
      PPSFM_CUDA(ctx, cudaMemsetAsync(d.xyz, 0, sizeof(double) * 3 * (size_t)T, s));
      tri_ransac_kernel<<<(T + 127) / 128, 128, 0, s>>>(d, o);
      PPSFM_CUDA(ctx, cudaMemcpyAsync(xyz, d.xyz, sizeof(double) * 3 * (size_t)T, cudaMemcpyDeviceToHost, s));
      PPSFM_CUDA(ctx, cudaMemcpyAsync(success, d.success, (size_t)T, cudaMemcpyDeviceToHost, s));
      if (num_trials)
        PPSFM_CUDA(ctx, cudaMemcpyAsync(num_trials, d.num_trials, sizeof(unsigned) * (size_t)T,
                                        cudaMemcpyDeviceToHost, s));
    }
    if (O > 0)
      PPSFM_CUDA(ctx, cudaMemcpyAsync(inlier_mask, d.inlier_mask, (size_t)O, cudaMemcpyDeviceToHost, s));
    PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
    PPSFM_CUDA(ctx, cudaGetLastError());
    return PPSFM_OK;
  };
  rc = body();
  for (void* p : bufs) cudaFreeAsync(p, s);
  return rc;
}

}  // extern "C"
