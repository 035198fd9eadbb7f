// triangulation_oracle.cc — TEST INFRASTRUCTURE: CPU restatement of the reference's robust line
// triangulation of ONE track at a time.
//   EstimateTriangulation / TriangulationEstimator   src/estimators/triangulation.cc:55-149
//   TriangulateMultiViewPoint                         src/base/triangulation.cc:41-57
//   CalculateTriangulationAngle                       src/base/triangulation.cc:59-82
//   CalculateNormalizedLineAngularError,
//   CalculateSquaredLineReprojectionError             src/base/projection.cc:162-203, 241-260
//   LORANSAC::Estimate                                src/optim/loransac.h:91-234
//   CombinationSampler, NChooseK, NextCombination     src/optim/combination_sampler.cc:41-70,
//                                                     src/util/math.cc:36-42, math.h:140-176
//   RANSAC ctor cap / ComputeNumTrials                src/optim/ransac.h:144-176
// Eigen::JacobiSVD is replaced by a one-sided Jacobi iteration for the null vector of the n x 4
// system (eigen_restated.h; parity unpinned at that Eigen boundary).  Everything else is PINNED
// against the reference's own sources: oracle/build_ref.sh compiles triangulation.cc,
// base/triangulation.cc, base/projection.cc, camera.cc, loransac.h, combination_sampler.cc and
// math.cc from /root/reference (oracle/_ref/libref_tri.so) and tests/test_ref_triangulation.py
// requires success flags, points, inlier masks and trial counts to be bit-identical.
#include "camera_models_ext.h"
#include "eigen_restated.h"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <limits>
#include <numeric>
#include <vector>

namespace {

struct Problem {  // same layout as ppsfm_filter_problem
  int32_t num_images; const double* qvecs; const double* tvecs; const int32_t* image_camera;
  int32_t num_cameras; const int32_t* camera_model; const double* camera_params;
  const int32_t* camera_width; const int32_t* camera_height;
  int32_t num_points; const double* points; const int64_t* track_start;
  int64_t num_obs; const int32_t* obs_image; const double* obs_line; const uint8_t* obs_aligned;
};
struct Options {  // same layout as ppsfm_triangulation_options
  double min_tri_angle; int32_t residual_type; double max_error, min_inlier_ratio, confidence,
      dyn_num_trials_multiplier; uint64_t min_num_trials, max_num_trials; int32_t exhaustive_threshold;
};

struct Pose { double P[12]; double c[3]; };

Pose MakePose(const double* qv, const double* t) {
  const double n = std::sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2] + qv[3] * qv[3]);
  const double w = qv[0] / n, x = qv[1] / n, y = qv[2] / n, z = qv[3] / n;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  const double R[9] = {1.0 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.0 - (txx + tzz),
                       tyz - twx, txz - twy, tyz + twx, 1.0 - (txx + tyy)};
  Pose p;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) p.P[4 * r + c] = R[3 * r + c];
    p.P[4 * r + 3] = t[r];
  }
  for (int k = 0; k < 3; ++k) p.c[k] = -(R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2]);
  return p;
}

void WorldToImage(int model, const double* p, double u, double v, double* x, double* y) {
  if (model >= 5) {  // the fisheye / FOV / full-OpenCV / thin-prism models (camera_models_ext.h)
    orc_cam::WorldToImageExt(model, p, u, v, x, y, [](double c) { return c; },
                             [&](int k) { return p[k]; });
    return;
  }
  switch (model) {
    case 0: *x = p[0] * u + p[1]; *y = p[0] * v + p[2]; break;
    case 1: *x = p[0] * u + p[2]; *y = p[1] * v + p[3]; break;
    case 2:
    case 3: {
      const double k1 = p[3], k2 = (model == 3) ? p[4] : 0.0;
      const double u2 = u * u, v2 = v * v, r2 = u2 + v2, radial = k1 * r2 + k2 * r2 * r2;
      *x = p[0] * (u + u * radial) + p[1];
      *y = p[0] * (v + v * radial) + p[2];
      break;
    }
    default: {
      const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
      const double u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2, radial = k1 * r2 + k2 * r2 * r2;
      const double du = u * radial + 2.0 * p1 * uv + p2 * (r2 + 2.0 * u2);
      const double dv = v * radial + 2.0 * p2 * uv + p1 * (r2 + 2.0 * v2);
      *x = p[0] * (u + du) + p[2];
      *y = p[1] * (v + dv) + p[3];
    }
  }
}

// TriangulateMultiViewPoint: smallest right singular vector of the n x 4 system (one-sided Jacobi)
void MultiViewPoint(const std::vector<const double*>& lines, const std::vector<const Pose*>& poses,
                    double X[3]) {
  const int n = (int)lines.size();
  std::vector<double> W(4 * (size_t)n);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 4; ++c)
      W[4 * i + c] = lines[i][0] * poses[i]->P[c] + lines[i][1] * poses[i]->P[4 + c] +
                     lines[i][2] * poses[i]->P[8 + c];
  double v[4];
  eigen_restated::NullVectorNx4(W, n, v);  // JacobiSVD(A, ComputeFullV).matrixV().col(3)
  X[0] = v[0] / v[3]; X[1] = v[1] / v[3]; X[2] = v[2] / v[3];  // hnormalized()
}

double TriAngle(const double* c1, const double* c2, const double* X) {
  double b2 = 0, r1 = 0, r2 = 0;
  for (int k = 0; k < 3; ++k) {
    b2 += (c1[k] - c2[k]) * (c1[k] - c2[k]);
    r1 += (X[k] - c1[k]) * (X[k] - c1[k]);
    r2 += (X[k] - c2[k]) * (X[k] - c2[k]);
  }
  const double den = 2.0 * std::sqrt(r1 * r2);
  if (den == 0.0) return 0.0;
  const double angle = std::fabs(std::acos((r1 + r2 - b2) / den));
  return std::fmin(angle, M_PI - angle);
}

struct Track {
  const Problem* pb;
  const Options* opt;
  std::vector<Pose> pose;        // per observation
  std::vector<const double*> line;
  std::vector<int> cam;

  double Residual(size_t i, const double* X) const {
    const double* P = pose[i].P;
    const double* l = line[i];
    const double r0 = P[0] * X[0] + P[1] * X[1] + P[2] * X[2] + P[3];
    const double r1 = P[4] * X[0] + P[5] * X[1] + P[6] * X[2] + P[7];
    const double r2 = P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11];
    const double* prm = pb->camera_params + 12 * (size_t)cam[i];
    const int model = pb->camera_model[cam[i]];
    const double W = pb->camera_width[cam[i]], H = pb->camera_height[cam[i]];
    double x1, y1;
    if (opt->residual_type == 0) {
      if (r2 < 0) return DBL_MAX;
      WorldToImage(model, prm, r0 / r2, r1 / r2, &x1, &y1);
      if (x1 < 0 || x1 >= W || y1 < 0 || y1 >= H) return DBL_MAX;
      const double nl = std::sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
      const double nr = std::sqrt(r0 * r0 + r1 * r1 + r2 * r2);
      const double dot = (l[0] / nl) * (r0 / nr) + (l[1] / nl) * (r1 / nr) + (l[2] / nl) * (r2 / nr);
      const double e = std::fabs(M_PI_2 - std::acos(std::fabs(dot)));
      return e * e;
    }
    if (r2 < DBL_EPSILON) return DBL_MAX;
    const double inv = 1.0 / r2, u = inv * r0, v = inv * r1;
    const double alpha = l[0] * u + l[1] * v + l[2];
    WorldToImage(model, prm, u, v, &x1, &y1);
    if (!(x1 >= 0.0 && x1 < W && y1 >= 0.0 && y1 < H)) return DBL_MAX;
    double x2, y2;
    WorldToImage(model, prm, u - l[0] * alpha, v - l[1] * alpha, &x2, &y2);
    return (x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2);
  }

  // TriangulationEstimator::Estimate on a subset
  bool Estimate(const std::vector<size_t>& idx, double X[3]) const {
    std::vector<const double*> ls;
    std::vector<const Pose*> ps;
    for (size_t i : idx) { ls.push_back(line[i]); ps.push_back(&pose[i]); }
    MultiViewPoint(ls, ps, X);
    for (const Pose* p : ps)
      if (!(p->P[8] * X[0] + p->P[9] * X[1] + p->P[10] * X[2] + p->P[11] >= DBL_EPSILON)) return false;
    for (size_t i = 0; i < ps.size(); ++i)
      for (size_t j = 0; j < i; ++j)
        if (TriAngle(ps[i]->c, ps[j]->c, X) >= opt->min_tri_angle) return true;
    return false;
  }
};

size_t NChooseK(size_t n, size_t k) { return k == 0 ? 1 : (n * NChooseK(n - 1, k - 1)) / k; }

size_t ComputeNumTrials(size_t num_inliers, size_t num_samples, double confidence, double multiplier) {
  const double inlier_ratio = num_inliers / static_cast<double>(num_samples);
  const double nom = 1 - confidence;
  if (nom <= 0) return std::numeric_limits<size_t>::max();
  const double denom = 1 - std::pow(inlier_ratio, 3);
  if (denom <= 0) return 1;
  return static_cast<size_t>(std::ceil(std::log(nom) / std::log(denom) * multiplier));
}

}  // namespace

extern "C" int orc_estimate_triangulation_batch(const Problem* pbp, const Options* optp, double* xyz,
                                                uint8_t* success, uint8_t* inlier_mask,
                                                uint32_t* num_trials) {
  const Problem& pb = *pbp;
  Options opt = *optp;
  const size_t kNumSamples = 100000;  // RANSAC constructor cap (ransac.h:144-156)
  opt.max_num_trials = std::min<uint64_t>(
      opt.max_num_trials, ComputeNumTrials(static_cast<size_t>(opt.min_inlier_ratio * kNumSamples),
                                           kNumSamples, opt.confidence, opt.dyn_num_trials_multiplier));
  const double max_residual = opt.max_error * opt.max_error;
  for (int t = 0; t < pb.num_points; ++t) {
    const int64_t k0 = pb.track_start[t], k1 = pb.track_start[t + 1];
    const size_t n = (size_t)(k1 - k0);
    success[t] = 0;
    if (num_trials) num_trials[t] = 0;
    for (int64_t k = k0; k < k1; ++k) inlier_mask[k] = 0;
    if (n < 3) continue;
    Track tr;
    tr.pb = &pb;
    tr.opt = &opt;
    for (int64_t k = k0; k < k1; ++k) {
      const int img = pb.obs_image[k];
      tr.pose.push_back(MakePose(pb.qvecs + 4 * (size_t)img, pb.tvecs + 3 * (size_t)img));
      tr.line.push_back(pb.obs_line + 3 * (size_t)k);
      tr.cam.push_back(pb.image_camera[img]);
    }
    uint64_t min_trials = opt.min_num_trials;
    if ((int)n <= opt.exhaustive_threshold) min_trials = NChooseK(n, 3);
    // CombinationSampler
    std::vector<size_t> total(n);
    std::iota(total.begin(), total.end(), 0);
    std::vector<bool> pick(n, false);
    std::fill(pick.begin(), pick.begin() + 3, true);
    const size_t max_trials = std::min<size_t>(opt.max_num_trials, NChooseK(n, 3));
    size_t dyn_max = max_trials;
    size_t best_num = 0;
    double best_sum = DBL_MAX, best[3] = {0, 0, 0};
    bool abort = false;
    size_t trial = 0;
    std::vector<double> res(n);
    auto evaluate = [&](const double* X, size_t* num, double* sum) {
      *num = 0;
      *sum = 0;
      for (size_t i = 0; i < n; ++i) {
        res[i] = tr.Residual(i, X);
        if (res[i] <= max_residual) { ++*num; *sum += res[i]; }
      }
    };
    for (; trial < max_trials; ++trial) {
      if (abort) { trial += 1; break; }
      std::vector<size_t> sample;
      for (size_t i = 0; i < n; ++i) if (pick[i]) sample.push_back(i);
      if (!std::prev_permutation(pick.begin(), pick.end())) {  // lexicographic next combination
        std::fill(pick.begin(), pick.end(), false);
        std::fill(pick.begin(), pick.begin() + 3, true);
      }
      double X[3];
      if (!tr.Estimate(sample, X)) continue;
      size_t num;
      double sum;
      evaluate(X, &num, &sum);
      if (num > best_num || (num == best_num && sum < best_sum)) {
        best_num = num; best_sum = sum;
        std::copy(X, X + 3, best);
        if (num > 3) {
          std::vector<size_t> inl;
          for (size_t i = 0; i < n; ++i) if (res[i] <= max_residual) inl.push_back(i);
          double XL[3];
          if (tr.Estimate(inl, XL)) {
            size_t numl;
            double suml;
            evaluate(XL, &numl, &suml);
            if (numl > best_num || (numl == best_num && suml < best_sum)) {
              best_num = numl; best_sum = suml;
              std::copy(XL, XL + 3, best);
            }
          }
        }
        dyn_max = ComputeNumTrials(best_num, n, opt.confidence, opt.dyn_num_trials_multiplier);
      }
      if (trial >= dyn_max && trial >= min_trials) abort = true;
    }
    if (num_trials) num_trials[t] = (uint32_t)trial;
    if (best_num < 3) continue;
    success[t] = 1;
    std::copy(best, best + 3, xyz + 3 * (size_t)t);
    for (size_t i = 0; i < n; ++i) inlier_mask[k0 + i] = tr.Residual(i, best) <= max_residual ? 1 : 0;
  }
  return 0;
}
