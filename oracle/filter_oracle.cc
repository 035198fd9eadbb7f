// filter_oracle.cc — TEST INFRASTRUCTURE: CPU restatement of the reference's post-BA filters.
//   Reconstruction::FilterPoints3DWithLargeReprojectionError  src/base/reconstruction.cc:650-719
//   Reconstruction::FilterPoints3DWithSmallTriangulationAngle src/base/reconstruction.cc:594-648
//   Reconstruction::FilterObservationsWithNegativeDepth       src/base/reconstruction.cc:442-460
//   CalculateSquaredLineReprojectionError                     src/base/projection.cc:162-203
//   CalculateTriangulationAngle                               src/base/triangulation.cc:59-82
//   ProjectionCenterFromPose / QuaternionToRotationMatrix     src/base/pose.cc:46-62, 94-101
//   CameraModel::WorldToImage                                 src/base/camera_models.h:615-904
// The reference's hash-map Reconstruction is replaced by a track-major SoA view (the loops visit a
// point's track in Track::Elements() order, as here).  PINNED against the reference's own
// sources: oracle/build_ref.sh compiles src/base/reconstruction.cc and the classes it uses into
// oracle/_ref/libref_filter.so; tests/test_ref_filters.py requires num_filtered, the deleted
// observations / points and Point3D::Error() to be identical (the projection centres differ in
// the last bits: Eigen's quaternion-vector product there, -R^T t here).
#include "camera_models_ext.h"
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

namespace {

struct FilterProblem {  // same layout as ppsfm_filter_problem
  int32_t num_images;
  const double* qvecs;
  const double* tvecs;
  const int32_t* image_camera;
  int32_t num_cameras;
  const int32_t* camera_model;
  const double* camera_params;
  const int32_t* camera_width;
  const int32_t* camera_height;
  int32_t num_points;
  const double* points;
  const int64_t* track_start;
  int64_t num_obs;
  const int32_t* obs_image;
  const double* obs_line;
  const uint8_t* obs_aligned;
};

void RotationOf(const double* qv, double R[9]) {
  const double n = std::sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2] + qv[3] * qv[3]);
  const double w = qv[0] / n, x = qv[1] / n, y = qv[2] / n, z = qv[3] / n;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

void WorldToImage(int model, const double* p, double u, double v, double* x, double* y) {
  if (model >= 5) {  // the fisheye / FOV / full-OpenCV / thin-prism models (camera_models_ext.h)
    orc_cam::WorldToImageExt(model, p, u, v, x, y, [](double c) { return c; },
                             [&](int k) { return p[k]; });
    return;
  }
  switch (model) {
    case 0: *x = p[0] * u + p[1]; *y = p[0] * v + p[2]; break;           // SIMPLE_PINHOLE
    case 1: *x = p[0] * u + p[2]; *y = p[1] * v + p[3]; break;           // PINHOLE
    case 2:
    case 3: {                                                            // SIMPLE_RADIAL, RADIAL
      const double k1 = p[3], k2 = (model == 3) ? p[4] : 0.0;
      const double u2 = u * u, v2 = v * v, r2 = u2 + v2;
      const double radial = k1 * r2 + k2 * r2 * r2;
      *x = p[0] * (u + u * radial) + p[1];
      *y = p[0] * (v + v * radial) + p[2];
      break;
    }
    default: {                                                           // OPENCV
      const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
      const double u2 = u * u, uv = u * v, v2 = v * v, r2 = u2 + v2;
      const double radial = k1 * r2 + k2 * r2 * r2;
      const double du = u * radial + 2.0 * p1 * uv + p2 * (r2 + 2.0 * u2);
      const double dv = v * radial + 2.0 * p2 * uv + p1 * (r2 + 2.0 * v2);
      *x = p[0] * (u + du) + p[2];
      *y = p[1] * (v + dv) + p[3];
    }
  }
}

double SquaredLineError(const FilterProblem& pb, int img, const double* l, const double* X) {
  double R[9];
  RotationOf(pb.qvecs + 4 * (size_t)img, R);
  const double* t = pb.tvecs + 3 * (size_t)img;
  const double pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  if (pz < DBL_EPSILON) return DBL_MAX;
  const double px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  const double py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  const double inv = 1.0 / pz;
  const double u = inv * px, v = inv * py;
  const double alpha = l[0] * u + l[1] * v + l[2];
  const double lu = u - l[0] * alpha, lv = v - l[1] * alpha;
  const int cam = pb.image_camera[img];
  const double* prm = pb.camera_params + 12 * (size_t)cam;
  double x1, y1, x2, y2;
  WorldToImage(pb.camera_model[cam], prm, u, v, &x1, &y1);
  if (!(x1 >= 0.0 && x1 < (double)pb.camera_width[cam] && y1 >= 0.0 &&
        y1 < (double)pb.camera_height[cam]))
    return DBL_MAX;
  WorldToImage(pb.camera_model[cam], prm, lu, lv, &x2, &y2);
  const double dx = x1 - x2, dy = y1 - y2;
  return dx * dx + dy * dy;
}

double TriangulationAngle(const double* c1, const double* c2, const double* X) {
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

}  // namespace

extern "C" {

int orc_filter_points3d(const FilterProblem* pbp, double max_reproj_error, double min_tri_angle_deg,
                        uint8_t* obs_deleted, uint8_t* point_deleted, double* point_error,
                        uint64_t* num_filtered, double* sq_errors /* [O] or null */) {
  const FilterProblem& pb = *pbp;
  const double max_sq = max_reproj_error * max_reproj_error;
  const double min_rad = min_tri_angle_deg * 0.0174532925199432954743716805978692718781530857086181640625;
  std::vector<double> centers(3 * (size_t)pb.num_images);
  for (int i = 0; i < pb.num_images; ++i) {
    double R[9];
    RotationOf(pb.qvecs + 4 * (size_t)i, R);
    const double* t = pb.tvecs + 3 * (size_t)i;
    for (int k = 0; k < 3; ++k) centers[3 * (size_t)i + k] = -(R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2]);
  }
  uint64_t total = 0;
  for (int64_t k = 0; k < pb.num_obs; ++k) obs_deleted[k] = 0;
  for (int p = 0; p < pb.num_points; ++p) {
    const int64_t k0 = pb.track_start[p], k1 = pb.track_start[p + 1], len = k1 - k0;
    point_deleted[p] = 0;
    if (len == 0) continue;
    const double* X = pb.points + 3 * (size_t)p;
    bool deleted = false, have_non_aligned = false;
    for (int64_t k = k0; k < k1; ++k) have_non_aligned |= pb.obs_aligned[k] == 0;
    if (!have_non_aligned || len < 3) {
      deleted = true;
      total += (uint64_t)len;
    } else {
      double sum = 0.0;
      int64_t nd = 0;
      for (int64_t k = k0; k < k1; ++k) {
        const double e = SquaredLineError(pb, pb.obs_image[k], pb.obs_line + 3 * (size_t)k, X);
        if (sq_errors) sq_errors[k] = e;
        if (e > max_sq) { obs_deleted[k] = 1; ++nd; } else { sum += std::sqrt(e); }
      }
      if (nd >= len - 3) {
        deleted = true;
        total += (uint64_t)len;
      } else {
        total += (uint64_t)nd;
        point_error[p] = sum / (double)(len - nd);  // after DeleteObservation (:706-713)
      }
    }
    if (!deleted) {
      bool keep = false;
      for (int64_t i1 = k0; i1 < k1 && !keep; ++i1) {
        if (obs_deleted[i1]) continue;
        for (int64_t i2 = k0; i2 < i1; ++i2) {
          if (obs_deleted[i2]) continue;
          if (TriangulationAngle(&centers[3 * (size_t)pb.obs_image[i1]],
                                 &centers[3 * (size_t)pb.obs_image[i2]], X) >= min_rad) {
            keep = true;
            break;
          }
        }
      }
      if (!keep) {
        deleted = true;
        total += 1;
      }
    }
    if (deleted)
      for (int64_t k = k0; k < k1; ++k) obs_deleted[k] = 1;
    point_deleted[p] = deleted ? 1 : 0;
  }
  *num_filtered = total;
  return 0;
}

// reconstruction.cc:442-460 with DeleteObservation (:255-275) simulated literally: images in
// index order, the observations of an image in input order; an observation whose point is gone
// "has no point" and is skipped.
int orc_filter_negative_depth(const FilterProblem* pbp, uint8_t* obs_deleted,
                              uint8_t* point_deleted, uint64_t* num_filtered) {
  const FilterProblem& pb = *pbp;
  std::vector<int> obs_point((size_t)pb.num_obs);
  std::vector<int64_t> length((size_t)pb.num_points);
  std::vector<uint8_t> alive((size_t)pb.num_points, 1);
  for (int p = 0; p < pb.num_points; ++p) {
    length[p] = pb.track_start[p + 1] - pb.track_start[p];
    for (int64_t k = pb.track_start[p]; k < pb.track_start[p + 1]; ++k) obs_point[k] = p;
  }
  std::vector<std::vector<int64_t>> by_image((size_t)pb.num_images);
  for (int64_t k = 0; k < pb.num_obs; ++k) by_image[pb.obs_image[k]].push_back(k);
  for (int64_t k = 0; k < pb.num_obs; ++k) obs_deleted[k] = 0;
  uint64_t total = 0;
  for (int img = 0; img < pb.num_images; ++img) {
    double R[9];
    RotationOf(pb.qvecs + 4 * (size_t)img, R);
    for (const int64_t k : by_image[img]) {
      const int p = obs_point[k];
      if (!alive[p] || obs_deleted[k]) continue;  // !line.HasPoint3D()
      const double* X = pb.points + 3 * (size_t)p;
      const double pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + pb.tvecs[3 * (size_t)img + 2];
      if (pz >= DBL_EPSILON) continue;            // HasPointPositiveDepth
      // DeleteObservation
      if (length[p] <= 3) {
        alive[p] = 0;                             // DeletePoint3D: every line of the track reset
        for (int64_t j = pb.track_start[p]; j < pb.track_start[p + 1]; ++j) obs_deleted[j] = 1;
      } else {
        length[p] -= 1;
        obs_deleted[k] = 1;
      }
      total += 1;
    }
  }
  if (point_deleted)
    for (int p = 0; p < pb.num_points; ++p) point_deleted[p] = alive[p] ? 0 : 1;
  *num_filtered = total;
  return 0;
}

}  // extern "C"
