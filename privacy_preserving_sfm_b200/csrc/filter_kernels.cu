// filter_kernels.cu — observation / point filters that run after every bundle adjustment
// (SURVEY.md §8 f2).  Replaces, on a track-major SoA view of the reconstruction,
//   Reconstruction::FilterPoints3D                         src/base/reconstruction.cc:425-440
//     FilterPoints3DWithLargeReprojectionError             :650-719
//     FilterPoints3DWithSmallTriangulationAngle            :594-648
//   Reconstruction::FilterObservationsWithNegativeDepth    :442-460
// with CalculateSquaredLineReprojectionError (src/base/projection.cc:162-203: cheirality and
// in-image test), CalculateTriangulationAngle (src/base/triangulation.cc:59-82) and
// ProjectionCenterFromPose (src/base/pose.cc:94-101).
// One thread per 3-D point walks its track in order, so the per-point error sums round exactly
// like the reference's loops; the TU is compiled --fmad=false (the reference build has no FMA),
// which makes the squared errors and therefore every keep / delete decision bit-identical to the
// CPU path (the arccosine of the angle test is the only libm-dependent value).
#include <cfloat>
#include <cmath>
#include <vector>

#include "camera_models.cuh"
#include "common.h"

namespace ppsfm {

namespace {

struct FilterDev {
  int C, P, num_cameras;
  int64_t O;
  const double *q, *t, *X, *cam_params, *obs_line;
  const int *img_cam, *cam_model, *cam_w, *cam_h, *obs_image;
  const int64_t* track_start;
  const uint8_t* obs_aligned;
  double* centers;  // [C][3] projection centres
  uint8_t *obs_deleted, *point_deleted;
  double* point_error;
  unsigned long long* num_filtered;
};

// rotation matrix of the (unnormalised-safe) quaternion, Eigen convention (base/pose.cc:46-62)
__device__ __forceinline__ void rotation_of(const double* qv, double R[9]) {
  const double n = sqrt(qv[0] * qv[0] + qv[1] * qv[1] + qv[2] * qv[2] + qv[3] * qv[3]);
  const double w = qv[0] / n, x = qv[1] / n, y = qv[2] / n, z = qv[3] / n;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

__global__ void filter_centers_kernel(FilterDev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.C) return;
  double R[9];
  rotation_of(d.q + 4 * (size_t)i, R);
  const double* t = d.t + 3 * (size_t)i;
  // C = -R^T t
  for (int k = 0; k < 3; ++k)
    d.centers[3 * (size_t)i + k] = -(R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2]);
}

// CalculateSquaredLineReprojectionError (projection.cc:162-203)
__device__ __forceinline__ double squared_line_error(const FilterDev& d, int img, const double* l,
                                                     const double* X) {
  double R[9];
  rotation_of(d.q + 4 * (size_t)img, R);
  const double* t = d.t + 3 * (size_t)img;
  const double pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
  if (pz < DBL_EPSILON) return DBL_MAX;
  const double px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
  const double py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
  const double inv = 1.0 / pz;
  const double u = inv * px, v = inv * py;
  const double alpha = l[0] * u + l[1] * v + l[2];
  const double lu = u - l[0] * alpha, lv = v - l[1] * alpha;
  const int cam = d.img_cam[img];
  const double* prm = d.cam_params + 12 * (size_t)cam;
  const int model = d.cam_model[cam];
  double x1, y1, x2, y2, j0, j1, j2, j3;
  world_to_image<false>(model, prm, u, v, x1, y1, j0, j1, j2, j3);
  if (!(x1 >= 0.0 && x1 < (double)d.cam_w[cam] && y1 >= 0.0 && y1 < (double)d.cam_h[cam]))
    return DBL_MAX;
  world_to_image<false>(model, prm, lu, lv, x2, y2, j0, j1, j2, j3);
  const double dx = x1 - x2, dy = y1 - y2;
  return dx * dx + dy * dy;
}

// CalculateTriangulationAngle (triangulation.cc:59-82)
__device__ __forceinline__ double triangulation_angle(const double* c1, const double* c2,
                                                      const double* X) {
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

__global__ void __launch_bounds__(128)
filter_points_kernel(FilterDev d, double max_sq_error, double min_tri_angle_rad) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.P) return;
  const int64_t k0 = d.track_start[p], k1 = d.track_start[p + 1];
  const int64_t len = k1 - k0;
  if (len == 0) {  // point without a track: nothing to do (it is not in point3D_ids)
    d.point_deleted[p] = 0;
    return;
  }
  const double* X = d.X + 3 * (size_t)p;
  unsigned long long filtered = 0;
  bool deleted = false;
  bool have_non_aligned = false;
  for (int64_t k = k0; k < k1; ++k) have_non_aligned |= d.obs_aligned[k] == 0;
  if (!have_non_aligned || len < 3) {
    deleted = true;
    filtered = (unsigned long long)len;
  } else {
    double sum = 0.0;
    int64_t nd = 0;
    for (int64_t k = k0; k < k1; ++k) {
      const double e = squared_line_error(d, d.obs_image[k], d.obs_line + 3 * (size_t)k, X);
      const bool del = e > max_sq_error;
      d.obs_deleted[k] = del ? 1 : 0;
      if (del) ++nd; else sum += sqrt(e);
    }
    if (nd >= len - 3) {
      deleted = true;
      filtered = (unsigned long long)len;
    } else {
      filtered = (unsigned long long)nd;
      // SetError after the DeleteObservation calls: the track is already shorter (:706-713)
      d.point_error[p] = sum / (double)(len - nd);
    }
  }
  if (!deleted) {
    // small triangulation angle: keep the point if any pair of remaining views is wide enough
    bool keep = false;
    for (int64_t i1 = k0; i1 < k1 && !keep; ++i1) {
      if (d.obs_deleted[i1]) continue;
      const double* c1 = d.centers + 3 * (size_t)d.obs_image[i1];
      for (int64_t i2 = k0; i2 < i1; ++i2) {
        if (d.obs_deleted[i2]) continue;
        if (triangulation_angle(c1, d.centers + 3 * (size_t)d.obs_image[i2], X) >=
            min_tri_angle_rad) {
          keep = true;
          break;
        }
      }
    }
    if (!keep) {
      deleted = true;
      filtered += 1;
    }
  }
  if (deleted)
    for (int64_t k = k0; k < k1; ++k) d.obs_deleted[k] = 1;
  d.point_deleted[p] = deleted ? 1 : 0;
  if (filtered) atomicAdd(d.num_filtered, filtered);
}

// FilterObservationsWithNegativeDepth (reconstruction.cc:442-460): thread per point.
// Every negative-depth observation goes through DeleteObservation (:255-275), which deletes the
// whole point when its track is down to <= 3 elements; the later observations of that point no
// longer "have a point" and are neither visited nor counted.  Per point with track length len and
// n negative-depth observations the sequential loop therefore counts min(n, max(1, len - 2))
// deletions, and the point dies iff n >= max(1, len - 2) -- whatever the image order.
__global__ void __launch_bounds__(128) filter_depth_kernel(FilterDev d) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= d.P) return;
  const int64_t k0 = d.track_start[p], k1 = d.track_start[p + 1];
  const int64_t len = k1 - k0;
  d.point_deleted[p] = 0;
  if (len == 0) return;
  const double* X = d.X + 3 * (size_t)p;
  int64_t n = 0;
  for (int64_t k = k0; k < k1; ++k) {
    double R[9];
    const int img = d.obs_image[k];
    rotation_of(d.q + 4 * (size_t)img, R);
    const double pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + d.t[3 * (size_t)img + 2];
    const bool del = !(pz >= DBL_EPSILON);
    d.obs_deleted[k] = del ? 1 : 0;
    n += del ? 1 : 0;
  }
  if (n == 0) return;
  const int64_t fatal = len - 2 > 1 ? len - 2 : 1;
  if (n >= fatal) {
    for (int64_t k = k0; k < k1; ++k) d.obs_deleted[k] = 1;
    d.point_deleted[p] = 1;
    n = fatal;
  }
  atomicAdd(d.num_filtered, (unsigned long long)n);
}

struct Uploader {
  ppsfm_ctx* ctx;
  std::vector<void*> bufs;
  cudaError_t err = cudaSuccess;
  template <typename T>
  T* up(const T* host, size_t count) {
    T* p = nullptr;
    if (err != cudaSuccess) return nullptr;
    err = cudaMallocAsync((void**)&p, (count ? count : 1) * sizeof(T), ctx->stream);
    if (err != cudaSuccess) return nullptr;
    bufs.push_back(p);
    if (count && host)
      err = cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    return p;
  }
  ~Uploader() {
    for (void* p : bufs) cudaFreeAsync(p, ctx->stream);
  }
};

int validate(ppsfm_ctx* ctx, const ppsfm_filter_problem* pb) {
  if (!ctx || !pb) return PPSFM_ERR_INVALID;
  if (pb->num_images < 0 || pb->num_points < 0 || pb->num_obs < 0 || pb->num_cameras < 0)
    return fail(ctx, PPSFM_ERR_INVALID, "negative size");
  for (int i = 0; i < pb->num_images; ++i) {
    const int cam = pb->image_camera[i];
    if (cam < 0 || cam >= pb->num_cameras)
      return fail(ctx, PPSFM_ERR_INVALID, "image %d references a missing camera", i);
    if (pb->camera_model[cam] < 0 || pb->camera_model[cam] > 10)
      return fail(ctx, PPSFM_ERR_INVALID, "camera model %d not supported", pb->camera_model[cam]);
  }
  if (pb->track_start[0] != 0 || pb->track_start[pb->num_points] != pb->num_obs)
    return fail(ctx, PPSFM_ERR_INVALID, "track_start does not cover the observations");
  for (int64_t k = 0; k < pb->num_obs; ++k) {
    if (pb->obs_image[k] < 0 || pb->obs_image[k] >= pb->num_images)
      return fail(ctx, PPSFM_ERR_INVALID, "observation %lld references a missing image",
                  (long long)k);
    const double* l = pb->obs_line + 3 * k;
    const double n2 = l[0] * l[0] + l[1] * l[1];
    if (!(n2 > 0.999998 && n2 < 1.000002) && std::fabs(std::sqrt(n2) - 1.0) > 1e-6)
      return fail(ctx, PPSFM_ERR_INVALID, "observation %lld: line normal is not unit length",
                  (long long)k);  // CHECK_NEAR, projection.cc:166
  }
  return PPSFM_OK;
}

int run(ppsfm_ctx* ctx, const ppsfm_filter_problem* pb, bool depth_only, double max_reproj_error,
        double min_tri_angle_deg, uint8_t* obs_deleted, uint8_t* point_deleted,
        double* point_error, size_t* num_filtered) {
  int rc = validate(ctx, pb);
  if (rc != PPSFM_OK) return rc;
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  const int C = pb->num_images, P = pb->num_points;
  const int64_t O = pb->num_obs;
  Uploader u{ctx};
  FilterDev d;
  d.C = C; d.P = P; d.num_cameras = pb->num_cameras; d.O = O;
  d.q = u.up(pb->qvecs, 4 * (size_t)C);
  d.t = u.up(pb->tvecs, 3 * (size_t)C);
  d.X = u.up(pb->points, 3 * (size_t)P);
  d.cam_params = u.up(pb->camera_params, 12 * (size_t)pb->num_cameras);
  d.obs_line = u.up(pb->obs_line, 3 * (size_t)O);
  d.img_cam = u.up(pb->image_camera, (size_t)C);
  d.cam_model = u.up(pb->camera_model, (size_t)pb->num_cameras);
  d.cam_w = u.up(pb->camera_width, (size_t)pb->num_cameras);
  d.cam_h = u.up(pb->camera_height, (size_t)pb->num_cameras);
  d.obs_image = u.up(pb->obs_image, (size_t)O);
  d.track_start = u.up(pb->track_start, (size_t)P + 1);
  d.obs_aligned = u.up(pb->obs_aligned, (size_t)O);
  d.centers = u.up((const double*)nullptr, 3 * (size_t)C);
  d.obs_deleted = u.up((const uint8_t*)nullptr, (size_t)O);
  d.point_deleted = u.up((const uint8_t*)nullptr, (size_t)P);
  d.point_error = u.up(point_error, (size_t)P);
  d.num_filtered = u.up((const unsigned long long*)nullptr, 1);
  PPSFM_CUDA(ctx, u.err);
  PPSFM_CUDA(ctx, cudaMemsetAsync(d.obs_deleted, 0, (size_t)(O ? O : 1), s));
  PPSFM_CUDA(ctx, cudaMemsetAsync(d.num_filtered, 0, sizeof(unsigned long long), s));
  if (depth_only) {
    if (P > 0) filter_depth_kernel<<<(P + 127) / 128, 128, 0, s>>>(d);
  } else {
    if (C > 0) filter_centers_kernel<<<(C + 127) / 128, 128, 0, s>>>(d);
    if (P > 0)
      filter_points_kernel<<<(P + 127) / 128, 128, 0, s>>>(
          d, max_reproj_error * max_reproj_error, min_tri_angle_deg * 0.0174532925199432954743716805978692718781530857086181640625);
  }
  unsigned long long nf = 0;
  if (obs_deleted && O > 0)
    PPSFM_CUDA(ctx, cudaMemcpyAsync(obs_deleted, d.obs_deleted, (size_t)O, cudaMemcpyDeviceToHost, s));
  if (point_deleted && P > 0)
    PPSFM_CUDA(ctx, cudaMemcpyAsync(point_deleted, d.point_deleted, (size_t)P, cudaMemcpyDeviceToHost, s));
  if (!depth_only) {
    if (point_error && P > 0)
      PPSFM_CUDA(ctx, cudaMemcpyAsync(point_error, d.point_error, sizeof(double) * (size_t)P,
                                      cudaMemcpyDeviceToHost, s));
  }
  PPSFM_CUDA(ctx, cudaMemcpyAsync(&nf, d.num_filtered, sizeof(nf), cudaMemcpyDeviceToHost, s));
  PPSFM_CUDA(ctx, cudaStreamSynchronize(s));
  PPSFM_CUDA(ctx, cudaGetLastError());
  if (num_filtered) *num_filtered = (size_t)nf;
  return PPSFM_OK;
}

}  // namespace
}  // namespace ppsfm

extern "C" {

int ppsfm_filter_points3d(ppsfm_ctx* ctx, const ppsfm_filter_problem* problem,
                          double max_reproj_error, double min_tri_angle_deg, uint8_t* obs_deleted,
                          uint8_t* point_deleted, double* point_error, size_t* num_filtered) {
  return ppsfm::run(ctx, problem, false, max_reproj_error, min_tri_angle_deg, obs_deleted,
                    point_deleted, point_error, num_filtered);
}

int ppsfm_filter_observations_with_negative_depth(ppsfm_ctx* ctx,
                                                  const ppsfm_filter_problem* problem,
                                                  uint8_t* obs_deleted, uint8_t* point_deleted,
                                                  size_t* num_filtered) {
  return ppsfm::run(ctx, problem, true, 0.0, 0.0, obs_deleted, point_deleted, nullptr,
                    num_filtered);
}

}  // extern "C"
