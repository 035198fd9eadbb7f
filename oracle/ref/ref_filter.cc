// ref_filter.cc — TEST INFRASTRUCTURE.  The REFERENCE's own post-BA filters,
// compiled from where they lie under /root/reference:
//   src/base/reconstruction.cc   Reconstruction::FilterPoints3D (-> ...WithLargeReprojectionError,
//                                ...WithSmallTriangulationAngle), FilterObservationsWithNegativeDepth,
//                                AddCamera / AddImage / RegisterImage / AddPoint3D,
//                                DeleteObservation / DeletePoint3D
//   src/base/{image,point3d,track,camera,camera_models,pose,projection,triangulation}.cc
// against the stand-ins of oracle/ref/shim/.  This translation unit builds a colmap::Reconstruction
// from the flat track-major problem of the C-ABI through the reference's own Add* members, runs
// the reference's filter, and reads back what it deleted.
// Built by oracle/build_ref.sh into oracle/_ref/libref_filter.so.
#include <cstdint>
#include <unordered_set>
#include <vector>

#include "base/database.h"
#include "base/reconstruction.h"

// The one definition of base/database.cc (SQLite behind it) that the compiled code names: the
// constant of Database::ImagePairToPairId, reached from Reconstruction::SetObservationAsTriangulated
// (which returns early here: no correspondence graph).  Restated from database.cc:168-169.
namespace colmap {
const size_t Database::kMaxNumImages = static_cast<size_t>(std::numeric_limits<int32_t>::max());
}

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

struct Built {
  colmap::Reconstruction rec;
  std::vector<colmap::point2D_t> obs_line_idx;  // per observation: its index among the image's lines
  std::vector<colmap::point3D_t> point_id;      // per point (0: not added, empty track)
};

void Build(const FilterProblem& pb, Built* b) {
  for (int c = 0; c < pb.num_cameras; ++c) {
    colmap::Camera cam;
    cam.SetCameraId(c + 1);
    cam.SetModelId(pb.camera_model[c]);
    cam.SetWidth(pb.camera_width[c]);
    cam.SetHeight(pb.camera_height[c]);
    cam.SetParams(std::vector<double>(pb.camera_params + 12 * c,
                                      pb.camera_params + 12 * c + cam.NumParams()));
    b->rec.AddCamera(cam);
  }
  std::vector<colmap::FeatureLines> lines(pb.num_images);
  b->obs_line_idx.resize(pb.num_obs);
  for (int64_t k = 0; k < pb.num_obs; ++k) {
    const double* l = pb.obs_line + 3 * k;
    auto& v = lines[pb.obs_image[k]];
    b->obs_line_idx[k] = static_cast<colmap::point2D_t>(v.size());
    v.emplace_back(Eigen::Vector3d(l[0], l[1], l[2]), pb.obs_aligned[k] != 0);
  }
  for (int i = 0; i < pb.num_images; ++i) {
    colmap::Image img;
    img.SetImageId(i + 1);
    img.SetCameraId(pb.image_camera[i] + 1);
    img.SetLines(lines[i]);
    img.Qvec() = Eigen::Vector4d(pb.qvecs[4 * i], pb.qvecs[4 * i + 1], pb.qvecs[4 * i + 2], pb.qvecs[4 * i + 3]);
    img.Tvec() = Eigen::Vector3d(pb.tvecs[3 * i], pb.tvecs[3 * i + 1], pb.tvecs[3 * i + 2]);
    b->rec.AddImage(img);
    b->rec.RegisterImage(i + 1);  // registration order = image index order
  }
  b->point_id.assign(pb.num_points, 0);
  for (int p = 0; p < pb.num_points; ++p) {
    if (pb.track_start[p + 1] == pb.track_start[p]) continue;
    colmap::Track track;
    for (int64_t k = pb.track_start[p]; k < pb.track_start[p + 1]; ++k)
      track.AddElement(pb.obs_image[k] + 1, b->obs_line_idx[k]);
    b->point_id[p] = b->rec.AddPoint3D(
        Eigen::Vector3d(pb.points[3 * p], pb.points[3 * p + 1], pb.points[3 * p + 2]), track);
  }
}

void ReadBack(const FilterProblem& pb, const Built& b, uint8_t* obs_deleted, uint8_t* point_deleted,
              double* point_error) {
  for (int p = 0; p < pb.num_points; ++p) {
    const bool alive = b.point_id[p] != 0 && b.rec.ExistsPoint3D(b.point_id[p]);
    if (point_deleted) point_deleted[p] = (b.point_id[p] != 0 && !alive) ? 1 : 0;
    if (point_error && alive) point_error[p] = b.rec.Point3D(b.point_id[p]).Error();
    for (int64_t k = pb.track_start[p]; k < pb.track_start[p + 1]; ++k)
      obs_deleted[k] = b.rec.Image(pb.obs_image[k] + 1).Line(b.obs_line_idx[k]).HasPoint3D() ? 0 : 1;
  }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int ref_filter_points3d(
    const FilterProblem* pb, double max_reproj_error, double min_tri_angle_deg,
    uint8_t* obs_deleted, uint8_t* point_deleted, double* point_error, uint64_t* num_filtered) {
  Built b;
  Build(*pb, &b);
  std::unordered_set<colmap::point3D_t> ids;
  for (const auto id : b.point_id)
    if (id != 0) ids.insert(id);
  *num_filtered = b.rec.FilterPoints3D(max_reproj_error, min_tri_angle_deg, ids);
  ReadBack(*pb, b, obs_deleted, point_deleted, point_error);
  return 0;
}

extern "C" __attribute__((visibility("default"))) int ref_filter_negative_depth(
    const FilterProblem* pb, uint8_t* obs_deleted, uint8_t* point_deleted, uint64_t* num_filtered) {
  Built b;
  Build(*pb, &b);
  *num_filtered = b.rec.FilterObservationsWithNegativeDepth();
  ReadBack(*pb, b, obs_deleted, point_deleted, nullptr);
  return 0;
}
