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

// ---------------------------------------------------------------------------------------------
// Text model format: the reference's own Reconstruction::WriteText / ReadText
// (src/base/reconstruction.cc:543-553, 721-1095) — cameras.txt, images.txt with
// LINES2D[] as (A, B, C, is_aligned, POINT3D_ID), points3D.txt with TRACK[] as (IMAGE_ID, line_idx).
// ref_model_write_text builds the reconstruction of a FilterProblem through the reference's own
// members (optionally runs its FilterPoints3D first, so that lines without a point and deleted
// points appear) and lets the reference write it; ref_model_read_text lets the reference read a
// directory and hands the result back as flat arrays sorted by id.
// ---------------------------------------------------------------------------------------------
#include <algorithm>
#include <cstdio>
#include <cstring>

namespace {

struct ReadModel {
  colmap::Reconstruction rec;
  std::vector<colmap::camera_t> cams;
  std::vector<colmap::image_t> imgs;
  std::vector<colmap::point3D_t> pts;
  int64_t num_lines = 0, num_track = 0;
};

void Index(ReadModel* m) {
  for (const auto& c : m->rec.Cameras()) m->cams.push_back(c.first);
  for (const auto& i : m->rec.Images()) {
    m->imgs.push_back(i.first);
    m->num_lines += static_cast<int64_t>(i.second.Lines().size());
  }
  for (const auto& p : m->rec.Points3D()) {
    m->pts.push_back(p.first);
    m->num_track += static_cast<int64_t>(p.second.Track().Length());
  }
  std::sort(m->cams.begin(), m->cams.end());
  std::sort(m->imgs.begin(), m->imgs.end());
  std::sort(m->pts.begin(), m->pts.end());
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int ref_model_write_text(
    const FilterProblem* pb, int apply_filter, double max_reproj_error, double min_tri_angle_deg,
    const char* path) {
  Built b;
  Build(*pb, &b);
  for (int i = 0; i < pb->num_images; ++i) {
    char name[32];
    std::snprintf(name, sizeof(name), "image%06d.jpg", i);
    b.rec.Image(i + 1).SetName(name);
  }
  if (apply_filter) {
    std::unordered_set<colmap::point3D_t> ids;
    for (const auto id : b.point_id)
      if (id != 0) ids.insert(id);
    b.rec.FilterPoints3D(max_reproj_error, min_tri_angle_deg, ids);
  }
  b.rec.WriteText(path);
  return 0;
}

extern "C" __attribute__((visibility("default"))) void* ref_model_read_text(const char* path) {
  auto* m = new ReadModel;
  m->rec.ReadText(path);
  Index(m);
  return m;
}

// the reference writes what it has read: text -> Reconstruction -> text
extern "C" __attribute__((visibility("default"))) void ref_model_rewrite_text(void* h, const char* path) {
  static_cast<ReadModel*>(h)->rec.WriteText(path);
}

extern "C" __attribute__((visibility("default"))) void ref_model_free(void* h) {
  delete static_cast<ReadModel*>(h);
}

// sizes: cameras, images, lines (all images), points, track elements (all points), registered images
extern "C" __attribute__((visibility("default"))) void ref_model_sizes(void* h, int64_t* sizes) {
  const auto* m = static_cast<ReadModel*>(h);
  sizes[0] = static_cast<int64_t>(m->cams.size());
  sizes[1] = static_cast<int64_t>(m->imgs.size());
  sizes[2] = m->num_lines;
  sizes[3] = static_cast<int64_t>(m->pts.size());
  sizes[4] = m->num_track;
  sizes[5] = static_cast<int64_t>(m->rec.NumRegImages());
}

extern "C" __attribute__((visibility("default"))) void ref_model_dump(
    void* h, int64_t* cam_id, int32_t* cam_model, int64_t* cam_size, int32_t* cam_num_params,
    double* cam_params /* 12 per camera */, int64_t* img_id, double* img_qvec, double* img_tvec,
    int64_t* img_camera, char* img_name /* 64 per image */, int64_t* line_start /* images + 1 */,
    double* lines, uint8_t* aligned, int64_t* line_point /* -1: none */, int64_t* pt_id,
    double* pt_xyz, uint8_t* pt_color, double* pt_error, int64_t* track_start /* points + 1 */,
    int64_t* track_image, int64_t* track_line) {
  const auto* m = static_cast<ReadModel*>(h);
  for (size_t c = 0; c < m->cams.size(); ++c) {
    const auto& cam = m->rec.Camera(m->cams[c]);
    cam_id[c] = cam.CameraId();
    cam_model[c] = cam.ModelId();
    cam_size[2 * c] = static_cast<int64_t>(cam.Width());
    cam_size[2 * c + 1] = static_cast<int64_t>(cam.Height());
    cam_num_params[c] = static_cast<int32_t>(cam.NumParams());
    for (size_t k = 0; k < cam.NumParams() && k < 12; ++k) cam_params[12 * c + k] = cam.Params()[k];
  }
  int64_t l = 0;
  for (size_t i = 0; i < m->imgs.size(); ++i) {
    const auto& img = m->rec.Image(m->imgs[i]);
    img_id[i] = img.ImageId();
    for (int k = 0; k < 4; ++k) img_qvec[4 * i + k] = img.Qvec(k);
    for (int k = 0; k < 3; ++k) img_tvec[3 * i + k] = img.Tvec(k);
    img_camera[i] = img.CameraId();
    std::strncpy(img_name + 64 * i, img.Name().c_str(), 63);
    line_start[i] = l;
    for (const auto& fl : img.Lines()) {
      for (int k = 0; k < 3; ++k) lines[3 * l + k] = fl.Line()(k);
      aligned[l] = fl.IsAligned() ? 1 : 0;
      line_point[l] = fl.HasPoint3D() ? static_cast<int64_t>(fl.Point3DId()) : -1;
      ++l;
    }
  }
  line_start[m->imgs.size()] = l;
  int64_t e = 0;
  for (size_t p = 0; p < m->pts.size(); ++p) {
    const auto& pt = m->rec.Point3D(m->pts[p]);
    pt_id[p] = static_cast<int64_t>(m->pts[p]);
    for (int k = 0; k < 3; ++k) {
      pt_xyz[3 * p + k] = pt.XYZ()(k);
      pt_color[3 * p + k] = pt.Color(k);
    }
    pt_error[p] = pt.Error();
    track_start[p] = e;
    for (const auto& el : pt.Track().Elements()) {
      track_image[e] = el.image_id;
      track_line[e] = el.line_idx;
      ++e;
    }
  }
  track_start[m->pts.size()] = e;
}

// ---------------------------------------------------------------------------------------------
// Reconstruction::Normalize (src/base/reconstruction.cc:302-398), called by the mapper after
// every global bundle adjustment (src/sfm/incremental_mapper.cc:934-936): the reference's own
// member on the reconstruction of a FilterProblem; translations and points are handed back.
// ---------------------------------------------------------------------------------------------
extern "C" __attribute__((visibility("default"))) int ref_normalize(
    const FilterProblem* pb, double extent, double p0, double p1, int use_images, double* tvecs,
    double* points) {
  Built b;
  Build(*pb, &b);
  b.rec.Normalize(extent, p0, p1, use_images != 0);
  for (int i = 0; i < pb->num_images; ++i)
    for (int k = 0; k < 3; ++k) tvecs[3 * i + k] = b.rec.Image(i + 1).Tvec(k);
  for (int p = 0; p < pb->num_points; ++p)
    for (int k = 0; k < 3; ++k)
      points[3 * p + k] = b.point_id[p] != 0 ? b.rec.Point3D(b.point_id[p]).XYZ()(k) : pb->points[3 * p + k];
  return 0;
}
