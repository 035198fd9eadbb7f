// adaptor_test.cc — exercises the header-only C++ adaptor (privacy_preserving_sfm_b200/cpp/
// ppsfm_adaptor.h) the way IncrementalMapper uses the reference API
// (src/sfm/incremental_mapper.cc:673-735, 907-930).  Inputs / outputs are flat binary files so
// that the pytest driver can compare against the CPU oracle.
//   adaptor_test pose <in.bin> <out.bin>
//   adaptor_test ba   <in.bin> <out.bin>
//   adaptor_test filter | depth | tri <in.bin> <out.bin>   (same scene file as `ba` + aligned flags)
#include <cstdio>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "ppsfm_adaptor.h"

using namespace ppsfm;

static std::vector<double> ReadDoubles(FILE* f, size_t n) {
  std::vector<double> v(n);
  if (n && fread(v.data(), sizeof(double), n, f) != n) std::abort();
  return v;
}
static int64_t ReadI64(FILE* f) {
  int64_t v;
  if (fread(&v, sizeof(v), 1, f) != 1) std::abort();
  return v;
}

// --- minimal stand-ins for colmap::Camera / Image / Point3D / Track / Reconstruction ------------
struct TestCamera {
  int model_id = 1;
  std::vector<double> params;
  size_t width = 1000, height = 1000;
  size_t Width() const { return width; }
  size_t Height() const { return height; }
  int ModelId() const { return model_id; }
  size_t NumParams() const { return params.size(); }
  const double* ParamsData() const { return params.data(); }
  double* ParamsData() { return params.data(); }
};
struct TestImage {
  camera_t camera_id = 1;
  Vector4d qvec{};
  Vector3d tvec{};
  FeatureLines lines;
  camera_t CameraId() const { return camera_id; }
  void NormalizeQvec() {
    double n = 0;
    for (int k = 0; k < 4; ++k) n += qvec[k] * qvec[k];
    n = std::sqrt(n);
    if (n > 0) for (int k = 0; k < 4; ++k) qvec[k] /= n;
  }
  Vector4d& Qvec() { return qvec; }
  Vector3d& Tvec() { return tvec; }
  const FeatureLines& Lines() const { return lines; }
};
struct TrackElement { image_t image_id; uint32_t line_idx; };
struct TestTrack {
  std::vector<TrackElement> elements;
  size_t Length() const { return elements.size(); }
  const std::vector<TrackElement>& Elements() const { return elements; }
};
struct TestPoint3D {
  Vector3d xyz{};
  TestTrack track;
  double error = -1.0;
  Vector3d& XYZ() { return xyz; }
  TestTrack& Track() { return track; }
  void SetError(double e) { error = e; }
};
struct TestReconstruction {
  std::unordered_map<camera_t, TestCamera> cameras;
  std::unordered_map<image_t, TestImage> images;
  std::unordered_map<point3D_t, TestPoint3D> points;
  TestCamera& Camera(camera_t id) { return cameras.at(id); }
  TestImage& Image(image_t id) { return images.at(id); }
  TestPoint3D& Point3D(point3D_t id) { return points.at(id); }
  bool ExistsPoint3D(point3D_t id) const { return points.count(id) != 0; }
  // Reconstruction::DeletePoint3D / DeleteObservation (src/base/reconstruction.cc:243-275)
  void DeletePoint3D(point3D_t id) {
    for (const auto& el : points.at(id).track.elements)
      images.at(el.image_id).lines[el.line_idx].SetPoint3DId(kInvalidPoint3DId);
    points.erase(id);
  }
  void DeleteObservation(image_t image_id, uint32_t line_idx) {
    const point3D_t pid = images.at(image_id).lines[line_idx].Point3DId();
    TestPoint3D& p = points.at(pid);
    if (p.track.Length() <= 3) {
      DeletePoint3D(pid);
      return;
    }
    auto& els = p.track.elements;
    for (size_t i = 0; i < els.size(); ++i)
      if (els[i].image_id == image_id && els[i].line_idx == line_idx) {
        els.erase(els.begin() + i);
        break;
      }
    images.at(image_id).lines[line_idx].SetPoint3DId(kInvalidPoint3DId);
  }
};

// scene file of the `ba` mode (+ one aligned flag per observation) -> TestReconstruction
struct Scene {
  int C = 0, P = 0;
  int64_t O = 0;
  TestReconstruction rec;
  std::vector<std::pair<image_t, uint32_t>> obs_ref;  // observation o -> (image, line)
  std::vector<point3D_t> obs_point;
};
static bool LoadScene(const char* in, bool with_aligned, Scene* sc) {
  FILE* f = fopen(in, "rb");
  if (!f) return false;
  sc->C = (int)ReadI64(f);
  sc->P = (int)ReadI64(f);
  sc->O = ReadI64(f);
  const int C = sc->C, P = sc->P;
  const int64_t O = sc->O;
  const std::vector<double> q = ReadDoubles(f, 4 * C), t = ReadDoubles(f, 3 * C),
                            X = ReadDoubles(f, 3 * P), oc = ReadDoubles(f, O), op = ReadDoubles(f, O),
                            ol = ReadDoubles(f, 3 * O), cam = ReadDoubles(f, 4);
  std::vector<double> al(O, 0.0);
  if (with_aligned) al = ReadDoubles(f, O);
  fclose(f);
  TestReconstruction& rec = sc->rec;
  rec.cameras[1].params = cam;
  for (int i = 0; i < C; ++i) {
    TestImage& im = rec.images[i + 1];
    im.qvec = Vector4d{q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]};
    im.tvec = Vector3d{t[3 * i], t[3 * i + 1], t[3 * i + 2]};
  }
  for (int p = 0; p < P; ++p) rec.points[p + 100].xyz = Vector3d{X[3 * p], X[3 * p + 1], X[3 * p + 2]};
  for (int64_t o = 0; o < O; ++o) {
    TestImage& im = rec.images[(image_t)oc[o] + 1];
    const point3D_t pid = (point3D_t)op[o] + 100;
    im.lines.emplace_back(Vector3d{ol[3 * o], ol[3 * o + 1], ol[3 * o + 2]}, al[o] != 0, pid);
    rec.points[pid].track.elements.push_back({(image_t)oc[o] + 1, (uint32_t)im.lines.size() - 1});
    sc->obs_ref.emplace_back((image_t)oc[o] + 1, (uint32_t)im.lines.size() - 1);
    sc->obs_point.push_back(pid);
  }
  return true;
}

// Reconstruction::FilterPoints3D / FilterObservationsWithNegativeDepth through the adaptor
// (controllers/incremental_mapper.cc:120-128, sfm/incremental_mapper.cc:904)
static int RunFilter(const char* in, const char* out, bool depth) {
  Scene sc;
  if (!LoadScene(in, true, &sc)) return 2;
  std::vector<point3D_t> ids;
  for (int p = 0; p < sc.P; ++p) ids.push_back((point3D_t)p + 100);
  const size_t nf = depth ? FilterObservationsWithNegativeDepth(&sc.rec, ids)
                          : FilterPoints3D(&sc.rec, 4.0, 1.5, ids);
  FILE* g = fopen(out, "wb");
  const double hdr[2] = {(double)nf, ImageToWorldThreshold(sc.rec.cameras[1], 12.0)};
  fwrite(hdr, sizeof(double), 2, g);
  for (int p = 0; p < sc.P; ++p) {  // alive flag, error
    const bool alive = sc.rec.ExistsPoint3D((point3D_t)p + 100);
    const double v[2] = {alive ? 1.0 : 0.0, alive ? sc.rec.points[p + 100].error : -1.0};
    fwrite(v, sizeof(double), 2, g);
  }
  for (int64_t o = 0; o < sc.O; ++o) {  // observation still attached to its point?
    const auto& line = sc.rec.images[sc.obs_ref[o].first].lines[sc.obs_ref[o].second];
    const double v = line.HasPoint3D() ? 1.0 : 0.0;
    fwrite(&v, sizeof(double), 1, g);
  }
  fclose(g);
  return 0;
}

// EstimateTriangulation per track and EstimateTriangulationBatch (triangulation.h:143-147,
// options of sfm/incremental_triangulator.cc:518-533)
static int RunTri(const char* in, const char* out) {
  Scene sc;
  if (!LoadScene(in, false, &sc)) return 2;
  typedef TriangulationEstimator::PoseData<TestCamera> PoseData;
  std::vector<std::vector<TriangulationEstimator::PointData>> point_data(sc.P);
  std::vector<std::vector<PoseData>> pose_data(sc.P);
  for (int p = 0; p < sc.P; ++p)
    for (const auto& el : sc.rec.points[p + 100].track.elements) {
      TestImage& im = sc.rec.images[el.image_id];
      const double w = im.qvec[0], x = im.qvec[1], y = im.qvec[2], z = im.qvec[3];
      Matrix3x4d Pm{};  // column-major [R | t], R = QuaternionToRotationMatrix (base/pose.cc:46-51)
      double* m = Pm.data();
      m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y + w * z);     m[2] = 2 * (x * z - w * y);
      m[3] = 2 * (x * y - w * z);     m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z + w * x);
      m[6] = 2 * (x * z + w * y);     m[7] = 2 * (y * z - w * x);     m[8] = 1 - 2 * (x * x + y * y);
      for (int k = 0; k < 3; ++k) m[9 + k] = im.tvec[k];
      Vector3d c{};
      for (int k = 0; k < 3; ++k) c[k] = -(m[3 * k] * m[9] + m[3 * k + 1] * m[10] + m[3 * k + 2] * m[11]);
      point_data[p].emplace_back(im.lines[el.line_idx].Line());
      pose_data[p].emplace_back(Pm, c, &sc.rec.cameras[1]);
    }
  EstimateTriangulationOptions o;
  o.min_tri_angle = 1.5 * 3.14159265358979323846 / 180.0;
  o.residual_type = TriangulationEstimator::ResidualType::ANGULAR_ERROR;
  o.ransac_options.max_error = 2.0 * 3.14159265358979323846 / 180.0;
  o.ransac_options.confidence = 0.9999;
  o.ransac_options.min_inlier_ratio = 0.02;
  o.ransac_options.max_num_trials = 10000;
  o.exhaustive_threshold = 15;
  std::vector<char> ok;
  std::vector<std::vector<char>> masks;
  std::vector<Vector3d> xyz;
  EstimateTriangulationBatch<TestCamera>(o, point_data, pose_data, &ok, &masks, &xyz);
  // the per-track signature on the first track must agree with the batch
  std::vector<char> m0;
  Vector3d x0{};
  const bool ok0 = EstimateTriangulation<TestCamera>(o, point_data[0], pose_data[0], &m0, &x0);
  FILE* g = fopen(out, "wb");
  const double hdr[2] = {ok0 ? 1.0 : 0.0, (ok0 == (ok[0] != 0) && (!ok0 || (m0 == masks[0] &&
                         x0[0] == xyz[0][0] && x0[1] == xyz[0][1] && x0[2] == xyz[0][2]))) ? 1.0 : 0.0};
  fwrite(hdr, sizeof(double), 2, g);
  for (int p = 0; p < sc.P; ++p) {
    const double v[4] = {ok[p] ? 1.0 : 0.0, xyz[p][0], xyz[p][1], xyz[p][2]};
    fwrite(v, sizeof(double), 4, g);
  }
  for (int p = 0; p < sc.P; ++p)
    for (char c : masks[p]) {
      const double v = c ? 1.0 : 0.0;
      fwrite(&v, sizeof(double), 1, g);
    }
  fclose(g);
  return 0;
}

static int RunPose(const char* in, const char* out) {
  FILE* f = fopen(in, "rb");
  if (!f) return 2;
  const size_t n = (size_t)ReadI64(f);
  const std::vector<double> lines = ReadDoubles(f, 3 * n), pts = ReadDoubles(f, 3 * n),
                            al = ReadDoubles(f, n);
  fclose(f);
  FeatureLines lines2D(n);
  std::vector<Vector3d> points3D(n);
  for (size_t i = 0; i < n; ++i) {
    lines2D[i] = FeatureLine(Vector3d{lines[3 * i], lines[3 * i + 1], lines[3 * i + 2]}, al[i] != 0);
    points3D[i] = Vector3d{pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
  }
  // mapper options, src/sfm/incremental_mapper.cc:673-681
  RANSACOptions options;
  options.max_error = 12.0 / 1000.0;
  options.min_inlier_ratio = 0.25;
  options.min_num_trials = 100;
  options.max_num_trials = 10000;
  options.confidence = 0.99999;
  SetPRNGSeed(0);
  Vector4d qvec;
  Vector3d tvec;
  size_t num_inliers = 0;
  std::vector<char> mask;
  const bool ok = EstimateAbsolutePoseFromLines(options, lines2D, points3D, &qvec, &tvec,
                                                &num_inliers, &mask);
  Vector4d q_ref = qvec;
  Vector3d t_ref = tvec;
  bool ok_ref = false;
  if (ok) {
    std::vector<Vector3d> l3(n);
    for (size_t i = 0; i < n; ++i) l3[i] = lines2D[i].Line();
    TestCamera cam;
    cam.params = {1000.0, 1000.0, 500.0, 500.0};
    AbsolutePoseRefinementOptions ro;
    ro.print_summary = false;
    ok_ref = RefineAbsolutePoseFromLines(ro, mask, l3, points3D, &q_ref, &t_ref, &cam);
  }
  // also the estimator concept on the first 6 correspondences
  std::vector<FeatureLine> x6(lines2D.begin(), lines2D.begin() + 6);
  std::vector<Vector3d> y6(points3D.begin(), points3D.begin() + 6);
  const std::vector<Matrix3x4d> models = P6LEstimator::Estimate(x6, y6);
  FILE* g = fopen(out, "wb");
  const double hdr[4] = {ok ? 1.0 : 0.0, (double)num_inliers, ok_ref ? 1.0 : 0.0, (double)models.size()};
  fwrite(hdr, sizeof(double), 4, g);
  fwrite(qvec.data(), sizeof(double), 4, g);
  fwrite(tvec.data(), sizeof(double), 3, g);
  fwrite(q_ref.data(), sizeof(double), 4, g);
  fwrite(t_ref.data(), sizeof(double), 3, g);
  std::vector<double> m(n, 0.0);
  for (size_t i = 0; i < mask.size(); ++i) m[i] = mask[i] ? 1.0 : 0.0;
  fwrite(m.data(), sizeof(double), n, g);
  fclose(g);
  return 0;
}

// refine: ParameterizeCameras with refine_extra_params on a SIMPLE_RADIAL camera (id 1) next to a
// second, constant camera (id 2, config.SetConstantCamera) that the odd images use
static int RunBA(const char* in, const char* out, bool refine = false) {
  FILE* f = fopen(in, "rb");
  if (!f) return 2;
  const int C = (int)ReadI64(f), P = (int)ReadI64(f);
  const int64_t O = ReadI64(f);
  const std::vector<double> q = ReadDoubles(f, 4 * C), t = ReadDoubles(f, 3 * C),
                            X = ReadDoubles(f, 3 * P), oc = ReadDoubles(f, O), op = ReadDoubles(f, O),
                            ol = ReadDoubles(f, 3 * O), cam = ReadDoubles(f, 4);
  fclose(f);
  TestReconstruction rec;
  rec.cameras[1].params = cam;
  if (refine) {
    rec.cameras[1].model_id = 2;
    rec.cameras[1].params = {cam[0], cam[2], cam[3], 0.08};
    rec.cameras[2] = rec.cameras[1];
  }
  for (int i = 0; i < C; ++i) {
    TestImage& im = rec.images[i + 1];
    if (refine) im.camera_id = 1 + i % 2;
    im.qvec = Vector4d{q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]};
    im.tvec = Vector3d{t[3 * i], t[3 * i + 1], t[3 * i + 2]};
  }
  for (int p = 0; p < P; ++p) rec.points[p + 100].xyz = Vector3d{X[3 * p], X[3 * p + 1], X[3 * p + 2]};
  for (int64_t o = 0; o < O; ++o) {
    TestImage& im = rec.images[(image_t)oc[o] + 1];
    const point3D_t pid = (point3D_t)op[o] + 100;
    im.lines.emplace_back(Vector3d{ol[3 * o], ol[3 * o + 1], ol[3 * o + 2]}, false, pid);
    rec.points[pid].track.elements.push_back({(image_t)oc[o] + 1, (uint32_t)im.lines.size() - 1});
  }
  // IncrementalMapper::AdjustGlobalBundle (src/sfm/incremental_mapper.cc:893-939)
  BundleAdjustmentOptions options;
  options.solver_options.max_num_iterations = refine ? 6 : 20;
  options.solver_options.gradient_tolerance = 1e-4;
  options.print_summary = false;
  options.refine_extra_params = refine;
  BundleAdjustmentConfig config;
  if (refine) config.SetConstantCamera(2);
  for (int i = 0; i < C; ++i) config.AddImage(i + 1);
  config.SetConstantPose(1);
  config.SetConstantTvec(2, {0});
  BundleAdjuster<TestReconstruction> adjuster(options, config);
  const bool ok = adjuster.Solve(&rec);
  FILE* g = fopen(out, "wb");
  const double hdr[4] = {ok ? 1.0 : 0.0, adjuster.Summary().initial_cost,
                         adjuster.Summary().final_cost,
                         (double)(adjuster.Summary().num_successful_steps +
                                  adjuster.Summary().num_unsuccessful_steps)};
  fwrite(hdr, sizeof(double), 4, g);
  for (int i = 0; i < C; ++i) fwrite(rec.images[i + 1].qvec.data(), sizeof(double), 4, g);
  for (int i = 0; i < C; ++i) fwrite(rec.images[i + 1].tvec.data(), sizeof(double), 3, g);
  for (int p = 0; p < P; ++p) fwrite(rec.points[p + 100].xyz.data(), sizeof(double), 3, g);
  if (refine) {
    fwrite(rec.cameras[1].params.data(), sizeof(double), 4, g);
    fwrite(rec.cameras[2].params.data(), sizeof(double), 4, g);
  }
  fclose(g);
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 4) return 1;
  if (std::string(argv[1]) == "pose") return RunPose(argv[2], argv[3]);
  if (std::string(argv[1]) == "ba") return RunBA(argv[2], argv[3]);
  if (std::string(argv[1]) == "ba_refine") return RunBA(argv[2], argv[3], true);
  if (std::string(argv[1]) == "filter") return RunFilter(argv[2], argv[3], false);
  if (std::string(argv[1]) == "depth") return RunFilter(argv[2], argv[3], true);
  if (std::string(argv[1]) == "tri") return RunTri(argv[2], argv[3]);
  return 1;
}
