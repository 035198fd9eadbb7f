"""ctypes binding of oracle/_ref/libref_p6l.so: the REFERENCE's own sources of the absolute-pose
RANSAC path (P6LEstimator, re3q3, ComputeSquaredLineReprojectionError, RANSAC<>::Estimate,
RandomSampler, support measurers) compiled from /root/reference by oracle/build_ref.sh against the
Eigen / glog stand-ins of oracle/ref/shim/.

TEST INFRASTRUCTURE ONLY (same rule as the oracle: tests/, smoke() and bench.py's CPU arms).  The
library is built in the container that has /root/reference and travels to the GPU box as a
git-ignored file; nothing here reads /root/reference at run time.  Same call surface as the
matching functions of ``oracle/__init__.py`` so that a test can run both side by side.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import RansacOptions, RansacReport, make_options  # noqa: F401  (same C structs)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_p6l.so")
COST_LIB_PATH = os.path.join(_HERE, "_ref", "libref_cost.so")
TRI_LIB_PATH = os.path.join(_HERE, "_ref", "libref_tri.so")
BA_SETUP_LIB_PATH = os.path.join(_HERE, "_ref", "libref_ba_setup.so")
FILTER_LIB_PATH = os.path.join(_HERE, "_ref", "libref_filter.so")
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_lib = None


def build(reference_root="/root/reference"):
    """Compile oracle/_ref/ where the reference tree exists; returns True if the library is there."""
    srcs = [os.path.join(_HERE, "ref", "ref_p6l.cc"), os.path.join(_HERE, "ref", "ref_cost.cc"),
            os.path.join(_HERE, "ref", "ref_triangulation.cc"),
            os.path.join(_HERE, "ref", "ref_ba_setup.cc"),
            os.path.join(_HERE, "ref", "ref_filter.cc"),
            os.path.join(_HERE, "ref", "ref_corr_graph.cc"),
            os.path.join(_HERE, "..", "privacy_preserving_sfm_b200", "cpp", "ppsfm_adaptor.h"),
            os.path.join(_HERE, "triangulation_oracle.cc"),
            os.path.join(_HERE, "ref", "shim", "minieigen.h"),
            os.path.join(_HERE, "ref", "shim", "ceres", "ceres.h"),
            os.path.join(_HERE, "ref", "shim", "glog", "logging.h"),
            os.path.join(_HERE, "eigen_restated.h"), os.path.join(_HERE, "build_ref.sh")]
    fresh = all(os.path.exists(lp) and all(os.path.getmtime(lp) >= os.path.getmtime(s)
                                           for s in srcs)
                for lp in (LIB_PATH, COST_LIB_PATH, TRI_LIB_PATH, BA_SETUP_LIB_PATH,
                           FILTER_LIB_PATH))
    if not fresh and os.path.isdir(os.path.join(reference_root, "src", "estimators")):
        subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")],
                              stdout=subprocess.DEVNULL)
    return all(os.path.exists(lp)
               for lp in (LIB_PATH, COST_LIB_PATH, TRI_LIB_PATH, BA_SETUP_LIB_PATH,
                          FILTER_LIB_PATH))


def available():
    return build()


def lib():
    global _lib
    if _lib is None:
        if not build():
            raise RuntimeError("oracle/_ref/libref_p6l.so is not built (needs /root/reference)")
        L = C.CDLL(LIB_PATH)
        L.ref_set_prng_seed.argtypes = [C.c_uint32]
        L.ref_prng_peek.restype = C.c_uint32
        L.ref_line_residuals.argtypes = [_dp, _dp, C.c_size_t, _dp, _dp]
        L.ref_inlier_support.argtypes = [_dp, C.c_size_t, C.c_double, C.POINTER(C.c_uint64), _dp]
        L.ref_inlier_support_compare.argtypes = [C.c_uint64, C.c_double, C.c_uint64, C.c_double]
        L.ref_inlier_support_compare.restype = C.c_int
        L.ref_mestimator_support.argtypes = [_dp, C.c_size_t, C.c_double,
                                             C.POINTER(C.c_uint64), _dp]
        L.ref_compute_num_trials.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double]
        L.ref_compute_num_trials.restype = C.c_uint64
        L.ref_sample_table.argtypes = [C.c_size_t, C.c_size_t, _u32p]
        L.ref_re3q3.argtypes = [_dp, _dp]
        L.ref_re3q3.restype = C.c_int
        L.ref_p6l_estimate.argtypes = [_dp, _u8p, _dp, _dp]
        L.ref_p6l_estimate.restype = C.c_int
        L.ref_ransac_p6l.argtypes = [_dp, _u8p, _dp, C.c_size_t, C.POINTER(RansacOptions),
                                     C.POINTER(RansacReport), _u8p]
        L.ref_estimate_absolute_pose_from_lines.argtypes = [
            _dp, _u8p, _dp, C.c_size_t, C.POINTER(RansacOptions), _dp, _dp,
            C.POINTER(C.c_uint64), _u8p]
        L.ref_estimate_absolute_pose_from_lines.restype = C.c_int
        L.ref_refine_absolute_pose_setup.argtypes = [
            _dp, _dp, _u8p, C.c_size_t, C.c_int, _dp, C.c_int, C.c_int, C.c_double, C.c_int,
            C.c_double, _dp, _dp, _dp]
        L.ref_refine_absolute_pose_setup.restype = C.c_int
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_u8p)


def set_prng_seed(seed):
    lib().ref_set_prng_seed(seed)


def prng_peek():
    return int(lib().ref_prng_peek())


def line_residuals(lines, points, model):
    lines, lp = _d(lines)
    points, pp = _d(points)
    model, mp = _d(model)
    out = np.empty(lines.shape[0], dtype=np.float64)
    lib().ref_line_residuals(lp, pp, lines.shape[0], mp, out.ctypes.data_as(_dp))
    return out


def inlier_support(residuals, max_residual):
    residuals, rp = _d(residuals)
    cnt, s = C.c_uint64(), C.c_double()
    lib().ref_inlier_support(rp, residuals.shape[0], max_residual, C.byref(cnt), C.byref(s))
    return int(cnt.value), float(s.value)


def inlier_support_compare(n1, s1, n2, s2):
    return bool(lib().ref_inlier_support_compare(n1, s1, n2, s2))


def mestimator_support(residuals, max_residual):
    residuals, rp = _d(residuals)
    cnt, s = C.c_uint64(), C.c_double()
    lib().ref_mestimator_support(rp, residuals.shape[0], max_residual, C.byref(cnt), C.byref(s))
    return int(cnt.value), float(s.value)


def compute_num_trials(num_inliers, num_samples, confidence, multiplier):
    return int(lib().ref_compute_num_trials(num_inliers, num_samples, confidence, multiplier))


def sample_table(n, num_trials):
    out = np.empty((num_trials, 6), dtype=np.uint32)
    lib().ref_sample_table(n, num_trials, out.ctypes.data_as(_u32p))
    return out


def re3q3(coeffs):
    coeffs, cp = _d(coeffs)
    assert coeffs.shape == (3, 10)
    sol = np.zeros(24, dtype=np.float64)
    n = lib().ref_re3q3(cp, sol.ctypes.data_as(_dp))
    return sol.reshape(8, 3)[:n].copy()


def p6l_estimate(lines6, aligned6, points6):
    lines6, lp = _d(lines6)
    points6, pp = _d(points6)
    aligned6, ap = _u8(aligned6)
    out = np.zeros((8, 12), dtype=np.float64)
    n = lib().ref_p6l_estimate(lp, ap, pp, out.ctypes.data_as(_dp))
    return out[:n].copy()


def ransac_p6l(lines, aligned, points, options):
    """colmap::RANSAC<P6LEstimator>(options).Estimate.  best_trial / best_model_idx /
    num_models_scored of the report are not observable from outside the loop (-1 / 0)."""
    lines, lp = _d(lines)
    points, pp = _d(points)
    aligned, ap = _u8(aligned)
    n = lines.shape[0]
    rep = RansacReport()
    mask = np.zeros(n, dtype=np.uint8)
    lib().ref_ransac_p6l(lp, ap, pp, n, C.byref(options), C.byref(rep), mask.ctypes.data_as(_u8p))
    return rep, mask


def estimate_absolute_pose_from_lines(lines, aligned, points, options):
    """colmap::EstimateAbsolutePoseFromLines (src/estimators/pose.cc:52-94):
    (ok, qvec, tvec, num_inliers, inlier_mask)."""
    lines, lp = _d(lines)
    points, pp = _d(points)
    aligned, ap = _u8(aligned)
    n = lines.shape[0]
    q, t = np.zeros(4), np.zeros(3)
    ninl = C.c_uint64()
    mask = np.zeros(n, dtype=np.uint8)
    ok = lib().ref_estimate_absolute_pose_from_lines(
        lp, ap, pp, n, C.byref(options), q.ctypes.data_as(_dp), t.ctypes.data_as(_dp),
        C.byref(ninl), mask.ctypes.data_as(_u8p))
    return bool(ok), q, t, int(ninl.value), mask


def refine_absolute_pose_setup(lines, points, mask, model, cam_params, qvec, tvec,
                               refine_focal_length=False, refine_extra_params=False,
                               gradient_tolerance=1.0, max_num_iterations=100, loss_scale=1.0):
    """What colmap::RefineAbsolutePoseFromLines (src/estimators/pose.cc:96-213) hands to Ceres,
    as recorded by the stand-in ceres::Problem / ceres::Solve.  Returns a dict."""
    lines, lp = _d(lines)
    points, pp = _d(points)
    mask, mp = _u8(mask)
    cam = np.zeros(12)
    cam[:len(cam_params)] = cam_params
    q, qp = _d(qvec)
    t, tp = _d(tvec)
    out = np.zeros(17)
    lib().ref_refine_absolute_pose_setup(lp, pp, mp, lines.shape[0], model,
                                         cam.ctypes.data_as(_dp), int(refine_focal_length),
                                         int(refine_extra_params), gradient_tolerance,
                                         max_num_iterations, loss_scale, qp, tp,
                                         out.ctypes.data_as(_dp))
    return dict(residual_blocks=int(out[0]), uniform_blocks=bool(out[1]), loss_kind=int(out[2]),
                loss_scale=float(out[3]), constant_points=int(out[4]),
                points_in_inlier_order=bool(out[5]), quaternion_parameterization=bool(out[6]),
                tvec_free=bool(out[7]), camera_variable_mask=int(out[8]),
                linear_solver_type=int(out[9]), gradient_tolerance=float(out[10]),
                max_num_iterations=int(out[11]), num_threads=int(out[12]), qvec=out[13:17].copy())


# ---------------------------------------------------------------------------------------------
# oracle/_ref/libref_cost.so: the reference's line cost functors and camera models
# ---------------------------------------------------------------------------------------------
_cost = None
_ip = C.POINTER(C.c_int)


def cost_lib():
    global _cost
    if _cost is None:
        if not build():
            raise RuntimeError("oracle/_ref/libref_cost.so is not built (needs /root/reference)")
        L = C.CDLL(COST_LIB_PATH)
        L.ref_camera_num_params.argtypes = [C.c_int]
        L.ref_camera_param_idxs.argtypes = [C.c_int, C.c_int, _ip]
        L.ref_world_to_image.argtypes = [C.c_int, _dp, C.c_double, C.c_double, _dp]
        L.ref_image_to_world_threshold.argtypes = [C.c_int, _dp, C.c_double]
        L.ref_image_to_world.argtypes = [C.c_int, _dp, C.c_int, _dp, _dp]
        L.ref_has_bogus_params.argtypes = [C.c_int, _dp, C.c_uint64, C.c_uint64, C.c_double,
                                           C.c_double, C.c_double]
        L.ref_image_to_world_threshold.restype = C.c_double
        L.ref_line_cost.argtypes = [C.c_int] + [_dp] * 10
        L.ref_constant_pose_line_cost.argtypes = [C.c_int] + [_dp] * 8
        _cost = L
    return _cost


def camera_num_params(model):
    return int(cost_lib().ref_camera_num_params(model))


def camera_param_idxs(model, group):
    """group 0: focal length, 1: principal point, 2: extra parameters."""
    out = (C.c_int * 12)()
    n = cost_lib().ref_camera_param_idxs(model, group, out)
    return [int(out[i]) for i in range(n)]


def world_to_image(model, params, u, v):
    params, pp = _d(params)
    xy = np.zeros(2)
    cost_lib().ref_world_to_image(model, pp, u, v, xy.ctypes.data_as(_dp))
    return xy


def image_to_world(model, params, xy):
    """The reference's CameraModelImageToWorld on pixels [n, 2]."""
    params, pp = _d(params)
    xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
    uv = np.zeros_like(xy)
    cost_lib().ref_image_to_world(model, pp, len(xy), xy.ctypes.data_as(_dp), uv.ctypes.data_as(_dp))
    return uv


def has_bogus_params(model, params, width, height, min_focal_length_ratio, max_focal_length_ratio,
                     max_extra_param):
    params, pp = _d(params)
    return bool(cost_lib().ref_has_bogus_params(model, pp, width, height, min_focal_length_ratio,
                                                max_focal_length_ratio, max_extra_param))


def image_to_world_threshold(model, params, threshold):
    params, pp = _d(params)
    return float(cost_lib().ref_image_to_world_threshold(model, pp, threshold))


def line_cost_intr(model, cam_params, line, q, t, X):
    """BundleAdjustmentLineCostFunction<Model>: r, Jq[2x4], Jt[2x3], JX[2x3], Jcamera[2x12]."""
    cam = np.zeros(12)
    cam[:len(cam_params)] = cam_params
    line, lp = _d(line)
    q, qp = _d(q)
    t, tp = _d(t)
    X, xp = _d(X)
    r, jq, jt, jx = np.zeros(2), np.zeros((2, 4)), np.zeros((2, 3)), np.zeros((2, 3))
    jc = np.zeros((2, 12))
    ok = cost_lib().ref_line_cost(model, cam.ctypes.data_as(_dp), lp, qp, tp, xp,
                                  r.ctypes.data_as(_dp), jq.ctypes.data_as(_dp),
                                  jt.ctypes.data_as(_dp), jx.ctypes.data_as(_dp),
                                  jc.ctypes.data_as(_dp))
    assert ok == 1
    return r, jq, jt, jx, jc


def line_residual(model, cam_params, line, q, t, X):
    """The functor on plain doubles (no Jacobians): r[2]."""
    cam = np.zeros(12)
    cam[:len(cam_params)] = cam_params
    line, lp = _d(line)
    q, qp = _d(q)
    t, tp = _d(t)
    X, xp = _d(X)
    r = np.zeros(2)
    ok = cost_lib().ref_line_cost(model, cam.ctypes.data_as(_dp), lp, qp, tp, xp,
                                  r.ctypes.data_as(_dp), None, None, None, None)
    assert ok == 1
    return r


def constant_pose_line_cost(model, cam_params, line, q, t, X):
    """BundleAdjustmentConstantPoseLineCostFunction<Model>: r, JX[2x3], Jcamera[2x12]."""
    cam = np.zeros(12)
    cam[:len(cam_params)] = cam_params
    line, lp = _d(line)
    q, qp = _d(q)
    t, tp = _d(t)
    X, xp = _d(X)
    r, jx, jc = np.zeros(2), np.zeros((2, 3)), np.zeros((2, 12))
    ok = cost_lib().ref_constant_pose_line_cost(model, cam.ctypes.data_as(_dp), lp, qp, tp, xp,
                                                r.ctypes.data_as(_dp), jx.ctypes.data_as(_dp),
                                                jc.ctypes.data_as(_dp))
    assert ok == 1
    return r, jx, jc


# ---------------------------------------------------------------------------------------------
# oracle/_ref/libref_tri.so: the reference's LORANSAC / CombinationSampler around the oracle's
# per-track triangulation estimator
# ---------------------------------------------------------------------------------------------
_tri = None


def estimate_triangulation_batch(tracks, options):
    """Same call as oracle.estimate_triangulation_batch (tracks: filters.FilterProblem, options:
    triangulation.EstimateTriangulationOptions)."""
    global _tri
    if _tri is None:
        if not build():
            raise RuntimeError("oracle/_ref/libref_tri.so is not built (needs /root/reference)")
        _tri = C.CDLL(TRI_LIB_PATH)
        _tri.ref_estimate_triangulation_batch.argtypes = [C.c_void_p, C.c_void_p, _dp, _u8p, _u8p,
                                                          _u32p]
    T, O = len(tracks.points), len(tracks.obs_image)
    xyz = np.zeros((max(T, 1), 3))
    ok, mask = np.zeros(max(T, 1), np.uint8), np.zeros(max(O, 1), np.uint8)
    nt = np.zeros(max(T, 1), np.uint32)
    _tri.ref_estimate_triangulation_batch(C.byref(tracks.struct), C.byref(options),
                                          xyz.ctypes.data_as(_dp), ok.ctypes.data_as(_u8p),
                                          mask.ctypes.data_as(_u8p), nt.ctypes.data_as(_u32p))
    return ok[:T].astype(bool), xyz[:T], mask[:O].astype(bool), nt[:T]


# ---------------------------------------------------------------------------------------------
# oracle/_ref/libref_ba_setup.so: the reference's BundleAdjuster::SetUp against a recording
# ceres::Problem, and the product's adaptor on the same colmap::Reconstruction
# ---------------------------------------------------------------------------------------------
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class _BaScene(C.Structure):
    _fields_ = [("num_cameras", C.c_int32), ("camera_model", _i32p), ("camera_params", _dp),
                ("num_images", C.c_int32), ("image_camera", _i32p), ("qvecs", _dp), ("tvecs", _dp),
                ("num_points", C.c_int32), ("points", _dp), ("num_lines", C.c_int64),
                ("image_line_start", _i64p), ("lines", _dp), ("line_point", _i32p)]


class _BaConfig(C.Structure):
    _fields_ = [("num_config_images", C.c_int32), ("config_images", _i32p),
                ("num_constant_poses", C.c_int32), ("constant_poses", _i32p),
                ("num_constant_tvecs", C.c_int32), ("constant_tvec_image", _i32p),
                ("constant_tvec_mask", _i32p),
                ("num_variable_points", C.c_int32), ("variable_points", _i32p),
                ("num_constant_points", C.c_int32), ("constant_points", _i32p),
                ("num_constant_cameras", C.c_int32), ("constant_cameras", _i32p),
                ("loss_type", C.c_int32), ("loss_scale", C.c_double),
                ("refine_focal_length", C.c_int32), ("refine_principal_point", C.c_int32),
                ("refine_extra_params", C.c_int32), ("refine_extrinsics", C.c_int32)]


_ba_setup = None


def ba_setup_compare(scene, config):
    """scene: dict(camera_model [C], camera_params [C, 12], image_camera [I], qvecs [I, 4],
    tvecs [I, 3], points [P, 3], image_line_start [I + 1], lines [L, 3], line_point [L] (-1: no
    3-D point)); config: dict(images, constant_poses, constant_tvecs {image: [idx]},
    variable_points, constant_points, constant_cameras, loss_type, loss_scale, refine_focal_length,
    refine_principal_point, refine_extra_params, refine_extrinsics), all indices 0-based.
    Returns (rc, reference_text, product_text): the canonical descriptions of what the reference's
    BundleAdjuster::SetUp hands to Ceres and of what the product's adaptor assembles."""
    global _ba_setup
    if _ba_setup is None:
        if not build():
            raise RuntimeError("oracle/_ref/libref_ba_setup.so is not built (needs /root/reference)")
        _ba_setup = C.CDLL(BA_SETUP_LIB_PATH)
        _ba_setup.ref_ba_setup_compare.argtypes = [C.POINTER(_BaScene), C.POINTER(_BaConfig),
                                                   C.c_char_p, C.c_char_p, C.c_size_t]
    keep = []

    def i32(a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        keep.append(a)
        return len(a), a.ctypes.data_as(_i32p)

    def f64(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        keep.append(a)
        return a.ctypes.data_as(_dp)

    s = _BaScene()
    s.num_cameras, s.camera_model = i32(scene["camera_model"])
    s.camera_params = f64(scene["camera_params"])
    s.num_images, s.image_camera = i32(scene["image_camera"])
    s.qvecs, s.tvecs, s.points = f64(scene["qvecs"]), f64(scene["tvecs"]), f64(scene["points"])
    s.num_points = len(scene["points"])
    ils = np.ascontiguousarray(scene["image_line_start"], dtype=np.int64)
    keep.append(ils)
    s.image_line_start = ils.ctypes.data_as(_i64p)
    s.lines = f64(scene["lines"])
    s.num_lines, s.line_point = i32(scene["line_point"])
    c = _BaConfig()
    c.num_config_images, c.config_images = i32(config.get("images", []))
    c.num_constant_poses, c.constant_poses = i32(config.get("constant_poses", []))
    tv = config.get("constant_tvecs", {})
    c.num_constant_tvecs, c.constant_tvec_image = i32(list(tv.keys()))
    _, c.constant_tvec_mask = i32([sum(1 << k for k in idxs) for idxs in tv.values()])
    c.num_variable_points, c.variable_points = i32(config.get("variable_points", []))
    c.num_constant_points, c.constant_points = i32(config.get("constant_points", []))
    c.num_constant_cameras, c.constant_cameras = i32(config.get("constant_cameras", []))
    c.loss_type, c.loss_scale = config.get("loss_type", 0), config.get("loss_scale", 1.0)
    c.refine_focal_length = int(config.get("refine_focal_length", False))
    c.refine_principal_point = int(config.get("refine_principal_point", False))
    c.refine_extra_params = int(config.get("refine_extra_params", False))
    c.refine_extrinsics = int(config.get("refine_extrinsics", True))
    cap = 1 << 22
    a, b = C.create_string_buffer(cap), C.create_string_buffer(cap)
    rc = _ba_setup.ref_ba_setup_compare(C.byref(s), C.byref(c), a, b, cap)
    return rc, a.value.decode(), b.value.decode()


# ---------------------------------------------------------------------------------------------
# oracle/_ref/libref_filter.so: the reference's Reconstruction::FilterPoints3D /
# FilterObservationsWithNegativeDepth on a Reconstruction built through its own Add* members
# ---------------------------------------------------------------------------------------------
_filter = None


def _filter_lib():
    global _filter
    if _filter is None:
        if not build():
            raise RuntimeError("oracle/_ref/libref_filter.so is not built (needs /root/reference)")
        _filter = C.CDLL(FILTER_LIB_PATH)
        _filter.ref_filter_points3d.argtypes = [C.c_void_p, C.c_double, C.c_double, _u8p, _u8p, _dp,
                                                C.POINTER(C.c_uint64)]
        _filter.ref_filter_negative_depth.argtypes = [C.c_void_p, _u8p, _u8p, C.POINTER(C.c_uint64)]
        _filter.ref_model_write_text.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_char_p]
        _filter.ref_model_read_text.argtypes = [C.c_char_p]
        _filter.ref_model_read_text.restype = C.c_void_p
        _filter.ref_model_rewrite_text.argtypes = [C.c_void_p, C.c_char_p]
        _filter.ref_model_free.argtypes = [C.c_void_p]
        _filter.ref_model_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        _filter.ref_model_dump.argtypes = [C.c_void_p] * 22
        _filter.ref_normalize.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp]
    return _filter


def filter_points3d(problem, max_reproj_error, min_tri_angle):
    """Same call as oracle.filter_points3d: (num_filtered, obs_deleted, point_deleted, point_error)."""
    O, P = len(problem.obs_image), len(problem.points)
    od, pd = np.zeros(max(O, 1), np.uint8), np.zeros(max(P, 1), np.uint8)
    pe = np.full(max(P, 1), -1.0)
    nf = C.c_uint64(0)
    _filter_lib().ref_filter_points3d(C.byref(problem.struct), max_reproj_error, min_tri_angle,
                                      od.ctypes.data_as(_u8p), pd.ctypes.data_as(_u8p),
                                      pe.ctypes.data_as(_dp), C.byref(nf))
    return nf.value, od[:O], pd[:P], pe[:P]


def filter_negative_depth(problem):
    O, P = len(problem.obs_image), len(problem.points)
    od, pd = np.zeros(max(O, 1), np.uint8), np.zeros(max(P, 1), np.uint8)
    nf = C.c_uint64(0)
    _filter_lib().ref_filter_negative_depth(C.byref(problem.struct), od.ctypes.data_as(_u8p),
                                            pd.ctypes.data_as(_u8p), C.byref(nf))
    return nf.value, od[:O], pd[:P]


# ---------------------------------------------------------------------------------------------
# the reference's Reconstruction::WriteText / ReadText (text model format, row f3), same library
# ---------------------------------------------------------------------------------------------
def model_write_text(problem, path, filter_thresholds=None):
    """The reference builds the reconstruction of a filters.FilterProblem through its own members
    (ids = index + 1, names image%06d.jpg), optionally runs its own FilterPoints3D
    (filter_thresholds = (max_reproj_error, min_tri_angle)), and writes it with WriteText."""
    os.makedirs(path, exist_ok=True)
    mx, ang = filter_thresholds if filter_thresholds is not None else (0.0, 0.0)
    _filter_lib().ref_model_write_text(C.byref(problem.struct), int(filter_thresholds is not None),
                                       mx, ang, path.encode())


def normalize(problem, extent=10.0, p0=0.1, p1=0.9, use_images=True):
    """The reference's Reconstruction::Normalize on the reconstruction of a FilterProblem:
    (tvecs [n, 3], points [p, 3]) afterwards (points with an empty track are not part of it)."""
    tv, pts = np.zeros_like(problem.tvecs), np.zeros_like(problem.points)
    _filter_lib().ref_normalize(C.byref(problem.struct), extent, p0, p1, int(use_images),
                                tv.ctypes.data_as(_dp), pts.ctypes.data_as(_dp))
    return tv, pts


def model_read_text(path, rewrite_to=None):
    """The reference's ReadText on directory ``path``, handed back as a dict of flat arrays sorted
    by id; ``rewrite_to``: the reference also writes what it has read into that directory."""
    L = _filter_lib()
    h = L.ref_model_read_text(path.encode())
    try:
        sz = (C.c_int64 * 6)()
        L.ref_model_sizes(h, sz)
        nc, ni, nl, npt, nt, nreg = (int(v) for v in sz)
        m = dict(
            cam_id=np.zeros(nc, np.int64), cam_model=np.zeros(nc, np.int32),
            cam_size=np.zeros((nc, 2), np.int64), cam_num_params=np.zeros(nc, np.int32),
            cam_params=np.zeros((nc, 12)), img_id=np.zeros(ni, np.int64), img_qvec=np.zeros((ni, 4)),
            img_tvec=np.zeros((ni, 3)), img_camera=np.zeros(ni, np.int64),
            img_name=np.zeros((ni, 64), np.uint8), line_start=np.zeros(ni + 1, np.int64),
            lines=np.zeros((nl, 3)), aligned=np.zeros(nl, np.uint8), line_point=np.zeros(nl, np.int64),
            pt_id=np.zeros(npt, np.int64), pt_xyz=np.zeros((npt, 3)), pt_color=np.zeros((npt, 3), np.uint8),
            pt_error=np.zeros(npt), track_start=np.zeros(npt + 1, np.int64),
            track_image=np.zeros(nt, np.int64), track_line=np.zeros(nt, np.int64))
        L.ref_model_dump(h, *[C.c_void_p(a.ctypes.data) for a in m.values()])
        m["num_reg_images"] = nreg
        m["img_name"] = [bytes(r).split(b"\0")[0].decode() for r in m["img_name"]]
        if rewrite_to is not None:
            os.makedirs(rewrite_to, exist_ok=True)
            L.ref_model_rewrite_text(h, rewrite_to.encode())
    finally:
        L.ref_model_free(h)
    return m


# ---------------------------------------------------------------------------------------------
# the reference's CorrespondenceGraph (src/base/correspondence_graph.cc), same library
# ---------------------------------------------------------------------------------------------
class CorrespondenceGraph:
    """The reference's own class behind the member names of
    privacy_preserving_sfm_b200.correspondence_graph.CorrespondenceGraph."""

    def __init__(self):
        L = self._L = _filter_lib()
        if not getattr(L, "_cg_declared", False):
            u32, u64 = C.c_uint32, C.c_uint64
            L.ref_cg_new.restype = C.c_void_p
            L.ref_cg_free.argtypes = [C.c_void_p]
            L.ref_cg_add_image.argtypes = [C.c_void_p, u32, u64]
            L.ref_cg_add_correspondences.argtypes = [C.c_void_p, u32, u32, C.c_void_p, u64]
            L.ref_cg_finalize.argtypes = [C.c_void_p]
            L.ref_cg_sizes.argtypes = [C.c_void_p, C.POINTER(u64)]
            L.ref_cg_image.argtypes = [C.c_void_p, u32, C.POINTER(u64)]
            L.ref_cg_num_between.argtypes = [C.c_void_p, u32, u32]
            L.ref_cg_num_between.restype = u64
            L.ref_cg_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, u64]
            L.ref_cg_pairs.restype = u64
            L.ref_cg_find.argtypes = [C.c_void_p, u32, u32, u64, C.c_void_p, u64]
            L.ref_cg_find.restype = u64
            L.ref_cg_between.argtypes = [C.c_void_p, u32, u32, C.c_void_p, u64]
            L.ref_cg_between.restype = u64
            L.ref_cg_line.argtypes = [C.c_void_p, u32, u32, C.POINTER(u64)]
            L._cg_declared = True
        self._h = L.ref_cg_new()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ref_cg_free(self._h)
            self._h = None

    def AddImage(self, image_id, num_lines):
        self._L.ref_cg_add_image(self._h, image_id, num_lines)

    def AddCorrespondences(self, image_id1, image_id2, matches):
        m = np.ascontiguousarray(matches, np.uint32).reshape(-1, 2)
        self._L.ref_cg_add_correspondences(self._h, image_id1, image_id2, m.ctypes.data, len(m))

    def Finalize(self):
        self._L.ref_cg_finalize(self._h)

    def _sizes(self):
        out = (C.c_uint64 * 2)()
        self._L.ref_cg_sizes(self._h, out)
        return int(out[0]), int(out[1])

    def NumImages(self):
        return self._sizes()[0]

    def NumImagePairs(self):
        return self._sizes()[1]

    def _image(self, image_id):
        out = (C.c_uint64 * 3)()
        self._L.ref_cg_image(self._h, image_id, out)
        return [int(v) for v in out]

    def ExistsImage(self, image_id):
        return bool(self._image(image_id)[0])

    def NumObservationsForImage(self, image_id):
        return self._image(image_id)[1]

    def NumCorrespondencesForImage(self, image_id):
        return self._image(image_id)[2]

    def NumCorrespondencesBetweenImages(self, image_id1=None, image_id2=None):
        if image_id1 is not None:
            return int(self._L.ref_cg_num_between(self._h, image_id1, image_id2))
        n = self.NumImagePairs()
        ids, num = np.zeros(max(n, 1), np.uint64), np.zeros(max(n, 1), np.uint32)
        self._L.ref_cg_pairs(self._h, ids.ctypes.data, num.ctypes.data, n)
        return {int(i): int(c) for i, c in zip(ids[:n], num[:n])}

    def FindTransitiveCorrespondences(self, image_id, line_idx, transitivity):
        cap = 4096
        while True:
            out = np.zeros((cap, 2), np.uint32)
            n = int(self._L.ref_cg_find(self._h, image_id, line_idx, transitivity, out.ctypes.data, cap))
            if n <= cap:
                return [tuple(r) for r in out[:n].tolist()]
            cap = n

    def FindCorrespondences(self, image_id, line_idx):
        return self.FindTransitiveCorrespondences(image_id, line_idx, 1)

    def FindCorrespondencesBetweenImages(self, image_id1, image_id2):
        cap = max(1, self.NumCorrespondencesBetweenImages(image_id1, image_id2))
        out = np.zeros((cap, 2), np.uint32)
        n = int(self._L.ref_cg_between(self._h, image_id1, image_id2, out.ctypes.data, cap))
        return [tuple(r) for r in out[:n].tolist()]

    def _line(self, image_id, line_idx):
        out = (C.c_uint64 * 2)()
        self._L.ref_cg_line(self._h, image_id, line_idx, out)
        return bool(out[0]), bool(out[1])

    def HasCorrespondences(self, image_id, line_idx):
        return self._line(image_id, line_idx)[0]

    def IsTwoViewObservation(self, image_id, line_idx):
        return self._line(image_id, line_idx)[1]
