"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile liboracle.so (g++ only; a few seconds)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cc", ".h"))]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


class RansacOptions(C.Structure):
    _fields_ = [("max_error", C.c_double), ("min_inlier_ratio", C.c_double),
                ("confidence", C.c_double), ("dyn_num_trials_multiplier", C.c_double),
                ("min_num_trials", C.c_uint64), ("max_num_trials", C.c_uint64)]


class RansacReport(C.Structure):
    _fields_ = [("success", C.c_int32), ("num_trials", C.c_uint64), ("num_inliers", C.c_uint64),
                ("residual_sum", C.c_double), ("model", C.c_double * 12),
                ("best_trial", C.c_int64), ("best_model_idx", C.c_int32),
                ("num_models_scored", C.c_uint64)]


_lib = None
_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_set_prng_seed.argtypes = [C.c_uint32]
        L.orc_prng_peek.restype = C.c_uint32
        L.orc_line_residuals.argtypes = [_dp, _dp, C.c_size_t, _dp, _dp]
        L.orc_inlier_support.argtypes = [_dp, C.c_size_t, C.c_double, C.POINTER(C.c_uint64), _dp]
        L.orc_mestimator_support.argtypes = [_dp, C.c_size_t, C.c_double,
                                             C.POINTER(C.c_uint64), _dp]
        L.orc_compute_num_trials.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double]
        L.orc_compute_num_trials.restype = C.c_uint64
        L.orc_sample_table.argtypes = [C.c_size_t, C.c_size_t, _u32p]
        L.orc_re3q3.argtypes = [_dp, _dp]
        L.orc_re3q3.restype = C.c_int
        L.orc_re3q3_resultant.argtypes = [_dp, _dp, _dp]
        L.orc_re3q3_backsubstitute.argtypes = [_dp, C.c_double, _dp]
        L.orc_poly8_real_roots.argtypes = [_dp, _dp]
        L.orc_poly8_real_roots.restype = C.c_int
        L.orc_poly8_all_roots.argtypes = [_dp, _dp]
        L.orc_poly8_all_roots.restype = C.c_int
        L.orc_p6l_estimate.argtypes = [_dp, _u8p, _dp, _dp]
        L.orc_p6l_estimate.restype = C.c_int
        L.orc_rotation_matrix_to_quaternion.argtypes = [_dp, _dp]
        L.orc_ransac_p6l.argtypes = [_dp, _u8p, _dp, C.c_size_t, C.POINTER(RansacOptions),
                                     C.POINTER(RansacReport), _u8p]
        L.orc_estimate_absolute_pose_from_lines.argtypes = [
            _dp, _u8p, _dp, C.c_size_t, C.POINTER(RansacOptions), _dp, _dp,
            C.POINTER(C.c_uint64), _u8p, C.POINTER(RansacReport)]
        L.orc_estimate_absolute_pose_from_lines.restype = C.c_int
        L.orc_ransac_p6l_fixed_trials.argtypes = [_dp, _u8p, _dp, C.c_size_t, C.c_double,
                                                  C.c_uint64, C.POINTER(RansacReport)]
        L.orc_ransac_p6l_fixed_trials.restype = C.c_uint64
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_u8p)


def make_options(max_error, min_inlier_ratio=0.1, confidence=0.99, multiplier=3.0,
                 min_num_trials=0, max_num_trials=2**64 - 1):
    return RansacOptions(max_error, min_inlier_ratio, confidence, multiplier, min_num_trials,
                         max_num_trials)


def set_prng_seed(seed):
    lib().orc_set_prng_seed(seed)


def prng_peek():
    return int(lib().orc_prng_peek())


def line_residuals(lines, points, model):
    lines, lp = _d(lines)
    points, pp = _d(points)
    model, mp = _d(model)
    n = lines.shape[0]
    out = np.empty(n, dtype=np.float64)
    lib().orc_line_residuals(lp, pp, n, mp, out.ctypes.data_as(_dp))
    return out


def inlier_support(residuals, max_residual):
    residuals, rp = _d(residuals)
    cnt = C.c_uint64()
    s = C.c_double()
    lib().orc_inlier_support(rp, residuals.shape[0], max_residual, C.byref(cnt), C.byref(s))
    return int(cnt.value), float(s.value)


def mestimator_support(residuals, max_residual):
    residuals, rp = _d(residuals)
    cnt = C.c_uint64()
    s = C.c_double()
    lib().orc_mestimator_support(rp, residuals.shape[0], max_residual, C.byref(cnt), C.byref(s))
    return int(cnt.value), float(s.value)


def compute_num_trials(num_inliers, num_samples, confidence, multiplier):
    return int(lib().orc_compute_num_trials(num_inliers, num_samples, confidence, multiplier))


def sample_table(n, num_trials):
    out = np.empty((num_trials, 6), dtype=np.uint32)
    lib().orc_sample_table(n, num_trials, out.ctypes.data_as(_u32p))
    return out


def re3q3(coeffs):
    coeffs, cp = _d(coeffs)
    assert coeffs.shape == (3, 10)
    sol = np.zeros(24, dtype=np.float64)
    n = lib().orc_re3q3(cp, sol.ctypes.data_as(_dp))
    return sol.reshape(8, 3)[:n].copy()


def re3q3_resultant(P):
    """re3q3.h:84-150 on P (3x7) -> (a[33], c[9])."""
    P, pp = _d(np.asarray(P, np.float64).reshape(3, 7))
    a = np.zeros(33)
    c = np.zeros(9)
    lib().orc_re3q3_resultant(pp, a.ctypes.data_as(_dp), c.ctypes.data_as(_dp))
    return a, c


def re3q3_backsubstitute(a, x):
    """re3q3.h:177-188: (y, z) for the root x."""
    a, ap = _d(np.asarray(a, np.float64).reshape(33))
    yz = np.zeros(2)
    lib().orc_re3q3_backsubstitute(ap, float(x), yz.ctypes.data_as(_dp))
    return yz


def poly8_real_roots(c):
    c, cp = _d(c)
    out = np.zeros(8)
    n = lib().orc_poly8_real_roots(cp, out.ctypes.data_as(_dp))
    return out[:n].copy()


def poly8_all_roots(c):
    c, cp = _d(c)
    out = np.zeros(16)
    lib().orc_poly8_all_roots(cp, out.ctypes.data_as(_dp))
    return out[0::2] + 1j * out[1::2]


def p6l_estimate(lines6, aligned6, points6):
    lines6, lp = _d(lines6)
    points6, pp = _d(points6)
    aligned6, ap = _u8(aligned6)
    out = np.zeros((8, 12), dtype=np.float64)
    n = lib().orc_p6l_estimate(lp, ap, pp, out.ctypes.data_as(_dp))
    return out[:n].copy()


def rotation_matrix_to_quaternion(R):
    """R: 3x3 numpy (row/col semantic as usual); returns (w, x, y, z)."""
    Rc, rp = _d(np.asarray(R, dtype=np.float64).T)  # column-major flat
    q = np.zeros(4)
    lib().orc_rotation_matrix_to_quaternion(rp, q.ctypes.data_as(_dp))
    return q


def ransac_p6l(lines, aligned, points, options):
    lines, lp = _d(lines)
    points, pp = _d(points)
    aligned, ap = _u8(aligned)
    n = lines.shape[0]
    rep = RansacReport()
    mask = np.zeros(n, dtype=np.uint8)
    lib().orc_ransac_p6l(lp, ap, pp, n, C.byref(options), C.byref(rep), mask.ctypes.data_as(_u8p))
    return rep, mask


def estimate_absolute_pose_from_lines(lines, aligned, points, options):
    lines, lp = _d(lines)
    points, pp = _d(points)
    aligned, ap = _u8(aligned)
    n = lines.shape[0]
    rep = RansacReport()
    mask = np.zeros(n, dtype=np.uint8)
    q = np.zeros(4)
    t = np.zeros(3)
    ninl = C.c_uint64()
    ok = lib().orc_estimate_absolute_pose_from_lines(
        lp, ap, pp, n, C.byref(options), q.ctypes.data_as(_dp), t.ctypes.data_as(_dp),
        C.byref(ninl), mask.ctypes.data_as(_u8p), C.byref(rep))
    return bool(ok), q, t, int(ninl.value), mask, rep


def ransac_p6l_fixed_trials(lines, aligned, points, max_error, num_trials):
    lines, lp = _d(lines)
    points, pp = _d(points)
    aligned, ap = _u8(aligned)
    rep = RansacReport()
    scored = lib().orc_ransac_p6l_fixed_trials(lp, ap, pp, lines.shape[0], max_error, num_trials,
                                               C.byref(rep))
    return int(scored), rep


# ---------------------------------------------------------------------------------------------
# Bundle adjustment oracle (ba_oracle.h)
# ---------------------------------------------------------------------------------------------
_i32p = C.POINTER(C.c_int32)
BA_MAX_TRACE = 128


class BaProblem(C.Structure):
    _fields_ = [("num_images", C.c_int32), ("qvecs", _dp), ("tvecs", _dp), ("pose_flags", _u8p),
                ("image_camera", _i32p), ("num_cameras", C.c_int32), ("camera_model", _i32p),
                ("camera_params", _dp), ("num_points", C.c_int32), ("points", _dp),
                ("point_const", _u8p), ("num_obs", C.c_int64), ("obs_image", _i32p),
                ("obs_point", _i32p), ("obs_line", _dp), ("camera_const", _u8p)]


class BaOptions(C.Structure):
    _fields_ = [("loss_type", C.c_int32), ("loss_scale", C.c_double),
                ("max_num_iterations", C.c_int32), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int32),
                ("initial_trust_region_radius", C.c_double),
                ("max_trust_region_radius", C.c_double), ("min_trust_region_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("jacobi_scaling", C.c_int32),
                ("num_threads", C.c_int32), ("refine_focal_length", C.c_int32),
                ("refine_principal_point", C.c_int32), ("refine_extra_params", C.c_int32)]


class BaSummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("num_successful_steps", C.c_int32), ("num_unsuccessful_steps", C.c_int32),
                ("termination_type", C.c_int32), ("num_residuals", C.c_int64),
                ("num_residuals_reduced", C.c_int64),
                ("num_effective_parameters_reduced", C.c_int32), ("total_time_s", C.c_double),
                ("jacobian_time_s", C.c_double), ("linear_solver_time_s", C.c_double),
                ("final_gradient_max_norm", C.c_double), ("trace_len", C.c_int32),
                ("trace_cost", C.c_double * BA_MAX_TRACE),
                ("trace_radius", C.c_double * BA_MAX_TRACE),
                ("trace_accepted", C.c_int32 * BA_MAX_TRACE)]


_ba_ready = False


def _ba_lib():
    global _ba_ready
    L = lib()
    if not _ba_ready:
        L.orc_ba_options_default.argtypes = [C.POINTER(BaOptions)]
        L.orc_ba_solve.argtypes = [C.POINTER(BaProblem), C.POINTER(BaOptions),
                                   C.POINTER(BaSummary)]
        L.orc_ba_solve.restype = C.c_int
        L.orc_line_cost.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_line_cost_intr.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_line_cost_tangent.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
        L.orc_ba_cost.argtypes = [C.POINTER(BaProblem), C.POINTER(BaOptions)]
        L.orc_ba_cost.restype = C.c_double
        L.orc_quaternion_plus.argtypes = [_dp, _dp, _dp]
        L.orc_refine_absolute_pose.argtypes = [_dp, _dp, _u8p, C.c_size_t, C.c_int, _dp,
                                               C.c_double, C.c_int, C.c_double, _dp, _dp,
                                               C.POINTER(BaSummary)]
        L.orc_refine_absolute_pose.restype = C.c_int
        _ba_ready = True
    return L


def ba_default_options(**kw):
    o = BaOptions()
    _ba_lib().orc_ba_options_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class BaArrays:
    """Owns contiguous numpy buffers of a BA problem and the ctypes struct pointing at them."""

    def __init__(self, qvecs, tvecs, points, obs_image, obs_point, obs_line, camera_model,
                 camera_params, image_camera=None, pose_flags=None, point_const=None,
                 struct_cls=None, camera_const=None):
        self.qvecs = np.ascontiguousarray(qvecs, dtype=np.float64).copy()
        self.tvecs = np.ascontiguousarray(tvecs, dtype=np.float64).copy()
        self.points = np.ascontiguousarray(points, dtype=np.float64).copy()
        ni, npnt = self.qvecs.shape[0], self.points.shape[0]
        self.obs_image = np.ascontiguousarray(obs_image, dtype=np.int32)
        self.obs_point = np.ascontiguousarray(obs_point, dtype=np.int32)
        self.obs_line = np.ascontiguousarray(obs_line, dtype=np.float64)
        self.camera_model = np.ascontiguousarray(np.atleast_1d(camera_model), dtype=np.int32)
        cp = np.atleast_2d(np.asarray(camera_params, dtype=np.float64))
        self.camera_params = np.zeros((cp.shape[0], 12))
        self.camera_params[:, :cp.shape[1]] = cp
        self.image_camera = np.ascontiguousarray(
            image_camera if image_camera is not None else np.zeros(ni), dtype=np.int32)
        self.pose_flags = np.ascontiguousarray(
            pose_flags if pose_flags is not None else np.zeros(ni), dtype=np.uint8)
        self.point_const = np.ascontiguousarray(
            point_const if point_const is not None else np.zeros(npnt), dtype=np.uint8)
        self.camera_const = np.ascontiguousarray(
            camera_const if camera_const is not None else np.zeros(self.camera_model.shape[0]),
            dtype=np.uint8)
        cls = struct_cls or BaProblem
        p = cls()
        p.num_images = ni
        p.qvecs = self.qvecs.ctypes.data_as(_dp)
        p.tvecs = self.tvecs.ctypes.data_as(_dp)
        p.pose_flags = self.pose_flags.ctypes.data_as(_u8p)
        p.image_camera = self.image_camera.ctypes.data_as(_i32p)
        p.num_cameras = self.camera_model.shape[0]
        p.camera_model = self.camera_model.ctypes.data_as(_i32p)
        p.camera_params = self.camera_params.ctypes.data_as(_dp)
        p.num_points = npnt
        p.points = self.points.ctypes.data_as(_dp)
        p.point_const = self.point_const.ctypes.data_as(_u8p)
        p.num_obs = self.obs_image.shape[0]
        p.obs_image = self.obs_image.ctypes.data_as(_i32p)
        p.obs_point = self.obs_point.ctypes.data_as(_i32p)
        p.obs_line = self.obs_line.ctypes.data_as(_dp)
        if hasattr(p, "camera_const"):
            p.camera_const = self.camera_const.ctypes.data_as(_u8p)
        self.struct = p


def ba_solve(arrays, options):
    s = BaSummary()
    ok = _ba_lib().orc_ba_solve(C.byref(arrays.struct), C.byref(options), C.byref(s))
    return bool(ok), s


def ba_cost(arrays, options):
    return float(_ba_lib().orc_ba_cost(C.byref(arrays.struct), C.byref(options)))


def line_cost(model, cam_params, line, q, t, X):
    cam = np.zeros(12)
    cam[:len(cam_params)] = cam_params
    line, lp = _d(line)
    q, qp = _d(q)
    t, tp = _d(t)
    X, xp = _d(X)
    r, jq, jt, jx = np.zeros(2), np.zeros((2, 4)), np.zeros((2, 3)), np.zeros((2, 3))
    _ba_lib().orc_line_cost(model, cam.ctypes.data_as(_dp), lp, qp, tp, xp,
                            r.ctypes.data_as(_dp), jq.ctypes.data_as(_dp),
                            jt.ctypes.data_as(_dp), jx.ctypes.data_as(_dp))
    return r, jq, jt, jx


def line_cost_intr(model, cam_params, line, q, t, X):
    """The block (2; 4, 3, 3, k) of intrinsics refinement: r, Jq, Jt, JX, Jcamera[2 x 12]."""
    cam = np.zeros(12)
    cam[:len(cam_params)] = cam_params
    line, lp = _d(line)
    q, qp = _d(q)
    t, tp = _d(t)
    X, xp = _d(X)
    r, jq, jt, jx = np.zeros(2), np.zeros((2, 4)), np.zeros((2, 3)), np.zeros((2, 3))
    jc = np.zeros((2, 12))
    _ba_lib().orc_line_cost_intr(model, cam.ctypes.data_as(_dp), lp, qp, tp, xp,
                                 r.ctypes.data_as(_dp), jq.ctypes.data_as(_dp),
                                 jt.ctypes.data_as(_dp), jx.ctypes.data_as(_dp),
                                 jc.ctypes.data_as(_dp))
    return r, jq, jt, jx, jc


def line_cost_tangent(model, cam_params, line, q, t, X):
    cam = np.zeros(12)
    cam[:len(cam_params)] = cam_params
    line, lp = _d(line)
    q, qp = _d(q)
    t, tp = _d(t)
    X, xp = _d(X)
    r, jc, jx = np.zeros(2), np.zeros((2, 6)), np.zeros((2, 3))
    _ba_lib().orc_line_cost_tangent(model, cam.ctypes.data_as(_dp), lp, qp, tp, xp,
                                    r.ctypes.data_as(_dp), jc.ctypes.data_as(_dp),
                                    jx.ctypes.data_as(_dp))
    return r, jc, jx


def quaternion_plus(q, delta):
    q, qp = _d(q)
    delta, dp_ = _d(delta)
    out = np.zeros(4)
    _ba_lib().orc_quaternion_plus(qp, dp_, out.ctypes.data_as(_dp))
    return out


def refine_absolute_pose(lines, points, mask, model, cam_params, qvec, tvec,
                         gradient_tolerance=1.0, max_num_iterations=100, loss_scale=1.0):
    lines, lp = _d(lines)
    points, pp = _d(points)
    mask, mp = _u8(mask)
    cam = np.zeros(12)
    cam[:len(cam_params)] = cam_params
    q = np.array(qvec, dtype=np.float64)
    t = np.array(tvec, dtype=np.float64)
    s = BaSummary()
    ok = _ba_lib().orc_refine_absolute_pose(
        lp, pp, mp, lines.shape[0], model, cam.ctypes.data_as(_dp), gradient_tolerance,
        max_num_iterations, loss_scale, q.ctypes.data_as(_dp), t.ctypes.data_as(_dp), C.byref(s))
    return bool(ok), q, t, s


# ---- post-BA filters (filter_oracle.cc) ------------------------------------------------------
def filter_points3d(problem, max_reproj_error, min_tri_angle, point_error=None):
    """problem: privacy_preserving_sfm_b200.filters.FilterProblem (same struct layout).
    Returns (num_filtered, obs_deleted, point_deleted, point_error, squared_errors)."""
    L = lib()
    O, P = len(problem.obs_image), len(problem.points)
    od, pd = np.zeros(max(O, 1), np.uint8), np.zeros(max(P, 1), np.uint8)
    pe = np.full(max(P, 1), -1.0) if point_error is None else np.array(point_error, np.float64)
    sq = np.full(max(O, 1), -1.0)
    nf = C.c_uint64(0)
    L.orc_filter_points3d.argtypes = [C.c_void_p, C.c_double, C.c_double, _u8p, _u8p, _dp,
                                      C.POINTER(C.c_uint64), _dp]
    L.orc_filter_points3d(C.byref(problem.struct), max_reproj_error, min_tri_angle,
                          od.ctypes.data_as(_u8p), pd.ctypes.data_as(_u8p),
                          pe.ctypes.data_as(_dp), C.byref(nf), sq.ctypes.data_as(_dp))
    return nf.value, od[:O], pd[:P], pe[:P], sq[:O]


def filter_negative_depth(problem):
    L = lib()
    O = len(problem.obs_image)
    P = len(problem.points)
    od = np.zeros(max(O, 1), np.uint8)
    pd = np.zeros(max(P, 1), np.uint8)
    nf = C.c_uint64(0)
    L.orc_filter_negative_depth.argtypes = [C.c_void_p, _u8p, _u8p, C.POINTER(C.c_uint64)]
    L.orc_filter_negative_depth(C.byref(problem.struct), od.ctypes.data_as(_u8p),
                                pd.ctypes.data_as(_u8p), C.byref(nf))
    return nf.value, od[:O], pd[:P]


# ---- batched line triangulation (triangulation_oracle.cc) -------------------------------------
def estimate_triangulation_batch(tracks, options):
    """tracks: filters.FilterProblem, options: triangulation.EstimateTriangulationOptions (same
    struct layouts).  Returns (success, xyz, inlier_mask, num_trials)."""
    L = lib()
    T, O = len(tracks.points), len(tracks.obs_image)
    xyz = np.zeros((max(T, 1), 3))
    ok, mask = np.zeros(max(T, 1), np.uint8), np.zeros(max(O, 1), np.uint8)
    nt = np.zeros(max(T, 1), np.uint32)
    L.orc_estimate_triangulation_batch.argtypes = [C.c_void_p, C.c_void_p, _dp, _u8p, _u8p, _u32p]
    L.orc_estimate_triangulation_batch(C.byref(tracks.struct), C.byref(options),
                                       xyz.ctypes.data_as(_dp), ok.ctypes.data_as(_u8p),
                                       mask.ctypes.data_as(_u8p), nt.ctypes.data_as(_u32p))
    return ok[:T].astype(bool), xyz[:T], mask[:O].astype(bool), nt[:T]
