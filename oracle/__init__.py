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
