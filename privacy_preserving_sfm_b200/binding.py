"""ctypes binding of libppsfm_b200.so (include/ppsfm_b200.h).  No torch, no CPU fallback."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libppsfm_b200.so")

PPSFM_OK = 0
PPSFM_NO_SOLUTION = 1


class PpsfmError(RuntimeError):
    pass


def library_path():
    return _LIB


def build_library(force=False):
    """nvcc-compile csrc/ for sm_100a into libppsfm_b200.so (cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    cmd = ["make", "-C", csrc, "-s"]
    if force:
        cmd.append("-B")
    subprocess.check_call(cmd)
    return _LIB


class RANSACOptions(C.Structure):
    """RANSACOptions, src/optim/ransac.h:47-76 (same fields and defaults)."""
    _fields_ = [("max_error", C.c_double), ("min_inlier_ratio", C.c_double),
                ("confidence", C.c_double), ("dyn_num_trials_multiplier", C.c_double),
                ("min_num_trials", C.c_uint64), ("max_num_trials", C.c_uint64)]

    def __init__(self, max_error=0.0, min_inlier_ratio=0.1, confidence=0.99,
                 dyn_num_trials_multiplier=3.0, min_num_trials=0, max_num_trials=2**64 - 1):
        super().__init__(max_error, min_inlier_ratio, confidence, dyn_num_trials_multiplier,
                         min_num_trials, max_num_trials)

    def Check(self):
        # src/optim/ransac.h:68-75 (CHECK_* abort in the reference -> exception here)
        if not self.max_error > 0:
            raise PpsfmError("CHECK_GT(max_error, 0)")
        if not 0 <= self.min_inlier_ratio <= 1:
            raise PpsfmError("CHECK min_inlier_ratio in [0, 1]")
        if not 0 <= self.confidence <= 1:
            raise PpsfmError("CHECK confidence in [0, 1]")
        if self.min_num_trials > self.max_num_trials:
            raise PpsfmError("CHECK_LE(min_num_trials, max_num_trials)")


class RansacReport(C.Structure):
    """RANSAC<P6LEstimator>::Report, src/optim/ransac.h:82-99."""
    _fields_ = [("success", C.c_int32), ("num_trials", C.c_uint64), ("num_inliers", C.c_uint64),
                ("residual_sum", C.c_double), ("model", C.c_double * 12),
                ("best_trial", C.c_int64), ("best_model_idx", C.c_int32),
                ("num_models_scored", C.c_uint64)]


class RansacTiming(C.Structure):
    _fields_ = [("solve_ms", C.c_double), ("score_ms", C.c_double), ("exact_ms", C.c_double),
                ("total_ms", C.c_double), ("score_pairs", C.c_uint64),
                ("score_launches", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("comm_ms", C.c_double)]


_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_lib = None


def load_library():
    """Loads libppsfm_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise PpsfmError(
            f"{_LIB} not found: build it with __graft_entry__.build() or "
            "`make -C privacy_preserving_sfm_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(_LIB)
    vp = C.c_void_p
    L.ppsfm_version.restype = C.c_char_p
    L.ppsfm_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.ppsfm_ctx_destroy.argtypes = [vp]
    L.ppsfm_ctx_destroy.restype = None
    L.ppsfm_last_error.argtypes = [vp]
    L.ppsfm_last_error.restype = C.c_char_p
    L.ppsfm_set_prng_seed.argtypes = [vp, C.c_uint32]
    L.ppsfm_set_prng_seed.restype = None
    L.ppsfm_prng_peek.argtypes = [vp]
    L.ppsfm_prng_peek.restype = C.c_uint32
    L.ppsfm_ransac_options_default.argtypes = [C.POINTER(RANSACOptions)]
    L.ppsfm_ransac_options_default.restype = None
    L.ppsfm_compute_num_trials.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double]
    L.ppsfm_compute_num_trials.restype = C.c_uint64
    L.ppsfm_sample_table.argtypes = [vp, C.c_size_t, C.c_size_t, _u32p]
    L.ppsfm_line_residuals.argtypes = [vp, _dp, _dp, C.c_size_t, _dp, C.c_size_t, C.c_double,
                                       _dp, _u64p, _dp]
    L.ppsfm_p6l_solve_batch.argtypes = [vp, _dp, _u8p, _dp, C.c_size_t, _u32p, C.c_size_t, _dp,
                                        _i32p]
    L.ppsfm_ransac_p6l.argtypes = [vp, _dp, _u8p, _dp, C.c_size_t, C.POINTER(RANSACOptions),
                                   C.POINTER(RansacReport), _u8p]
    L.ppsfm_estimate_absolute_pose_from_lines.argtypes = [
        vp, _dp, _u8p, _dp, C.c_size_t, C.POINTER(RANSACOptions), _dp, _dp, _u64p, _u8p,
        C.POINTER(RansacReport)]
    L.ppsfm_corr_upload.argtypes = [vp, _dp, _u8p, _dp, C.c_size_t, C.POINTER(vp)]
    L.ppsfm_corr_free.argtypes = [vp, vp]
    L.ppsfm_corr_free.restype = None
    L.ppsfm_ransac_p6l_resident.argtypes = [vp, vp, C.POINTER(RANSACOptions),
                                            C.POINTER(RansacReport), _u8p]
    L.ppsfm_ransac_p6l_resident_sharded.argtypes = [vp, vp, C.POINTER(RANSACOptions),
                                                    C.POINTER(RansacReport), _u8p]
    L.ppsfm_ransac_p6l_sharded.argtypes = [vp, _dp, _u8p, _dp, C.c_size_t,
                                           C.POINTER(RANSACOptions), C.POINTER(RansacReport), _u8p]
    L.ppsfm_get_ransac_timing.argtypes = [vp, C.POINTER(RansacTiming)]
    L.ppsfm_get_ransac_timing.restype = None
    L.ppsfm_comm_get_unique_id.argtypes = [vp, C.c_char_p]
    L.ppsfm_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.ppsfm_comm_destroy.argtypes = [vp]
    L.ppsfm_comm_destroy.restype = None
    L.ppsfm_comm_rank.argtypes = [vp]
    L.ppsfm_comm_world_size.argtypes = [vp]
    L.ppsfm_comm_allreduce_sum_host.argtypes = [vp, _dp, C.c_size_t]
    L.ppsfm_bench_fp64_peak.argtypes = [vp, _dp, _dp]
    L.ppsfm_bench_fp32_peak.argtypes = [vp, _dp]
    L.ppsfm_bench_l2_flush.argtypes = [vp, C.c_size_t]
    L.ppsfm_bench_hbm_rw_peak.argtypes = [vp, _dp, _dp]
    _lib = L
    return L


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _u8(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(_u8p)


class Correspondences:
    """Handle to a correspondence set resident in HBM (ppsfm_corr)."""

    def __init__(self, ctx, handle, n):
        self._ctx, self._h, self.n = ctx, handle, n

    def free(self):
        if self._h is not None:
            load_library().ppsfm_corr_free(self._ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One context per host thread / per GPU (ppsfm_ctx)."""

    def __init__(self, device=0):
        L = load_library()
        h = C.c_void_p()
        rc = L.ppsfm_ctx_create(device, C.byref(h))
        if rc != PPSFM_OK:
            raise PpsfmError(f"ppsfm_ctx_create(device={device}) failed (rc={rc}): no usable CUDA "
                             "device — this library has no CPU fallback")
        self._h = h
        self._L = L

    def close(self):
        if self._h is not None:
            self._L.ppsfm_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow_no_solution=False):
        if rc == PPSFM_OK or (allow_no_solution and rc == PPSFM_NO_SOLUTION):
            return rc
        raise PpsfmError(f"rc={rc}: {self._L.ppsfm_last_error(self._h).decode()}")

    # -- PRNG (util/random.h) ----------------------------------------------------------------
    def set_prng_seed(self, seed=0):
        self._L.ppsfm_set_prng_seed(self._h, seed)

    def prng_peek(self):
        return int(self._L.ppsfm_prng_peek(self._h))

    def sample_table(self, n, num_trials):
        out = np.empty((num_trials, 6), dtype=np.uint32)
        self._check(self._L.ppsfm_sample_table(self._h, n, num_trials, out.ctypes.data_as(_u32p)))
        return out

    # -- kernels -----------------------------------------------------------------------------
    def line_residuals(self, lines, points, models, max_residual, want_residuals=True):
        lines, lp = _d(lines)
        points, pp = _d(points)
        models, mp = _d(np.asarray(models, dtype=np.float64).reshape(-1, 12))
        n, k = lines.shape[0], models.shape[0]
        res = np.empty((k, n), dtype=np.float64) if want_residuals else None
        cnt = np.zeros(k, dtype=np.uint64)
        sm = np.zeros(k, dtype=np.float64)
        self._check(self._L.ppsfm_line_residuals(
            self._h, lp, pp, n, mp, k, max_residual,
            res.ctypes.data_as(_dp) if want_residuals else None,
            cnt.ctypes.data_as(_u64p), sm.ctypes.data_as(_dp)))
        return res, cnt, sm

    def score_models(self, lines, points, models, max_residual):
        """Inlier counts through the RANSAC scoring kernel (test hook)."""
        lines, lp = _d(lines)
        points, pp = _d(points)
        models, mp = _d(np.asarray(models, dtype=np.float64).reshape(-1, 12))
        cnt = np.zeros(models.shape[0], dtype=np.uint32)
        self._L.ppsfm_score_models.argtypes = [C.c_void_p, _dp, _dp, C.c_size_t, _dp, C.c_size_t,
                                               C.c_double, _u32p]
        self._check(self._L.ppsfm_score_models(self._h, lp, pp, lines.shape[0], mp,
                                               models.shape[0], max_residual,
                                               cnt.ctypes.data_as(_u32p)))
        return cnt

    def p6l_solve_batch(self, lines, aligned, points, sample_idx):
        lines, lp = _d(lines)
        points, pp = _d(points)
        aligned, ap = _u8(aligned)
        sample_idx = np.ascontiguousarray(sample_idx, dtype=np.uint32).reshape(-1, 6)
        h = sample_idx.shape[0]
        models = np.zeros((h, 8, 12), dtype=np.float64)
        nm = np.zeros(h, dtype=np.int32)
        self._check(self._L.ppsfm_p6l_solve_batch(
            self._h, lp, ap, pp, lines.shape[0], sample_idx.ctypes.data_as(_u32p), h,
            models.ctypes.data_as(_dp), nm.ctypes.data_as(_i32p)))
        return models, nm

    def upload(self, lines, aligned, points):
        lines, lp = _d(lines)
        points, pp = _d(points)
        aligned, ap = _u8(aligned)
        h = C.c_void_p()
        self._check(self._L.ppsfm_corr_upload(self._h, lp, ap, pp, lines.shape[0], C.byref(h)))
        return Correspondences(self, h, lines.shape[0])

    def ransac_p6l(self, lines, aligned, points, options, want_mask=True):
        lines, lp = _d(lines)
        points, pp = _d(points)
        aligned, ap = _u8(aligned)
        n = lines.shape[0]
        rep = RansacReport()
        mask = np.zeros(n, dtype=np.uint8)
        self._check(self._L.ppsfm_ransac_p6l(self._h, lp, ap, pp, n, C.byref(options),
                                             C.byref(rep),
                                             mask.ctypes.data_as(_u8p) if want_mask else None))
        return rep, mask

    def ransac_p6l_resident(self, corr, options, want_mask=True, mask_out=None):
        rep = RansacReport()
        mask = mask_out if mask_out is not None else np.zeros(corr.n, dtype=np.uint8)
        self._check(self._L.ppsfm_ransac_p6l_resident(
            self._h, corr._h, C.byref(options), C.byref(rep),
            mask.ctypes.data_as(_u8p) if want_mask else None))
        return rep, mask

    def estimate_absolute_pose_from_lines(self, lines, aligned, points, options):
        lines, lp = _d(lines)
        points, pp = _d(points)
        aligned, ap = _u8(aligned)
        n = lines.shape[0]
        rep = RansacReport()
        mask = np.zeros(n, dtype=np.uint8)
        q = np.zeros(4)
        t = np.zeros(3)
        ninl = C.c_uint64()
        rc = self._check(self._L.ppsfm_estimate_absolute_pose_from_lines(
            self._h, lp, ap, pp, n, C.byref(options), q.ctypes.data_as(_dp),
            t.ctypes.data_as(_dp), C.byref(ninl), mask.ctypes.data_as(_u8p), C.byref(rep)),
            allow_no_solution=True)
        return rc == PPSFM_OK, q, t, int(ninl.value), mask, rep

    # -- multi-GPU ---------------------------------------------------------------------------
    def ransac_p6l_sharded(self, lines, aligned, points, options, want_mask=True):
        """ONE RANSAC call sharded over the communicator's GPUs (collective: every rank passes the
        same set, options and generator state and gets the single-GPU call's report back)."""
        lines, lp = _d(lines)
        points, pp = _d(points)
        aligned, ap = _u8(aligned)
        n = lines.shape[0]
        rep = RansacReport()
        mask = np.zeros(n, dtype=np.uint8)
        self._check(self._L.ppsfm_ransac_p6l_sharded(
            self._h, lp, ap, pp, n, C.byref(options), C.byref(rep),
            mask.ctypes.data_as(_u8p) if want_mask else None))
        return rep, mask

    def ransac_p6l_resident_sharded(self, corr, options, want_mask=True, mask_out=None):
        rep = RansacReport()
        mask = mask_out if mask_out is not None else np.zeros(corr.n, dtype=np.uint8)
        self._check(self._L.ppsfm_ransac_p6l_resident_sharded(
            self._h, corr._h, C.byref(options), C.byref(rep),
            mask.ctypes.data_as(_u8p) if want_mask else None))
        return rep, mask

    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self._check(self._L.ppsfm_comm_get_unique_id(self._h, buf))
        return buf.raw

    def comm_init(self, world_size, rank, unique_id):
        self._check(self._L.ppsfm_comm_init(self._h, world_size, rank, unique_id))

    def comm_init_from_torch(self, dist):
        """Initialises the NCCL communicator of this context from an initialised
        torch.distributed process group (any backend): rank 0's id is broadcast as an object."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self.comm_init(world, rank, box[0])

    def comm_rank(self):
        return int(self._L.ppsfm_comm_rank(self._h))

    def comm_world_size(self):
        return int(self._L.ppsfm_comm_world_size(self._h))

    def comm_allreduce_sum(self, array):
        a = np.ascontiguousarray(array, dtype=np.float64).copy()
        self._check(self._L.ppsfm_comm_allreduce_sum_host(self._h, a.ctypes.data_as(_dp), a.size))
        return a

    def bench_fp64_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._L.ppsfm_bench_fp64_peak(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def bench_fp32_peak(self):
        a = C.c_double()
        self._check(self._L.ppsfm_bench_fp32_peak(self._h, C.byref(a)))
        return float(a.value)

    def bench_hbm_rw_peak(self):
        a, b = C.c_double(), C.c_double()
        self._check(self._L.ppsfm_bench_hbm_rw_peak(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def bench_l2_flush(self, nbytes=512 << 20):
        self._check(self._L.ppsfm_bench_l2_flush(self._h, nbytes))

    def ransac_timing(self):
        t = RansacTiming()
        self._L.ppsfm_get_ransac_timing(self._h, C.byref(t))
        return t


_default_ctx = None


def ransac_shard_models(num_models, rank, world):
    """Host-only: models of a wave that `rank` of `world` scores in a sharded RANSAC call."""
    L = load_library()
    L.ppsfm_ransac_shard_models.argtypes = [C.c_uint64, C.c_int, C.c_int]
    L.ppsfm_ransac_shard_models.restype = C.c_uint64
    return int(L.ppsfm_ransac_shard_models(num_models, rank, world))


def default_context():
    """Lazily created context on cuda:LOCAL_RANK (or 0)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx
