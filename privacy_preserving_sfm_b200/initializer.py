"""Four-view initialisation from lifted lines: Python mirror of
``init::initialize_reconstruction`` (src/init/initializer.h:103-108) over the C-ABI.  With a
context the candidate models of both LO-MSAC loops are scored on the GPU
(``ppsfm_initialize_reconstruction_gpu``, identical results); without one everything runs on the
host."""
import ctypes as C

import numpy as np

from . import binding

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)


class InitOptions(C.Structure):
    """init::InitOptions (src/init/initializer.h:49-58)."""
    _fields_ = [("min_tri_angle", C.c_double), ("min_num_inliers", C.c_double),
                ("max_error", C.c_double)]

    def __init__(self, **kw):
        super().__init__()
        binding.load_library().ppsfm_init_options_default(C.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


class InitReport(C.Structure):
    _fields_ = [("num_aligned", C.c_int32), ("num_unaligned", C.c_int32),
                ("inliers_2d", C.c_int32), ("inliers_3d", C.c_int32),
                ("iterations_2d", C.c_uint32), ("iterations_3d", C.c_uint32),
                ("mean_tri_angle_deg", C.c_double)]


def _declare(L):
    if getattr(L, "_init_declared", False):
        return
    L.ppsfm_init_options_default.argtypes = [C.POINTER(InitOptions)]
    L.ppsfm_init_options_default.restype = None
    L.ppsfm_initialize_reconstruction.argtypes = [_dp, _u8p, C.c_size_t, _dp,
                                                  C.POINTER(InitOptions), _dp, _dp,
                                                  C.POINTER(InitReport)]
    L.ppsfm_initialize_reconstruction.restype = C.c_int
    L.ppsfm_initialize_reconstruction_gpu.argtypes = [C.c_void_p, _dp, _u8p, C.c_size_t, _dp,
                                                      C.POINTER(InitOptions), _dp, _dp,
                                                      C.POINTER(InitReport),
                                                      C.POINTER(C.c_int64)]
    L.ppsfm_initialize_reconstruction_gpu.restype = C.c_int
    L._init_declared = True


def initialize_reconstruction(lines, aligned, gravity, options=None, ctx=None):
    """lines [4, n, 3], aligned [4, n], gravity [4, 3] -> (ok, poses [4, 3, 4], inlier_ratio,
    report).  Raises ValueError where the reference CHECK-aborts.  ctx: a Context -> models are
    scored on its GPU (report.gpu_launches = kernels launched)."""
    L = binding.load_library()
    _declare(L)
    lines = np.ascontiguousarray(lines, np.float64)
    aligned = np.ascontiguousarray(aligned, np.uint8)
    gravity = np.ascontiguousarray(gravity, np.float64)
    if lines.ndim != 3 or lines.shape[0] != 4 or lines.shape[2] != 3:
        raise ValueError("lines must be [4, n, 3]")
    n = lines.shape[1]
    if aligned.shape != (4, n) or gravity.shape != (4, 3):
        raise ValueError("aligned must be [4, n], gravity [4, 3]")
    opt = options if options is not None else InitOptions()
    poses = np.zeros((4, 3, 4))
    ratio = C.c_double(0.0)
    rep = InitReport()
    args = (lines.ctypes.data_as(_dp), aligned.ctypes.data_as(_u8p), n,
            gravity.ctypes.data_as(_dp), C.byref(opt), poses.ctypes.data_as(_dp), C.byref(ratio),
            C.byref(rep))
    if ctx is not None:
        launches = C.c_int64(0)
        rc = L.ppsfm_initialize_reconstruction_gpu(ctx._h, *args, C.byref(launches))
        rep.gpu_launches = int(launches.value)
        if rc == -2:  # PPSFM_ERR_CUDA
            raise binding.PpsfmError("initialize_reconstruction: CUDA error "
                                     + ctx._L.ppsfm_last_error(ctx._h).decode())
    else:
        rc = L.ppsfm_initialize_reconstruction(*args)
    if rc < 0:
        raise ValueError("initialize_reconstruction: contract violation (the reference aborts)")
    return rc == binding.PPSFM_OK, poses, ratio.value, rep
