"""Batched robust line triangulation: Python mirror of ``EstimateTriangulation``
(src/estimators/triangulation.h:117-147) for many tracks at once, over the C-ABI (CUDA)."""
import ctypes as C

import numpy as np

from . import binding
from .filters import FilterProblemStruct

_dp = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)

ANGULAR_ERROR, REPROJECTION_ERROR = 0, 1   # TriangulationEstimator::ResidualType


class EstimateTriangulationOptions(C.Structure):
    """EstimateTriangulationOptions + its RANSACOptions (triangulation.h:117-135); angles in
    radians.  ``exhaustive_threshold``: tracks up to this length use min_num_trials = C(n, 3)
    (src/sfm/incremental_triangulator.cc:527-531)."""
    _fields_ = [("min_tri_angle", C.c_double), ("residual_type", C.c_int32),
                ("max_error", C.c_double), ("min_inlier_ratio", C.c_double),
                ("confidence", C.c_double), ("dyn_num_trials_multiplier", C.c_double),
                ("min_num_trials", C.c_uint64), ("max_num_trials", C.c_uint64),
                ("exhaustive_threshold", C.c_int32)]

    def __init__(self, **kw):
        super().__init__()
        L = binding.load_library()
        L.ppsfm_triangulation_options_default.argtypes = [C.POINTER(EstimateTriangulationOptions)]
        L.ppsfm_triangulation_options_default.restype = None
        L.ppsfm_triangulation_options_default(C.byref(self))
        for k, v in kw.items():
            setattr(self, k, v)


def EstimateTriangulationBatch(ctx, tracks, options):
    """tracks: filters.FilterProblem (track-major; ``points`` unused).  Returns
    (success [T] bool, xyz [T, 3], inlier_mask [O] bool, num_trials [T])."""
    L = binding.load_library()
    L.ppsfm_estimate_triangulation_batch.argtypes = [
        C.c_void_p, C.POINTER(FilterProblemStruct), C.POINTER(EstimateTriangulationOptions), _dp,
        _u8p, _u8p, _u32p]
    T, O = len(tracks.points), len(tracks.obs_image)
    xyz = np.zeros((max(T, 1), 3))
    ok, mask = np.zeros(max(T, 1), np.uint8), np.zeros(max(O, 1), np.uint8)
    nt = np.zeros(max(T, 1), np.uint32)
    ctx._check(L.ppsfm_estimate_triangulation_batch(
        ctx._h, C.byref(tracks.struct), C.byref(options), xyz.ctypes.data_as(_dp),
        ok.ctypes.data_as(_u8p), mask.ctypes.data_as(_u8p), nt.ctypes.data_as(_u32p)))
    return ok[:T].astype(bool), xyz[:T], mask[:O].astype(bool), nt[:T]
