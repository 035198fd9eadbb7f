"""Post-bundle-adjustment filters on a track-major view of the reconstruction: Python mirror of
``Reconstruction::FilterPoints3D`` / ``FilterObservationsWithNegativeDepth``
(src/base/reconstruction.cc:425-460, 594-719) over the C-ABI (CUDA kernels, no CPU fallback)."""
import ctypes as C

import numpy as np

from . import binding

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


class FilterProblemStruct(C.Structure):
    _fields_ = [("num_images", C.c_int32), ("qvecs", _dp), ("tvecs", _dp), ("image_camera", _ip),
                ("num_cameras", C.c_int32), ("camera_model", _ip), ("camera_params", _dp),
                ("camera_width", _ip), ("camera_height", _ip), ("num_points", C.c_int32),
                ("points", _dp), ("track_start", _lp), ("num_obs", C.c_int64), ("obs_image", _ip),
                ("obs_line", _dp), ("obs_aligned", _u8p)]


class FilterProblem:
    """Owns contiguous arrays and the C struct that points into them.  Observations must be
    grouped by point: ``track_start[p] .. track_start[p+1]`` are the track of point p."""

    def __init__(self, qvecs, tvecs, image_camera, camera_model, camera_params, camera_size,
                 points, track_start, obs_image, obs_line, obs_aligned):
        f64 = lambda a: np.ascontiguousarray(a, np.float64)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        self.qvecs, self.tvecs, self.points, self.obs_line = f64(qvecs), f64(tvecs), f64(points), f64(obs_line)
        self.image_camera, self.camera_model, self.obs_image = i32(image_camera), i32(camera_model), i32(obs_image)
        ncam = len(self.camera_model)
        prm = np.zeros((ncam, 12))
        for i, p in enumerate(camera_params):
            prm[i, :len(p)] = p
        self.camera_params = prm
        size = np.asarray(camera_size, np.int32).reshape(ncam, 2)
        self.camera_width, self.camera_height = i32(size[:, 0]), i32(size[:, 1])
        self.track_start = np.ascontiguousarray(track_start, np.int64)
        self.obs_aligned = np.ascontiguousarray(obs_aligned, np.uint8)
        s = FilterProblemStruct()
        s.num_images, s.num_cameras = len(self.qvecs), ncam
        s.num_points, s.num_obs = len(self.points), len(self.obs_image)
        for name, typ in (("qvecs", _dp), ("tvecs", _dp), ("image_camera", _ip),
                          ("camera_model", _ip), ("camera_params", _dp), ("camera_width", _ip),
                          ("camera_height", _ip), ("points", _dp), ("track_start", _lp),
                          ("obs_image", _ip), ("obs_line", _dp), ("obs_aligned", _u8p)):
            setattr(s, name, getattr(self, name).ctypes.data_as(typ))
        self.struct = s


def _declare(L):
    if getattr(L, "_filter_declared", False):
        return
    P = C.POINTER(FilterProblemStruct)
    L.ppsfm_filter_points3d.argtypes = [C.c_void_p, P, C.c_double, C.c_double, _u8p, _u8p, _dp,
                                        C.POINTER(C.c_size_t)]
    L.ppsfm_filter_observations_with_negative_depth.argtypes = [C.c_void_p, P, _u8p, _u8p,
                                                                C.POINTER(C.c_size_t)]
    L._filter_declared = True


def FilterPoints3D(ctx, problem, max_reproj_error, min_tri_angle, point_error=None):
    """Returns (num_filtered, obs_deleted [O], point_deleted [P], point_error [P])."""
    L = binding.load_library()
    _declare(L)
    O, P = len(problem.obs_image), len(problem.points)
    od, pd = np.zeros(max(O, 1), np.uint8), np.zeros(max(P, 1), np.uint8)
    pe = np.full(max(P, 1), -1.0) if point_error is None else np.ascontiguousarray(point_error, np.float64).copy()
    nf = C.c_size_t(0)
    ctx._check(L.ppsfm_filter_points3d(ctx._h, C.byref(problem.struct), max_reproj_error,
                                       min_tri_angle, od.ctypes.data_as(_u8p),
                                       pd.ctypes.data_as(_u8p), pe.ctypes.data_as(_dp), C.byref(nf)))
    return nf.value, od[:O], pd[:P], pe[:P]


def FilterObservationsWithNegativeDepth(ctx, problem):
    """Returns (num_filtered, obs_deleted [O], point_deleted [P]): DeleteObservation removes the
    whole point once its track is down to three elements (reconstruction.cc:255-275)."""
    L = binding.load_library()
    _declare(L)
    O, P = len(problem.obs_image), len(problem.points)
    od, pd = np.zeros(max(O, 1), np.uint8), np.zeros(max(P, 1), np.uint8)
    nf = C.c_size_t(0)
    ctx._check(L.ppsfm_filter_observations_with_negative_depth(
        ctx._h, C.byref(problem.struct), od.ctypes.data_as(_u8p), pd.ctypes.data_as(_u8p),
        C.byref(nf)))
    return nf.value, od[:O], pd[:P]
