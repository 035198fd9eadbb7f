"""Host-side mirror of the reference's bundle-adjustment API (src/optim/bundle_adjustment.{h,cc})
on top of libppsfm_b200.so.

  * ``BundleAdjustmentOptions``  — src/optim/bundle_adjustment.h:49-100 (+ the ceres solver options
    the reference sets in src/controllers/incremental_mapper.cc:196-243)
  * ``BundleAdjustmentConfig``   — src/optim/bundle_adjustment.h:103-167
  * ``BundleAdjuster``           — src/optim/bundle_adjustment.h:171-215; ``Solve`` flattens a
    ``Reconstruction`` exactly as SetUp/AddImageToProblem/AddPointToProblem/Parameterize* do and
    hands the arrays to ``ppsfm_ba_solve``; results are written back in place.
  * ``RefineAbsolutePoseFromLines`` — src/estimators/pose.cc:96-213
The ``Reconstruction`` / ``Image`` / ``Camera`` / ``Point3D`` classes are minimal in-memory
stand-ins for src/base (only the members BundleAdjuster touches).  All numerics run on the GPU.
"""
import ctypes as C
from enum import IntEnum

import numpy as np

from . import binding
from .binding import PpsfmError, _dp, _u8p, _i32p

BA_MAX_TRACE = 128
# src/base/camera_models.h:189-248
CAMERA_MODEL_IDS = {"SIMPLE_PINHOLE": 0, "PINHOLE": 1, "SIMPLE_RADIAL": 2, "RADIAL": 3, "OPENCV": 4,
                    "OPENCV_FISHEYE": 5, "FULL_OPENCV": 6, "FOV": 7, "SIMPLE_RADIAL_FISHEYE": 8,
                    "RADIAL_FISHEYE": 9, "THIN_PRISM_FISHEYE": 10}
CAMERA_NUM_PARAMS = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8, 5: 8, 6: 12, 7: 5, 8: 4, 9: 5, 10: 12}


class LossFunctionType(IntEnum):
    TRIVIAL = 0
    SOFT_L1 = 1
    CAUCHY = 2


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
                ("trace_accepted", C.c_int32 * BA_MAX_TRACE),
                ("jacobian_launches", C.c_int32), ("kernel_launches", C.c_int64),
                ("schur_time_s", C.c_double), ("cholesky_time_s", C.c_double),
                ("backsub_time_s", C.c_double)]

    def IsSolutionUsable(self):
        return self.termination_type != 2

    @property
    def num_iterations(self):
        return self.num_successful_steps + self.num_unsuccessful_steps


_ready = False


def _lib():
    global _ready
    L = binding.load_library()
    if not _ready:
        vp = C.c_void_p
        L.ppsfm_ba_options_default.argtypes = [C.POINTER(BaOptions)]
        L.ppsfm_ba_options_default.restype = None
        L.ppsfm_ba_solve.argtypes = [vp, C.POINTER(BaProblem), C.POINTER(BaOptions),
                                     C.POINTER(BaSummary)]
        L.ppsfm_ba_create.argtypes = [vp, C.POINTER(BaProblem), C.POINTER(BaOptions),
                                      C.POINTER(vp)]
        L.ppsfm_ba_run.argtypes = [vp, C.POINTER(BaSummary)]
        L.ppsfm_ba_reset.argtypes = [vp]
        L.ppsfm_ba_download.argtypes = [vp, C.POINTER(BaProblem)]
        L.ppsfm_ba_free.argtypes = [vp]
        L.ppsfm_ba_free.restype = None
        L.ppsfm_refine_absolute_pose_from_lines.argtypes = [
            vp, _u8p, _dp, _dp, C.c_size_t, C.c_int, _dp, C.c_double, C.c_int, C.c_double, _dp,
            _dp, C.POINTER(BaSummary)]
        L.ppsfm_refine_absolute_pose_from_lines_ex.argtypes = [
            vp, _u8p, _dp, _dp, C.c_size_t, C.c_int, _dp, C.c_int, C.c_int, C.c_double, C.c_int,
            C.c_double, _dp, _dp, C.POINTER(BaSummary)]
        L.ppsfm_ba_linearize.argtypes = [vp, C.POINTER(BaProblem), C.POINTER(BaOptions), _dp, _dp,
                                         _dp, _dp]
        L.ppsfm_dense_cholesky_solve.argtypes = [vp, _dp, C.c_int, _dp, _dp]
        L.ppsfm_ba_shard_stats.argtypes = [C.POINTER(BaProblem), C.c_int, C.c_int,
                                           C.POINTER(C.c_int64)]
        _ready = True
    return L


def default_solver_options(**kw):
    o = BaOptions()
    _lib().ppsfm_ba_options_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class BaArrays:
    """Contiguous numpy buffers of a flattened BA problem + the ctypes struct over them."""

    def __init__(self, qvecs, tvecs, points, obs_image, obs_point, obs_line, camera_model,
                 camera_params, image_camera=None, pose_flags=None, point_const=None, copy=True,
                 camera_const=None):
        f64 = dict(dtype=np.float64)
        self.qvecs = np.array(qvecs, **f64) if copy else np.ascontiguousarray(qvecs, **f64)
        self.tvecs = np.array(tvecs, **f64) if copy else np.ascontiguousarray(tvecs, **f64)
        self.points = np.array(points, **f64) if copy else np.ascontiguousarray(points, **f64)
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
        # config.IsConstantCamera per camera (only read when an options.refine_* flag is set)
        self.camera_const = np.ascontiguousarray(
            camera_const if camera_const is not None else np.zeros(self.camera_model.shape[0]),
            dtype=np.uint8)
        p = BaProblem()
        p.camera_const = self.camera_const.ctypes.data_as(_u8p)
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
        self.struct = p


def solve_arrays(ctx, arrays, solver_options):
    """ppsfm_ba_solve on flat arrays (updated in place). Returns (ok, summary)."""
    s = BaSummary()
    rc = ctx._check(_lib().ppsfm_ba_solve(ctx._h, C.byref(arrays.struct),
                                          C.byref(solver_options), C.byref(s)),
                    allow_no_solution=True)
    return rc == binding.PPSFM_OK, s


class ResidentProblem:
    """A BA problem resident in HBM (ppsfm_ba): create once, run / reset / download."""

    def __init__(self, ctx, arrays, solver_options):
        self._ctx, self._arrays = ctx, arrays
        h = C.c_void_p()
        ctx._check(_lib().ppsfm_ba_create(ctx._h, C.byref(arrays.struct),
                                          C.byref(solver_options), C.byref(h)))
        self._h = h

    def run(self):
        s = BaSummary()
        rc = self._ctx._check(_lib().ppsfm_ba_run(self._h, C.byref(s)), allow_no_solution=True)
        return rc == binding.PPSFM_OK, s

    def reset(self):
        self._ctx._check(_lib().ppsfm_ba_reset(self._h))

    def download(self):
        self._ctx._check(_lib().ppsfm_ba_download(self._h, C.byref(self._arrays.struct)))

    def free(self):
        if self._h is not None:
            _lib().ppsfm_ba_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def linearize_arrays(ctx, arrays, solver_options):
    """Test hook: per-observation residual[2], jac_cam[2,6], jac_point[2,3] and the cost."""
    n = arrays.obs_image.shape[0]
    r = np.zeros((n, 2))
    jc = np.zeros((n, 2, 6))
    jp = np.zeros((n, 2, 3))
    cost = C.c_double()
    ctx._check(_lib().ppsfm_ba_linearize(ctx._h, C.byref(arrays.struct), C.byref(solver_options),
                                         r.ctypes.data_as(_dp), jc.ctypes.data_as(_dp),
                                         jp.ctypes.data_as(_dp), C.byref(cost)))
    return r, jc, jp, float(cost.value)


def shard_stats(arrays, rank, world):
    """Host-only: (kept observations on `rank`, owned points, camera blocks, kept obs in total)."""
    out = (C.c_int64 * 4)()
    rc = _lib().ppsfm_ba_shard_stats(C.byref(arrays.struct), rank, world, out)
    if rc != 0:
        raise PpsfmError(f"ppsfm_ba_shard_stats rc={rc}")
    return tuple(int(v) for v in out)


def dense_cholesky_solve(ctx, A, b):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    rc = ctx._check(_lib().ppsfm_dense_cholesky_solve(ctx._h, A.ctypes.data_as(_dp), A.shape[0],
                                                      b.ctypes.data_as(_dp),
                                                      x.ctypes.data_as(_dp)),
                    allow_no_solution=True)
    return rc == binding.PPSFM_OK, x


# ---------------------------------------------------------------------------------------------
# Minimal data-model stand-ins (src/base/{camera,image,point3d,track,reconstruction}.h)
# ---------------------------------------------------------------------------------------------
kInvalidPoint3DId = -1


class Camera:
    def __init__(self, camera_id, model, params):
        self.camera_id = camera_id
        self.model_id = CAMERA_MODEL_IDS[model] if isinstance(model, str) else int(model)
        self.params = np.array(params, dtype=np.float64)

    def ModelId(self):
        return self.model_id

    def Params(self):
        return self.params


class FeatureLine:
    """src/feature/types.h:98-138"""

    def __init__(self, line, is_aligned=False, point3D_id=kInvalidPoint3DId):
        self.line = np.array(line, dtype=np.float64)
        self.is_aligned = bool(is_aligned)
        self.point3D_id = point3D_id

    def HasPoint3D(self):
        return self.point3D_id != kInvalidPoint3DId


class Image:
    def __init__(self, image_id, camera_id, qvec, tvec, lines=()):
        self.image_id, self.camera_id = image_id, camera_id
        self.qvec = np.array(qvec, dtype=np.float64)
        self.tvec = np.array(tvec, dtype=np.float64)
        self.lines = list(lines)

    def NormalizeQvec(self):
        n = np.linalg.norm(self.qvec)
        if n > 0:
            self.qvec /= n


class Point3D:
    def __init__(self, xyz, track=()):
        self.xyz = np.array(xyz, dtype=np.float64)
        self.track = list(track)  # [(image_id, line_idx)]


class Reconstruction:
    def __init__(self):
        self.cameras, self.images, self.points3D = {}, {}, {}

    def Camera(self, camera_id):
        return self.cameras[camera_id]

    def Image(self, image_id):
        return self.images[image_id]

    def Point3D(self, point3D_id):
        return self.points3D[point3D_id]


class BundleAdjustmentOptions:
    """src/optim/bundle_adjustment.h:49-100"""
    LossFunctionType = LossFunctionType

    def __init__(self):
        self.loss_function_type = LossFunctionType.TRIVIAL
        self.loss_function_scale = 1.0
        self.refine_focal_length = False
        self.refine_principal_point = False
        self.refine_extra_params = False
        self.refine_extrinsics = True
        self.print_summary = True
        self.min_num_residuals_for_multi_threading = 50000
        self.solver_options = default_solver_options()

    def Check(self):
        if self.loss_function_scale < 0:  # CHECK_OPTION_GE(loss_function_scale, 0)
            print("CHECK_OPTION_GE(loss_function_scale, 0) failed")
            return False
        return True


class BundleAdjustmentConfig:
    """src/optim/bundle_adjustment.h:103-167, .cc:81-248"""

    def __init__(self):
        self._constant_camera_ids = set()
        self._image_ids = set()
        self._variable_point3D_ids = set()
        self._constant_point3D_ids = set()
        self._constant_poses = set()
        self._constant_tvecs = {}

    def NumImages(self): return len(self._image_ids)
    def NumPoints(self): return len(self._variable_point3D_ids) + len(self._constant_point3D_ids)
    def NumConstantCameras(self): return len(self._constant_camera_ids)
    def NumConstantPoses(self): return len(self._constant_poses)
    def NumConstantTvecs(self): return len(self._constant_tvecs)
    def NumVariablePoints(self): return len(self._variable_point3D_ids)
    def NumConstantPoints(self): return len(self._constant_point3D_ids)

    def AddImage(self, image_id): self._image_ids.add(image_id)
    def HasImage(self, image_id): return image_id in self._image_ids

    def RemoveImage(self, image_id): self._image_ids.discard(image_id)

    def SetConstantCamera(self, camera_id): self._constant_camera_ids.add(camera_id)
    def SetVariableCamera(self, camera_id): self._constant_camera_ids.discard(camera_id)
    def IsConstantCamera(self, camera_id): return camera_id in self._constant_camera_ids

    def SetConstantPose(self, image_id):
        if not self.HasImage(image_id):
            raise PpsfmError("CHECK(HasImage(image_id))")
        if self.HasConstantTvec(image_id):
            raise PpsfmError("CHECK(!HasConstantTvec(image_id))")
        self._constant_poses.add(image_id)

    def SetVariablePose(self, image_id): self._constant_poses.discard(image_id)
    def HasConstantPose(self, image_id): return image_id in self._constant_poses

    def SetConstantTvec(self, image_id, idxs):
        idxs = list(idxs)
        if not idxs or len(idxs) > 3 or len(set(idxs)) != len(idxs):
            raise PpsfmError("CHECK: tvec indices must be unique and in [0, 2]")
        if not self.HasImage(image_id) or self.HasConstantPose(image_id):
            raise PpsfmError("CHECK(HasImage && !HasConstantPose)")
        self._constant_tvecs[image_id] = idxs

    def RemoveConstantTvec(self, image_id): self._constant_tvecs.pop(image_id, None)
    def HasConstantTvec(self, image_id): return image_id in self._constant_tvecs
    def ConstantTvec(self, image_id): return self._constant_tvecs[image_id]

    def AddVariablePoint(self, pid):
        if self.HasConstantPoint(pid):
            raise PpsfmError("CHECK(!HasConstantPoint(point3D_id))")
        self._variable_point3D_ids.add(pid)

    def AddConstantPoint(self, pid):
        if self.HasVariablePoint(pid):
            raise PpsfmError("CHECK(!HasVariablePoint(point3D_id))")
        self._constant_point3D_ids.add(pid)

    def HasPoint(self, pid): return self.HasVariablePoint(pid) or self.HasConstantPoint(pid)
    def HasVariablePoint(self, pid): return pid in self._variable_point3D_ids
    def HasConstantPoint(self, pid): return pid in self._constant_point3D_ids
    def RemoveVariablePoint(self, pid): self._variable_point3D_ids.discard(pid)
    def RemoveConstantPoint(self, pid): self._constant_point3D_ids.discard(pid)
    def Images(self): return self._image_ids
    def VariablePoints(self): return self._variable_point3D_ids
    def ConstantPoints(self): return self._constant_point3D_ids


class BundleAdjuster:
    """src/optim/bundle_adjustment.h:171-215.  Single use, like the reference (:262)."""

    def __init__(self, options, config, ctx=None):
        if not options.Check():
            raise PpsfmError("CHECK(options_.Check())")
        self._options, self._config = options, config
        self._ctx = ctx or binding.default_context()
        self._summary = None
        self._used = False

    def Summary(self):
        return self._summary

    def Solve(self, reconstruction):
        if self._used:
            raise PpsfmError("Cannot use the same BundleAdjuster multiple times")
        self._used = True
        arrays, img_ids, pt_ids, cam_ids = self._SetUp(reconstruction)
        if arrays.obs_image.shape[0] == 0:  # problem_->NumResiduals() == 0 -> false
            return False
        o = self._options.solver_options
        o.loss_type = int(self._options.loss_function_type)
        o.loss_scale = self._options.loss_function_scale
        # ParameterizeCameras (:490-528)
        o.refine_focal_length = int(bool(self._options.refine_focal_length))
        o.refine_principal_point = int(bool(self._options.refine_principal_point))
        o.refine_extra_params = int(bool(self._options.refine_extra_params))
        ok, self._summary = solve_arrays(self._ctx, arrays, o)
        if o.refine_focal_length or o.refine_principal_point or o.refine_extra_params:
            for i, cid in enumerate(cam_ids):  # camera.ParamsData() updated in place
                cam = reconstruction.Camera(cid)
                cam.params[:] = arrays.camera_params[i, :len(cam.params)]
        # results in place (image.Qvec().data(), image.Tvec().data(), point3D.XYZ().data())
        for i, iid in enumerate(img_ids):
            img = reconstruction.Image(iid)
            img.qvec[:] = arrays.qvecs[i]
            img.tvec[:] = arrays.tvecs[i]
        for i, pid in enumerate(pt_ids):
            reconstruction.Point3D(pid).xyz[:] = arrays.points[i]
        if self._options.print_summary:
            s = self._summary
            print(f"Bundle adjustment report: residuals {s.num_residuals_reduced}, iterations "
                  f"{s.num_iterations}, cost {s.initial_cost:.6g} -> {s.final_cost:.6g}")
        return True

    def _SetUp(self, rec):
        """bundle_adjustment.cc:326-542 flattened to arrays."""
        cfg, opt = self._config, self._options
        img_index, pt_index, cam_index = {}, {}, {}
        qv, tv, flags, icam, cmodel, cparams, pts = [], [], [], [], [], [], []
        oi, op, ol = [], [], []
        num_obs_of_point = {}
        config_cameras, outside_cameras = set(), set()

        def cam_idx(camera_id):
            if camera_id not in cam_index:
                cam = rec.Camera(camera_id)
                if cam.ModelId() not in CAMERA_NUM_PARAMS:
                    raise PpsfmError(f"camera model {cam.ModelId()} not supported")
                cam_index[camera_id] = len(cmodel)
                cmodel.append(cam.ModelId())
                row = np.zeros(12)
                row[:len(cam.params)] = cam.params
                cparams.append(row)
            return cam_index[camera_id]

        def img_idx(image_id, constant):
            if image_id not in img_index:
                img = rec.Image(image_id)
                img_index[image_id] = len(qv)
                qv.append(img.qvec)
                tv.append(img.tvec)
                f = 1 if constant else 0
                if not constant and cfg.HasConstantTvec(image_id):
                    for k in cfg.ConstantTvec(image_id):
                        f |= 2 << k
                flags.append(f)
                icam.append(cam_idx(img.camera_id))
            return img_index[image_id]

        def pt_idx(pid):
            if pid not in pt_index:
                pt_index[pid] = len(pts)
                pts.append(rec.Point3D(pid).xyz)
            return pt_index[pid]

        # AddImageToProblem (:348-435)
        for image_id in sorted(cfg.Images()):
            img = rec.Image(image_id)
            img.NormalizeQvec()
            constant_pose = (not opt.refine_extrinsics) or cfg.HasConstantPose(image_id)
            for line in img.lines:
                if not line.HasPoint3D():
                    continue
                if abs(np.linalg.norm(line.line[:2]) - 1.0) > 1e-6:
                    raise PpsfmError("CHECK_NEAR(line.Line().head<2>().norm(), 1.0, 1e-6)")
                num_obs_of_point[line.point3D_id] = num_obs_of_point.get(line.point3D_id, 0) + 1
                oi.append(img_idx(image_id, constant_pose))
                op.append(pt_idx(line.point3D_id))
                ol.append(line.line)
                config_cameras.add(img.camera_id)         # camera_ids_.insert (:432-434)
        # AddPointToProblem (:437-488): observations of configured points in images outside the
        # image set enter through the constant-pose functor
        for pid in list(sorted(cfg.VariablePoints())) + list(sorted(cfg.ConstantPoints())):
            p3 = rec.Point3D(pid)
            if num_obs_of_point.get(pid, 0) == len(p3.track):
                continue
            for image_id, line_idx in p3.track:
                if cfg.HasImage(image_id):
                    continue
                num_obs_of_point[pid] = num_obs_of_point.get(pid, 0) + 1
                img = rec.Image(image_id)
                # a camera that enters only through images outside the configuration is
                # constant (config_.SetConstantCamera, :476-479)
                if img.camera_id not in config_cameras:
                    outside_cameras.add(img.camera_id)
                oi.append(img_idx(image_id, True))
                op.append(pt_idx(pid))
                ol.append(img.lines[line_idx].line)
        # ParameterizePoints (:530-542)
        point_const = np.zeros(len(pts), dtype=np.uint8)
        for pid, n_obs in num_obs_of_point.items():
            if len(rec.Point3D(pid).track) > n_obs:
                point_const[pt_index[pid]] = 1
        for pid in cfg.ConstantPoints():
            if pid in pt_index:
                point_const[pt_index[pid]] = 1
        cam_ids = [None] * len(cmodel)
        for cid, i in cam_index.items():
            cam_ids[i] = cid
        z = lambda rows, w: np.array(rows, dtype=np.float64).reshape(-1, w)
        arrays = BaArrays(z(qv, 4), z(tv, 3), z(pts, 3), oi, op, z(ol, 3),
                          cmodel if cmodel else [1], z(cparams, 12) if cparams else np.zeros((1, 12)),
                          image_camera=icam, pose_flags=flags, point_const=point_const,
                          camera_const=[1 if (cfg.IsConstantCamera(c) or c in outside_cameras)
                                        else 0 for c in cam_ids]
                          if cam_ids else None)
        img_ids = [None] * len(qv)
        for iid, i in img_index.items():
            img_ids[i] = iid
        pt_ids = [None] * len(pts)
        for pid, i in pt_index.items():
            pt_ids[i] = pid
        return arrays, img_ids, pt_ids, cam_ids


class AbsolutePoseRefinementOptions:
    """src/estimators/pose.h:84-108"""

    def __init__(self):
        self.gradient_tolerance = 1.0
        self.max_num_iterations = 100
        self.loss_function_scale = 1.0
        self.refine_focal_length = False
        self.refine_extra_params = False
        self.print_summary = True

    def Check(self):
        if self.gradient_tolerance < 0 or self.max_num_iterations < 0 or \
                self.loss_function_scale < 0:
            raise PpsfmError("AbsolutePoseRefinementOptions::Check")


def RefineAbsolutePoseFromLines(options, inlier_mask, lines2D, points3D, qvec, tvec, camera,
                                ctx=None):
    """src/estimators/pose.cc:96-213.  qvec / tvec are numpy arrays updated in place; returns the
    reference's bool (summary.IsSolutionUsable())."""
    ctx = ctx or binding.default_context()
    options.Check()
    lines2D = np.ascontiguousarray(lines2D, dtype=np.float64)
    points3D = np.ascontiguousarray(points3D, dtype=np.float64)
    inlier_mask = np.ascontiguousarray(inlier_mask, dtype=np.uint8)
    if not (inlier_mask.shape[0] == lines2D.shape[0] == points3D.shape[0]):
        raise PpsfmError("CHECK_EQ(inlier_mask.size(), lines2D.size(), points3D.size())")
    q = np.ascontiguousarray(qvec, dtype=np.float64).copy()
    t = np.ascontiguousarray(tvec, dtype=np.float64).copy()
    params = np.zeros(12)
    params[:len(camera.params)] = camera.params
    s = BaSummary()
    refine = bool(options.refine_focal_length or options.refine_extra_params)
    rc = ctx._check(_lib().ppsfm_refine_absolute_pose_from_lines_ex(
        ctx._h, inlier_mask.ctypes.data_as(_u8p), lines2D.ctypes.data_as(_dp),
        points3D.ctypes.data_as(_dp), lines2D.shape[0], camera.ModelId(),
        params.ctypes.data_as(_dp), int(bool(options.refine_focal_length)),
        int(bool(options.refine_extra_params)), options.gradient_tolerance,
        options.max_num_iterations, options.loss_function_scale, q.ctypes.data_as(_dp),
        t.ctypes.data_as(_dp), C.byref(s)),
        allow_no_solution=True)
    qvec[:] = q
    tvec[:] = t
    if refine:  # camera->ParamsData() is a parameter block of the problem (pose.cc:108, 138)
        camera.params[:] = params[:len(camera.params)]
    RefineAbsolutePoseFromLines.last_summary = s
    return rc == binding.PPSFM_OK

