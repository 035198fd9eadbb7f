"""Line lifting (the producer side of ``FeatureLine``): host-side mirror of
what the reference's feature extraction does to every keypoint before anything of the path sees
it (src/feature/extraction.cc:440-504) — the point is taken to normalised camera coordinates
(``Camera::ImageToWorld``, src/base/camera_models.h: pinhole models :629-688, Newton
undistortion ``IterativeUndistortion`` :547-588 over the model's ``Distortion``, the closed-form
``FOVCameraModel::Undistortion`` :1176-1210, the thin-prism fisheye rescaling :1437-1457),
replaced by a line through it — ``gravity x (u, v, 1)`` for the features drawn to be
gravity-aligned, ``random_direction x (u, v, 1)`` for the others — and the line is scaled so that
||(a, b)|| = 1 (:497-501).  The result is the ``lines [n, 3]`` + ``aligned [n]`` layout of the
C-ABI (include/ppsfm_b200.h) and of images.txt / the database blob (float (a, b, c, aligned),
src/base/database.cc:55-73: ``feature_lines_to_blob`` / ``feature_lines_from_blob``).

Data preparation, vectorised numpy over the keypoints of one image; nothing here is on the
measured path.  ``ImageToWorld`` is checked against the reference's own camera models compiled
from /root/reference (oracle/_ref/libref_cost.so, tests/test_ref_lifting.py): bit-identical for
the polynomial models (every point runs its own Newton iteration and stops on its own step, as
the scalar code does), to 1e-14 where ``atan`` / ``tan`` / ``sin`` / ``cos`` of numpy and of the C
library may differ in the last bit (fisheye, FOV).  The reference draws its random directions
from ``Eigen::Vector3d::Random()`` and the aligned subset from ``RandomInteger``; here both are
arguments (``select_aligned_features`` follows the reference's rule for the subset).
"""
import numpy as np

_EPS = float(np.finfo(np.float64).eps)


def _distortion(model, extra, u, v):
    """CameraModel::Distortion of the models that undistort iteratively; operation order as in
    src/base/camera_models.h (:746-757, :815-827, :887-902, :962-990, :1057-1081, :1271-1290,
    :1347-1370, :1459-1481)."""
    if model in (2, 3):                                   # SIMPLE_RADIAL, RADIAL
        u2, v2 = u * u, v * v
        r2 = u2 + v2
        radial = extra[0] * r2 if model == 2 else extra[0] * r2 + extra[1] * r2 * r2
        return u * radial, v * radial
    if model == 4:                                        # OPENCV
        k1, k2, p1, p2 = extra[:4]
        u2, uv, v2 = u * u, u * v, v * v
        r2 = u2 + v2
        radial = k1 * r2 + k2 * r2 * r2
        return (u * radial + 2.0 * p1 * uv + p2 * (r2 + 2.0 * u2),
                v * radial + 2.0 * p2 * uv + p1 * (r2 + 2.0 * v2))
    if model == 6:                                        # FULL_OPENCV
        k1, k2, p1, p2, k3, k4, k5, k6 = extra[:8]
        u2, uv, v2 = u * u, u * v, v * v
        r2 = u2 + v2
        r4 = r2 * r2
        r6 = r4 * r2
        radial = (1.0 + k1 * r2 + k2 * r4 + k3 * r6) / (1.0 + k4 * r2 + k5 * r4 + k6 * r6)
        return (u * radial + 2.0 * p1 * uv + p2 * (r2 + 2.0 * u2) - u,
                v * radial + 2.0 * p2 * uv + p1 * (r2 + 2.0 * v2) - v)
    if model == 10:                                       # THIN_PRISM_FISHEYE
        k1, k2, p1, p2, k3, k4, sx1, sy1 = extra[:8]
        u2, uv, v2 = u * u, u * v, v * v
        r2 = u2 + v2
        r4 = r2 * r2
        r6 = r4 * r2
        r8 = r6 * r2
        radial = k1 * r2 + k2 * r4 + k3 * r6 + k4 * r8
        return (u * radial + 2.0 * p1 * uv + p2 * (r2 + 2.0 * u2) + sx1 * r2,
                v * radial + 2.0 * p2 * uv + p1 * (r2 + 2.0 * v2) + sy1 * r2)
    if model in (5, 8, 9):                                # OPENCV_FISHEYE, (SIMPLE_)RADIAL_FISHEYE
        r = np.sqrt(u * u + v * v)
        big = r > _EPS
        rs = np.where(big, r, 1.0)
        theta = np.arctan(rs)
        theta2 = theta * theta
        if model == 8:
            thetad = theta * (1.0 + extra[0] * theta2)
        elif model == 9:
            theta4 = theta2 * theta2
            thetad = theta * (1.0 + extra[0] * theta2 + extra[1] * theta4)
        else:
            theta4 = theta2 * theta2
            theta6 = theta4 * theta2
            theta8 = theta4 * theta4
            thetad = theta * (1.0 + extra[0] * theta2 + extra[1] * theta4 + extra[2] * theta6
                              + extra[3] * theta8)
        return np.where(big, u * thetad / rs - u, 0.0), np.where(big, v * thetad / rs - v, 0.0)
    raise ValueError("camera model %d has no iterative undistortion" % model)


def _iterative_undistortion(model, extra, u0, v0):
    """BaseCameraModel::IterativeUndistortion (:547-588): Newton on x + Distortion(x) = x0 with a
    central-difference Jacobian, at most 100 iterations, every point stopping on its own step."""
    x0, x1 = u0.copy(), v0.copy()
    active = np.ones(len(u0), bool)
    for _ in range(100):
        idx = np.flatnonzero(active)
        if len(idx) == 0:
            break
        a, b = x0[idx], x1[idx]
        step0 = np.maximum(_EPS, np.abs(1e-6 * a))
        step1 = np.maximum(_EPS, np.abs(1e-6 * b))
        dx0, dx1 = _distortion(model, extra, a, b)
        b0 = _distortion(model, extra, a - step0, b)
        f0 = _distortion(model, extra, a + step0, b)
        b1 = _distortion(model, extra, a, b - step1)
        f1 = _distortion(model, extra, a, b + step1)
        j00 = 1 + (f0[0] - b0[0]) / (2 * step0)
        j01 = (f1[0] - b1[0]) / (2 * step1)
        j10 = (f0[1] - b0[1]) / (2 * step0)
        j11 = 1 + (f1[1] - b1[1]) / (2 * step1)
        # J.inverse() * (x + dx - x0): adjugate times 1 / det, then the 2 x 2 product
        invdet = 1.0 / (j00 * j11 - j10 * j01)
        i00, i10, i01, i11 = j11 * invdet, -j10 * invdet, -j01 * invdet, j00 * invdet
        e0, e1 = a + dx0 - u0[idx], b + dx1 - v0[idx]
        s0, s1 = i00 * e0 + i01 * e1, i10 * e0 + i11 * e1
        x0[idx], x1[idx] = a - s0, b - s1
        active[idx[s0 * s0 + s1 * s1 < 1e-10]] = False
    return x0, x1


def ImageToWorld(camera_model, params, xy):
    """CameraModelImageToWorld for all 11 models: pixels [n, 2] -> normalised camera coordinates."""
    p = np.asarray(params, np.float64)
    xy = np.asarray(xy, np.float64).reshape(-1, 2)
    x, y = xy[:, 0], xy[:, 1]
    if camera_model in (0, 2, 3, 8, 9):                   # f, cx, cy, ...
        u, v, extra = (x - p[1]) / p[0], (y - p[2]) / p[0], p[3:]
    elif camera_model in (1, 4, 5, 6, 7, 10):             # fx, fy, cx, cy, ...
        u, v, extra = (x - p[2]) / p[0], (y - p[3]) / p[1], p[4:]
    else:
        raise ValueError("unknown camera model %d" % camera_model)
    if camera_model in (0, 1):
        return np.stack([u, v], 1)
    if camera_model == 7:                                 # FOVCameraModel::Undistortion
        omega = extra[0]
        radius2 = u * u + v * v
        omega2 = omega * omega
        if omega2 < 1e-4:
            factor = (omega2 * radius2) / 3.0 - omega2 / 12.0 + 1.0
        else:
            small = radius2 < 1e-4
            radius = np.sqrt(np.where(small, 1.0, radius2))
            factor = np.where(small,
                              (omega * (omega * omega * radius2 + 3.0)) / (6.0 * np.tan(omega / 2.0)),
                              np.tan(radius * omega) / (radius * 2.0 * np.tan(omega / 2.0)))
        return np.stack([u * factor, v * factor], 1)
    u, v = _iterative_undistortion(camera_model, extra, u, v)
    if camera_model == 10:                                # theta -> tan(theta) (:1450-1456)
        theta = np.sqrt(u * u + v * v)
        theta_cos_theta = theta * np.cos(theta)
        big = theta_cos_theta > _EPS
        scale = np.where(big, np.sin(theta) / np.where(big, theta_cos_theta, 1.0), 1.0)
        u, v = u * scale, v * scale
    return np.stack([u, v], 1)


def WorldToImage(camera_model, params, uv):
    """CameraModelWorldToImage for all 11 models: normalised camera coordinates [n, 2] -> pixels
    (the forward model of the filters' pixel-space reprojection error, src/base/projection.cc)."""
    p = np.asarray(params, np.float64)
    uv = np.asarray(uv, np.float64).reshape(-1, 2)
    u, v = uv[:, 0], uv[:, 1]
    single = camera_model in (0, 2, 3, 8, 9)
    if not single and camera_model not in (1, 4, 5, 6, 7, 10):
        raise ValueError("unknown camera model %d" % camera_model)
    f1, f2, c1, c2, extra = (p[0], p[0], p[1], p[2], p[3:]) if single else (p[0], p[1], p[2], p[3], p[4:])
    if camera_model in (0, 1):
        return np.stack([f1 * u + c1, f2 * v + c2], 1)
    if camera_model == 7:                                 # FOVCameraModel::Distortion (:1136-1171)
        omega = extra[0]
        radius2 = u * u + v * v
        omega2 = omega * omega
        if omega2 < 1e-4:
            factor = (omega2 * radius2) / 3.0 - omega2 / 12.0 + 1.0
        else:
            tan_half = np.tan(omega / 2.0)
            small = radius2 < 1e-4
            radius = np.sqrt(np.where(small, 1.0, radius2))
            factor = np.where(small,
                              (-2.0 * tan_half * (4.0 * radius2 * tan_half * tan_half - 3.0)) / (3.0 * omega),
                              np.arctan(radius * 2.0 * tan_half) / (radius * omega))
        return np.stack([f1 * (u * factor) + c1, f2 * (v * factor) + c2], 1)
    if camera_model == 10:                                # equidistant pre-mapping (:1412-1421)
        r = np.sqrt(u * u + v * v)
        big = r > _EPS
        rs = np.where(big, r, 1.0)
        theta = np.arctan(rs)
        u, v = np.where(big, theta * u / rs, u), np.where(big, theta * v / rs, v)
    du, dv = _distortion(camera_model, extra, u, v)
    return np.stack([f1 * (u + du) + c1, f2 * (v + dv) + c2], 1)


# focal-length / principal-point / extra parameter indices of the 11 models
# (src/base/camera_models.h: Initialize{FocalLength,PrincipalPoint,ExtraParams}Idxs)
_PARAM_GROUPS = {0: ((0,), (1, 2), ()), 1: ((0, 1), (2, 3), ()), 2: ((0,), (1, 2), (3,)),
                 3: ((0,), (1, 2), (3, 4)), 4: ((0, 1), (2, 3), (4, 5, 6, 7)),
                 5: ((0, 1), (2, 3), (4, 5, 6, 7)), 6: ((0, 1), (2, 3), tuple(range(4, 12))),
                 7: ((0, 1), (2, 3), (4,)), 8: ((0,), (1, 2), (3,)), 9: ((0,), (1, 2), (3, 4)),
                 10: ((0, 1), (2, 3), tuple(range(4, 12)))}


def HasBogusParams(camera_model, params, width, height, min_focal_length_ratio,
                   max_focal_length_ratio, max_extra_param):
    """CameraModelHasBogusParams (src/base/camera_models.h:471-531): a focal length outside
    [min, max] x max(width, height), a principal point outside the image or an extra parameter
    above max_extra_param in magnitude — what the mapper tests before it trusts a camera
    (sfm/incremental_mapper.cc:684-700, 953-955; sfm/incremental_triangulator.cc:767-781)."""
    p = np.asarray(params, np.float64)
    focal, pp, extra = _PARAM_GROUPS[camera_model]
    max_size = max(int(width), int(height))
    for k in focal:
        ratio = p[k] / max_size
        if ratio < min_focal_length_ratio or ratio > max_focal_length_ratio:
            return True
    cx, cy = p[pp[0]], p[pp[1]]
    if cx < 0 or cx > width or cy < 0 or cy > height:
        return True
    return any(abs(p[k]) > max_extra_param for k in extra)


def select_aligned_features(num_features, aligned_line_ratio, rng):
    """The reference's rule (extraction.cc:452-457): draw feature indices uniformly WITH
    replacement until the set holds at least ratio * n of them.  ``rng``: numpy Generator."""
    chosen = set()
    while num_features > 0 and len(chosen) / float(num_features) < aligned_line_ratio:
        chosen.add(int(rng.integers(0, num_features)))
    mask = np.zeros(num_features, bool)
    mask[list(chosen)] = True
    return mask


def LiftLines(camera_model, params, keypoints, directions, aligned=None, gravity=None):
    """keypoints [n, 2] (pixels) -> (lines [n, 3], aligned [n] uint8).  ``directions`` [n, 3]: the
    random direction of every feature (the reference: ``Eigen::Vector3d::Random()``, uniform in
    [-1, 1]^3); ``aligned`` [n] bool with ``gravity`` [3]: these features take
    ``gravity x point`` instead (without gravity nothing is aligned, :470-474, :484)."""
    uv = ImageToWorld(camera_model, params, keypoints)
    n = len(uv)
    point = np.concatenate([uv, np.ones((n, 1))], 1)
    d = np.array(directions, np.float64).reshape(n, 3)
    flags = np.zeros(n, bool)
    if gravity is not None and aligned is not None and not np.isnan(np.asarray(gravity, np.float64)).any():
        flags = np.asarray(aligned, bool).reshape(n).copy()
        d[flags] = np.asarray(gravity, np.float64)
    # Eigen's cross(): (a1 b2 - a2 b1, a2 b0 - a0 b2, a0 b1 - a1 b0)
    line = np.stack([d[:, 1] * point[:, 2] - d[:, 2] * point[:, 1],
                     d[:, 2] * point[:, 0] - d[:, 0] * point[:, 2],
                     d[:, 0] * point[:, 1] - d[:, 1] * point[:, 0]], 1)
    head = np.sqrt(line[:, 0] * line[:, 0] + line[:, 1] * line[:, 1])
    return line / head[:, None], flags.astype(np.uint8)


def feature_lines_to_blob(lines, aligned):
    """FeatureLinesToBlob (src/base/database.cc:55-62): row-major float32 [n, 4] = (a, b, c, aligned)."""
    blob = np.empty((len(lines), 4), np.float32)
    blob[:, :3] = np.asarray(lines, np.float64).reshape(-1, 3).astype(np.float32)
    blob[:, 3] = np.where(np.asarray(aligned).reshape(-1) != 0, 1.0, 0.0)
    return blob.tobytes()


def feature_lines_from_blob(data, rows=None):
    """FeatureLinesFromBlob (:64-74): (lines [n, 3] double, renormalised so that ||(a, b)|| = 1 —
    the blob holds floats —, aligned [n] uint8 = column 3 > 0)."""
    blob = np.frombuffer(data, np.float32).reshape(-1, 4)
    if rows is not None and rows != len(blob):
        raise ValueError("line_features blob: %d rows expected, %d found" % (rows, len(blob)))
    line = blob[:, :3].astype(np.float64)
    norm = np.sqrt(line[:, 0] * line[:, 0] + line[:, 1] * line[:, 1])
    return line / norm[:, None], (blob[:, 3] > 0).astype(np.uint8)
