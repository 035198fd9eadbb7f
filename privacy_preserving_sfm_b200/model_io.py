"""Text model format of this fork: host-side mirror of
``Reconstruction::ReadText`` / ``WriteText`` (src/base/reconstruction.cc:543-553) — the data
format on either side of the path: what the mapper writes after bundle adjustment and what a
caller loads to run the operators of this library (filters, bundle adjustment) on a model.

  cameras.txt   CAMERA_ID MODEL WIDTH HEIGHT PARAMS[]                (ReadCamerasText :721-768,
                                                                      WriteCamerasText :963-991)
  images.txt    IMAGE_ID QW QX QY QZ TX TY TZ CAMERA_ID NAME
                LINES2D[] as (A, B, C, is_aligned, POINT3D_ID)        (ReadImagesText :770-886,
                                                                      WriteImagesText :993-1057)
  points3D.txt  POINT3D_ID X Y Z R G B ERROR TRACK[] as (IMAGE_ID, line_idx)
                                                                     (ReadPoints3DText :888-961,
                                                                      WritePoints3DText :1059-1095)

What the reference's reader does to the numbers, and this one with it (``reference_precision``):
  * QVEC / TVEC / XYZ / ERROR / camera parameters go through ``std::stold`` and are stored as
    double (``strtold`` of the C library here, then the same narrowing);
  * the quaternion is normalised (``Image::NormalizeQvec`` -> ``NormalizeQuaternion``,
    src/base/pose.cc:53-62: a zero quaternion becomes (1, x, y, z));
  * line coefficients go through ``std::stof`` — they are read as FLOAT —, widened to double and
    divided by the norm of (A, B) (:846-870); so a model written with 17 digits does not come back
    with the lines it was written with, but with their float roundings, renormalised;
  * is_aligned must be the token ``1`` or ``0`` (CHECK at :855); POINT3D_ID ``-1`` is "no point";
  * items are separated by single spaces (``std::getline(stream, item, ' ')``); the line after an
    image's header line is its LINES2D row whatever it starts with; all images read are
    registered (:795-796).
``reference_precision=False`` keeps the 17 digits of the lines (double parse, no renormalisation):
the lossless reader the mapper's own round trip uses.

The writer (``reference_precision=True``) produces the reference's records byte for byte.  The
reference sets ``precision(17)`` on the FILE stream only (:968, :998, :1064): what it writes to
the file directly — the header means, a point's XYZ and ERROR — has 17 significant digits
(``%.17g``), but camera parameters, QVEC / TVEC and the line coefficients go through a local
``std::ostringstream`` at the default precision (:976, :1014) and reach the file with SIX digits
(``%g``).  ``reference_precision=False`` writes 17 digits everywhere (what the mapper of this
package writes: a lossless round trip with the lossless reader).  The trailing blank of every row
is removed (:985-986, :1052-1054); header comments carry ``ComputeMeanObservationsPerRegImage`` /
``ComputeMeanTrackLength`` (:494-509); records are written by ascending id, where the reference
walks an unordered map (the readers on both sides do not depend on the order).

Pinned against the reference's own ``ReadText`` / ``WriteText`` compiled from
/root/reference (oracle/_ref/libref_filter.so): tests/test_ref_model_io.py.
"""
import ctypes as C
import ctypes.util
import os
import re

import numpy as np

from .filters import FilterProblem

# src/base/camera_models.h:117-130 (model ids and names), NumParams of each model
CAMERA_MODEL_NAMES = {0: "SIMPLE_PINHOLE", 1: "PINHOLE", 2: "SIMPLE_RADIAL", 3: "RADIAL", 4: "OPENCV",
                      5: "OPENCV_FISHEYE", 6: "FULL_OPENCV", 7: "FOV", 8: "SIMPLE_RADIAL_FISHEYE",
                      9: "RADIAL_FISHEYE", 10: "THIN_PRISM_FISHEYE"}
CAMERA_MODEL_IDS = {v: k for k, v in CAMERA_MODEL_NAMES.items()}
CAMERA_MODEL_NUM_PARAMS = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8, 5: 8, 6: 12, 7: 5, 8: 4, 9: 5, 10: 12}

_libc = C.CDLL(ctypes.util.find_library("c") or None, use_errno=True)
_libc.strtold.restype = C.c_longdouble
_libc.strtold.argtypes = [C.c_void_p, C.c_void_p]
_libc.strtof.restype = C.c_float
_libc.strtof.argtypes = [C.c_void_p, C.c_void_p]


def _c_parse(fn, item):
    """strtold / strtof with the std::sto* contract: leading white space and trailing characters
    are accepted, no conversion at all throws (std::invalid_argument), ERANGE throws
    (std::out_of_range)."""
    raw = item.encode()
    buf = C.create_string_buffer(raw)
    end = C.c_void_p()
    C.set_errno(0)
    value = fn(buf, C.byref(end))
    if end.value is None or end.value == C.addressof(buf):
        raise ValueError("sto*: no conversion of %r" % item)
    if C.get_errno() == 34:                                       # ERANGE
        raise ValueError("sto*: %r is out of range" % item)
    return value


def _stold(item):
    """double(std::stold(item)): parsed to long double, narrowed to double on assignment."""
    return float(_c_parse(_libc.strtold, item))


_INT = re.compile(r"\s*[+-]?\d+")


def _stoi(item):
    """std::stoul / stoll / stoi on well-formed ids: the leading integer, trailing characters ignored."""
    m = _INT.match(item)
    if not m:
        raise ValueError("sto*: no conversion of %r" % item)
    return int(m.group(0))


def _stof_row(items):
    """double(std::stof(item)) for a row of tokens.  Vectorised as double parse + narrowing; the
    tokens whose double value sits exactly between two floats (where rounding twice and rounding
    once can differ) and the ends of the float range go through strtof itself: std::stof throws
    std::out_of_range where strtof reports ERANGE (overflow, inexact subnormal results).  Rows
    with tokens numpy does not read (hexadecimal floats, trailing characters: strtof does) are
    parsed token by token."""
    try:
        d = np.array(items, dtype=np.float64)
    except ValueError:
        return np.array([float(_c_parse(_libc.strtof, x)) for x in items], np.float64)
    with np.errstate(over="ignore"):
        f = d.astype(np.float32)
    bits = d.view(np.uint64)
    redo = ((bits & np.uint64(0x1FFFFFFF)) == np.uint64(0x10000000)) | \
           ((np.abs(d) < 1.1754943508222875e-38) & (d != 0)) | (np.abs(d) > 3.4028234e38)
    for k in np.flatnonzero(redo):
        f[k] = _c_parse(_libc.strtof, items[k])
    return f.astype(np.float64)


class Camera:
    def __init__(self, model_id, width, height, params):
        self.model_id, self.width, self.height = int(model_id), int(width), int(height)
        self.params = np.array(params, np.float64).reshape(-1)

    @property
    def model_name(self):
        return CAMERA_MODEL_NAMES[self.model_id]


class Image:
    """lines [n, 3] (A, B, C), aligned [n] bool, point3D_ids [n] int64 (-1: the line has no point)."""

    def __init__(self, qvec, tvec, camera_id, name, lines, aligned, point3D_ids):
        self.qvec, self.tvec = np.array(qvec, np.float64), np.array(tvec, np.float64)
        self.camera_id, self.name = int(camera_id), str(name)
        self.lines = np.array(lines, np.float64).reshape(-1, 3)
        self.aligned = np.array(aligned, bool).reshape(-1)
        self.point3D_ids = np.array(point3D_ids, np.int64).reshape(-1)


class Point3D:
    """track [m, 2]: (IMAGE_ID, line_idx) — the index of the line among the image's lines."""

    def __init__(self, xyz, track, error=-1.0, color=(0, 0, 0)):
        self.xyz = np.array(xyz, np.float64)
        self.track = np.array(track, np.int64).reshape(-1, 2)
        self.error = float(error)
        self.color = np.array(color, np.uint8)


class Model:
    """cameras / images / points3D keyed by id, as Reconstruction holds them."""

    def __init__(self, cameras=None, images=None, points3D=None):
        self.cameras = dict(cameras or {})
        self.images = dict(images or {})
        self.points3D = dict(points3D or {})

    def num_observations(self):                      # ComputeNumObservations (:486-492)
        return int(sum(int((im.point3D_ids >= 0).sum()) for im in self.images.values()))

    # ---- the flat track-major problem of the C-ABI (include/ppsfm_b200.h: ppsfm_filter_problem) --
    def to_filter_problem(self):
        """(FilterProblem, image_ids, camera_ids, point3D_ids): images, cameras and points in
        ascending id order; a point's observations in the order of its track."""
        cam_ids, img_ids, pt_ids = sorted(self.cameras), sorted(self.images), sorted(self.points3D)
        cam_index = {c: k for k, c in enumerate(cam_ids)}
        img_index = {i: k for k, i in enumerate(img_ids)}
        track_start, obs_image, obs_line, obs_aligned = [0], [], [], []
        for p in pt_ids:
            for image_id, line_idx in self.points3D[p].track:
                im = self.images[int(image_id)]
                obs_image.append(img_index[int(image_id)])
                obs_line.append(im.lines[line_idx])
                obs_aligned.append(im.aligned[line_idx])
            track_start.append(len(obs_image))
        pb = FilterProblem(
            np.array([self.images[i].qvec for i in img_ids]).reshape(-1, 4),
            np.array([self.images[i].tvec for i in img_ids]).reshape(-1, 3),
            [cam_index[self.images[i].camera_id] for i in img_ids],
            [self.cameras[c].model_id for c in cam_ids], [self.cameras[c].params for c in cam_ids],
            [(self.cameras[c].width, self.cameras[c].height) for c in cam_ids],
            np.array([self.points3D[p].xyz for p in pt_ids]).reshape(-1, 3), track_start, obs_image,
            np.array(obs_line).reshape(-1, 3), np.array(obs_aligned, np.uint8))
        return pb, img_ids, cam_ids, pt_ids

    @staticmethod
    def from_filter_problem(pb, obs_deleted=None, point_deleted=None, point_error=None):
        """The reconstruction oracle/ref/ref_filter.cc builds from a FilterProblem through the
        reference's own members: ids are index + 1, the lines of an image are its observations in
        observation order, names image%06d.jpg; ``obs_deleted`` / ``point_deleted`` /
        ``point_error`` apply the result of a filter call (lines lose their point, points go)."""
        n_img, n_pt = len(pb.qvecs), len(pb.points)
        od = np.zeros(len(pb.obs_image), bool) if obs_deleted is None else np.asarray(obs_deleted, bool)
        pd = np.zeros(n_pt, bool) if point_deleted is None else np.asarray(point_deleted, bool)
        obs_point = np.repeat(np.arange(n_pt), np.diff(pb.track_start))
        line_idx = np.zeros(len(pb.obs_image), np.int64)
        per_image = [[] for _ in range(n_img)]
        for k, i in enumerate(pb.obs_image):
            line_idx[k] = len(per_image[i])
            per_image[i].append(k)
        cams = {c + 1: Camera(pb.camera_model[c], pb.camera_width[c], pb.camera_height[c],
                              pb.camera_params[c, :CAMERA_MODEL_NUM_PARAMS[int(pb.camera_model[c])]])
                for c in range(len(pb.camera_model))}
        images = {}
        for i in range(n_img):
            ks = np.array(per_image[i], np.int64)
            ids = np.where(od[ks] | pd[obs_point[ks]], -1, obs_point[ks] + 1) if len(ks) else []
            images[i + 1] = Image(pb.qvecs[i], pb.tvecs[i], pb.image_camera[i] + 1,
                                  "image%06d.jpg" % i, pb.obs_line[ks], pb.obs_aligned[ks] != 0, ids)
        points = {}
        for p in range(n_pt):
            ks = np.arange(pb.track_start[p], pb.track_start[p + 1])
            if pd[p] or len(ks) == 0:
                continue
            ks = ks[~od[ks]]
            points[p + 1] = Point3D(pb.points[p], np.stack([pb.obs_image[ks] + 1, line_idx[ks]], 1),
                                    -1.0 if point_error is None else point_error[p])
        return Model(cams, images, points)


def _g17(x):
    return "%.17g" % x


def write_model_text(path, model, reference_precision=True):
    """Reconstruction::WriteText (:549-553): the three files into directory ``path``."""
    os.makedirs(path, exist_ok=True)
    _g6 = (lambda x: "%g" % x) if reference_precision else _g17   # the rows built in an ostringstream
    num_obs = model.num_observations()
    with open(os.path.join(path, "cameras.txt"), "w") as f:
        f.write("# Camera list with one line of data per camera:\n")
        f.write("#   CAMERA_ID, MODEL, WIDTH, HEIGHT, PARAMS[]\n")
        f.write("# Number of cameras: %d\n" % len(model.cameras))
        for cid in sorted(model.cameras):
            cam = model.cameras[cid]
            row = "%d %s %d %d " % (cid, cam.model_name, cam.width, cam.height)
            row += "".join(_g6(x) + " " for x in cam.params)
            f.write(row[:-1] + "\n")
    with open(os.path.join(path, "images.txt"), "w") as f:
        f.write("# Image list with two lines of data per image:\n")
        f.write("#   IMAGE_ID, QW, QX, QY, QZ, TX, TY, TZ, CAMERA_ID, NAME\n")
        f.write("#   LINES2D[] as (A, B, C, is_aligned, POINT3D_ID)\n")
        mean_obs = num_obs / float(len(model.images)) if model.images else 0.0
        f.write("# Number of images: %d, mean observations per image: %s\n"
                % (len(model.images), _g17(mean_obs)))
        for iid in sorted(model.images):
            im = model.images[iid]
            q = normalize_quaternion(im.qvec)
            f.write("%d %s %s %d %s\n" % (iid, " ".join(_g6(x) for x in q),
                                          " ".join(_g6(x) for x in im.tvec), im.camera_id, im.name))
            row = "".join("%s %s %s %s %d " % (_g6(l[0]), _g6(l[1]), _g6(l[2]), "1" if a else "0", p)
                          for l, a, p in zip(im.lines.tolist(), im.aligned.tolist(),
                                             im.point3D_ids.tolist()))
            f.write(row[:-1] + "\n")
    with open(os.path.join(path, "points3D.txt"), "w") as f:
        f.write("# 3D point list with one line of data per point:\n")
        f.write("#   POINT3D_ID, X, Y, Z, R, G, B, ERROR, TRACK[] as (IMAGE_ID, line_idx)\n")
        mean_len = num_obs / float(len(model.points3D)) if model.points3D else 0.0
        f.write("# Number of points: %d, mean track length: %s\n" % (len(model.points3D), _g17(mean_len)))
        for pid in sorted(model.points3D):
            pt = model.points3D[pid]
            head = "%d %s %s %s %d %d %d %s " % (pid, _g17(pt.xyz[0]), _g17(pt.xyz[1]), _g17(pt.xyz[2]),
                                                 pt.color[0], pt.color[1], pt.color[2], _g17(pt.error))
            row = "".join("%d %d " % (i, l) for i, l in pt.track.tolist())
            f.write(head + row[:-1] + "\n")


def normalize_quaternion(qvec):
    """NormalizeQuaternion (src/base/pose.cc:53-62)."""
    q = np.asarray(qvec, np.float64)
    norm = float(np.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]))
    if norm == 0:
        return np.array([1.0, q[1], q[2], q[3]])
    return q / norm


def _rotate(q, v):
    """Eigen's quaternion * vector (Quaternion.h, _transformVector): v + w 2(u x v) + u x 2(u x v)."""
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    v0, v1, v2 = v[..., 0], v[..., 1], v[..., 2]
    u0, u1, u2 = y * v2 - z * v1, z * v0 - x * v2, x * v1 - y * v0
    u0, u1, u2 = u0 + u0, u1 + u1, u2 + u2
    return np.stack([v0 + w * u0 + (y * u2 - z * u1), v1 + w * u1 + (z * u0 - x * u2),
                     v2 + w * u2 + (x * u1 - y * u0)], -1)


def projection_centers(qvecs, tvecs):
    """ProjectionCenterFromPose (src/base/pose.cc:94-101): conj(q / |q|) * (-t)."""
    q = np.asarray(qvecs, np.float64).reshape(-1, 4)
    norm = np.sqrt(q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2] + q[:, 3] * q[:, 3])
    qn = np.where(norm[:, None] == 0, np.concatenate([np.ones((len(q), 1)), q[:, 1:]], 1),
                  q / np.where(norm == 0, 1.0, norm)[:, None])
    return _rotate(qn * np.array([1.0, -1.0, -1.0, -1.0]), -np.asarray(tvecs, np.float64).reshape(-1, 3))


def normalize_scene(qvecs, tvecs, points, extent=10.0, p0=0.1, p1=0.9, use_images=True):
    """Reconstruction::Normalize (src/base/reconstruction.cc:302-398) on arrays: the registered
    images' (qvecs [n, 4], tvecs [n, 3]) and the points [p, 3] -> (tvecs, points, scale,
    translation).  The mapper calls it after every global bundle adjustment
    (src/sfm/incremental_mapper.cc:934-936).  As the reference: coordinates (projection centres,
    or points with ``use_images=False``) are cast to FLOAT and every axis is sorted on its own;
    the box spans the p0 / p1 percentiles, the translation is the mean of the sorted values
    between them, the scale ``extent`` over the box diagonal; a new tvec is q * (-(c - t) s) with
    the image's quaternion as stored."""
    q = np.asarray(qvecs, np.float64).reshape(-1, 4)
    t = np.array(tvecs, np.float64).reshape(-1, 3)
    X = np.array(points, np.float64).reshape(-1, 3)
    if (use_images and len(q) < 2) or (not use_images and len(X) < 2):
        return t, X, 1.0, np.zeros(3)
    centers = projection_centers(q, t)
    coords = np.sort((centers if use_images else X).astype(np.float32), axis=0)
    n = len(coords)
    P0 = int(p0 * (n - 1)) if n > 3 else 0
    P1 = int(p1 * (n - 1)) if n > 3 else n - 1
    bbox_min, bbox_max = coords[P0].astype(np.float64), coords[P1].astype(np.float64)
    translation = np.cumsum(coords[P0:P1 + 1].astype(np.float64), axis=0)[-1] / float(P1 - P0 + 1)
    d = bbox_max - bbox_min
    old_extent = float(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
    scale = 1.0 if old_extent < np.finfo(np.float64).eps else extent / old_extent
    new_t = _rotate(q, -((centers - translation) * scale))
    return new_t, (X - translation) * scale, scale, translation


def _rows(path):
    with open(path) as f:
        for ln in f:
            yield ln.strip()                         # StringTrim


def read_model_text(path, reference_precision=True):
    """Reconstruction::ReadText (:543-547).  Raises ValueError where the reference CHECK-fails or
    its std::sto* conversions throw."""
    num = _stold if reference_precision else float
    cameras, images, points = {}, {}, {}
    for ln in _rows(os.path.join(path, "cameras.txt")):
        if not ln or ln[0] == "#":
            continue
        t = ln.split(" ")
        if t[1] not in CAMERA_MODEL_IDS:
            raise ValueError("cameras.txt: unknown camera model %r" % t[1])
        cam = Camera(CAMERA_MODEL_IDS[t[1]], _stoi(t[2]), _stoi(t[3]), [num(x) for x in t[4:]])
        if len(cam.params) != CAMERA_MODEL_NUM_PARAMS[cam.model_id]:    # CHECK(camera.VerifyParams())
            raise ValueError("cameras.txt: %s takes %d parameters, %d given"
                             % (t[1], CAMERA_MODEL_NUM_PARAMS[cam.model_id], len(cam.params)))
        cameras[_stoi(t[0])] = cam
    rows = _rows(os.path.join(path, "images.txt"))
    for ln in rows:
        if not ln or ln[0] == "#":
            continue
        t = ln.split(" ")
        qvec = np.array([num(x) for x in t[1:5]])
        if reference_precision:
            qvec = normalize_quaternion(qvec)
        tvec = np.array([num(x) for x in t[5:8]])
        camera_id, name = _stoi(t[8]), (t[9] if len(t) > 9 else "")
        row = next(rows, None)                       # LINES2D: the next line, whatever it holds
        if row is None:
            break
        if row:
            items = row.split(" ")
            if len(items) % 5:
                raise ValueError("images.txt: LINES2D row of image %s is not a list of 5-tuples" % t[0])
            flags = items[3::5]
            if any(x not in ("0", "1") for x in flags):                 # CHECK(item == "0")
                raise ValueError("images.txt: is_aligned must be 0 or 1")
            coeff = [x for k in range(0, len(items), 5) for x in items[k:k + 3]]
            if reference_precision:
                lines = _stof_row(coeff).reshape(-1, 3)
                lines = lines / np.sqrt(lines[:, 0] * lines[:, 0] + lines[:, 1] * lines[:, 1])[:, None]
            else:
                lines = np.array(coeff, dtype=np.float64).reshape(-1, 3)
            ids = np.array([_stoi(x) for x in items[4::5]], np.int64)
            images[_stoi(t[0])] = Image(qvec, tvec, camera_id, name, lines,
                                      [x == "1" for x in flags], ids)
        else:
            images[_stoi(t[0])] = Image(qvec, tvec, camera_id, name, np.zeros((0, 3)), [], [])
    for ln in _rows(os.path.join(path, "points3D.txt")):
        if not ln or ln[0] == "#":
            continue
        t = ln.split(" ")
        track = []
        for k in range(8, len(t), 2):
            if not t[k].strip():
                break
            # a dangling IMAGE_ID: the second std::getline fails at the end of the row and leaves
            # the item as it was, so the reference reads the id again as the line index (:944-951)
            track.append((_stoi(t[k]), _stoi(t[k + 1] if k + 1 < len(t) else t[k])))
        points[_stoi(t[0])] = Point3D([num(x) for x in t[1:4]], track, num(t[7]),
                                    [_stoi(x) & 0xFF for x in t[4:7]])
    return Model(cameras, images, points)


def Read(path, reference_precision=True):
    """Reconstruction::Read (:527-536): this fork reads the text files only and fails without them."""
    if not all(os.path.exists(os.path.join(path, n)) for n in ("cameras.txt", "images.txt", "points3D.txt")):
        raise FileNotFoundError("cameras, images, points3D files do not exist at " + path)
    return read_model_text(path, reference_precision)


def Write(path, model, reference_precision=True):
    """Reconstruction::Write (:538-540): text."""
    write_model_text(path, model, reference_precision)
