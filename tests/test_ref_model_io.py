"""Text model format: privacy_preserving_sfm_b200/model_io.py against the
REFERENCE'S OWN Reconstruction::ReadText / WriteText (CPU).

oracle/ref/ref_filter.cc (built by oracle/build_ref.sh into oracle/_ref/libref_filter.so from
src/base/reconstruction.cc and the classes it uses, where they lie under /root/reference) builds a
colmap::Reconstruction through the reference's own members, lets the reference write it
(cameras.txt / images.txt / points3D.txt) and read directories back.  Pinned here:
  * the library's writer produces the reference's records and header comments byte for byte
    (record order aside: the reference walks unordered maps) — SIX significant digits of camera
    parameters, poses and lines (the reference sets precision(17) on the file stream only and
    builds those rows in an ostringstream), 17 of the points;
  * the reference reads what the library writes, the library reads what the reference writes,
    and both readers return the same numbers bit for bit — including what the reference's reader
    does to them (lines through std::stof and renormalised, quaternion normalised, std::stold);
  * tokens on which parsing to float / long double first and parsing to double differ;
  * the flat track-major problem of the C-ABI built from a model that was read.
The mapper's own write_text / read_text (tests/test_gpu_mapper.py runs them after a GPU
reconstruction) are checked on a hand-made mapper state.

Skipped where neither oracle/_ref/libref_filter.so nor /root/reference exists."""
from fractions import Fraction
import os

import numpy as np
import pytest

from privacy_preserving_sfm_b200 import model_io as IO
from test_ref_filters import MODELS, _problem


@pytest.fixture(scope="module")
def ref():
    import oracle.reference as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_filter.so not built and /root/reference absent")
    return R


def _records(path, name):
    with open(os.path.join(path, name)) as f:
        rows = f.read().split("\n")
    assert rows[-1] == ""
    rows = rows[:-1]
    head = [r for r in rows if r.startswith("#")]
    body = [r for r in rows if not r.startswith("#")]
    if name == "images.txt":                               # two rows per image
        body = [(body[k], body[k + 1]) for k in range(0, len(body), 2)]
    return head, sorted(body)


def _same_as_reference_dump(model, d):
    """model_io.Model against the flat dump of the reference's Reconstruction, bit for bit."""
    u = lambda a: np.ascontiguousarray(a, np.float64).view(np.uint64)
    assert sorted(model.cameras) == d["cam_id"].tolist()
    for k, c in enumerate(sorted(model.cameras)):
        cam = model.cameras[c]
        assert (cam.model_id, cam.width, cam.height) == (d["cam_model"][k], *d["cam_size"][k])
        assert len(cam.params) == d["cam_num_params"][k]
        assert np.array_equal(u(cam.params), u(d["cam_params"][k, :len(cam.params)]))
    assert sorted(model.images) == d["img_id"].tolist() and d["num_reg_images"] == len(model.images)
    for k, i in enumerate(sorted(model.images)):
        im = model.images[i]
        assert np.array_equal(u(im.qvec), u(d["img_qvec"][k])), (i, im.qvec, d["img_qvec"][k])
        assert np.array_equal(u(im.tvec), u(d["img_tvec"][k]))
        assert im.camera_id == d["img_camera"][k] and im.name == d["img_name"][k]
        s, e = d["line_start"][k], d["line_start"][k + 1]
        assert np.array_equal(u(im.lines), u(d["lines"][s:e]))
        assert np.array_equal(im.aligned, d["aligned"][s:e] != 0)
        assert np.array_equal(im.point3D_ids, d["line_point"][s:e])
    assert sorted(model.points3D) == d["pt_id"].tolist()
    for k, p in enumerate(sorted(model.points3D)):
        pt = model.points3D[p]
        assert np.array_equal(u(pt.xyz), u(d["pt_xyz"][k]))
        assert np.array_equal(u([pt.error]), u(d["pt_error"][k:k + 1]))
        assert np.array_equal(pt.color, d["pt_color"][k])
        s, e = d["track_start"][k], d["track_start"][k + 1]
        assert np.array_equal(pt.track[:, 0], d["track_image"][s:e])
        assert np.array_equal(pt.track[:, 1], d["track_line"][s:e])


@pytest.mark.parametrize("model,params", [MODELS[1], MODELS[4], MODELS[7]])
def test_writer_produces_the_references_records(oracle, ref, tmp_path, model, params):
    pb = _problem(9, 300, 6, seed=40 + model, model=model, params=params, varlen=True)
    _, od, pd, pe, _ = oracle.filter_points3d(pb, 4.0, 1.5)
    assert 0 < pd.sum() < len(pd) and od.sum() > 0        # lines without a point, deleted points
    a, b = str(tmp_path / "reference"), str(tmp_path / "library")
    ref.model_write_text(pb, a, filter_thresholds=(4.0, 1.5))
    IO.write_model_text(b, IO.Model.from_filter_problem(pb, od, pd, pe))
    for name in ("cameras.txt", "images.txt", "points3D.txt"):
        (ha, ra), (hb, rb) = _records(a, name), _records(b, name)
        assert ha == hb, name                               # comments incl. the two means
        assert ra == rb, name
    # ... and unfiltered (every line has a point, errors unset: -1)
    ref.model_write_text(pb, a)
    IO.write_model_text(b, IO.Model.from_filter_problem(pb))
    for name in ("cameras.txt", "images.txt", "points3D.txt"):
        assert _records(a, name) == _records(b, name), name


@pytest.mark.parametrize("model,params", [MODELS[1], MODELS[3], MODELS[7]])
def test_both_readers_agree_on_both_writers(oracle, ref, tmp_path, model, params):
    pb = _problem(8, 250, 5, seed=60 + model, model=model, params=params, varlen=True)
    _, od, pd, pe, _ = oracle.filter_points3d(pb, 4.0, 1.5)
    written = IO.Model.from_filter_problem(pb, od, pd, pe)
    a, b, c = (str(tmp_path / n) for n in ("library", "reference", "rewritten"))
    IO.write_model_text(a, written, reference_precision=False)               # 17 digits everywhere
    ref.model_write_text(pb, b, filter_thresholds=(4.0, 1.5))               # six digits of poses / lines
    for path in (a, b):
        d = ref.model_read_text(path, rewrite_to=c)
        _same_as_reference_dump(IO.read_model_text(path), d)
        # the reference writes what it has read (float-rounded lines): both readers again
        _same_as_reference_dump(IO.read_model_text(c), ref.model_read_text(c))
    # what the reader keeps of what was written: everything but the line coefficients, which come
    # back as renormalised floats (reconstruction.cc:846-870) ...
    back, exact = IO.read_model_text(a), IO.read_model_text(a, reference_precision=False)
    assert sorted(back.images) == sorted(written.images) and sorted(back.points3D) == sorted(written.points3D)
    for i, im in written.images.items():
        assert np.array_equal(exact.images[i].lines, im.lines)               # 17 digits: lossless
        assert np.array_equal(exact.images[i].qvec, IO.normalize_quaternion(im.qvec))
        assert np.array_equal(back.images[i].tvec, im.tvec)
        assert np.array_equal(back.images[i].point3D_ids, im.point3D_ids)
        assert np.array_equal(back.images[i].aligned, im.aligned)
        if len(im.lines):
            diff = np.abs(back.images[i].lines - im.lines).max()
            assert 0 < diff < 1e-6 * max(1.0, np.abs(im.lines).max())
            assert np.abs(np.hypot(back.images[i].lines[:, 0], back.images[i].lines[:, 1]) - 1).max() < 1e-15
    for p, pt in written.points3D.items():
        assert np.array_equal(back.points3D[p].xyz, pt.xyz) and back.points3D[p].error == pt.error
        assert np.array_equal(back.points3D[p].track, pt.track)


def _decimal(fr, digits=60):
    """Exact decimal expansion of a dyadic fraction."""
    n, d = fr.numerator, fr.denominator
    s = str(n // d) + "."
    r = n % d
    for _ in range(digits):
        r *= 10
        s += str(r // d)
        r %= d
    return s


def test_tokens_where_float_and_long_double_parsing_differ(ref, tmp_path):
    """std::stof rounds the decimal to float ONCE (through double it would be rounded twice);
    std::stold rounds to a 64-bit significand and the assignment to double rounds again."""
    just_above_float_tie = _decimal(Fraction(1) + Fraction(1, 2 ** 24) + Fraction(1, 2 ** 90), 100)
    assert np.float32(float(just_above_float_tie)) == 1.0                    # twice: ties to even
    double_tie_via_long_double = _decimal(Fraction(1) + Fraction(1, 2 ** 53) + Fraction(1, 2 ** 70), 80)
    assert float(double_tie_via_long_double) == 1.0 + 2.0 ** -52             # direct: rounds up
    path = str(tmp_path / "m")
    os.makedirs(path)
    with open(os.path.join(path, "cameras.txt"), "w") as f:
        f.write("# c\n1 SIMPLE_RADIAL 640 480 %s 320 240 0.01\n" % double_tie_via_long_double)
    with open(os.path.join(path, "images.txt"), "w") as f:
        f.write("7 0.5 0.5 0.5 0.5 %s 2 3 1 a.jpg\n" % double_tie_via_long_double)
        f.write("%s 0.25 3 1 -1 0.1 0.7 -3e-5 0 4\n" % just_above_float_tie)
        f.write("\n# comment\n9 2 0 0 0 0 0 0 1 b.jpg\n\n")                 # an image without lines
    with open(os.path.join(path, "points3D.txt"), "w") as f:
        f.write("4 1 2 %s 255 0 7 0.5 7 1\n" % double_tie_via_long_double)
    m = IO.read_model_text(path)
    _same_as_reference_dump(m, ref.model_read_text(path))
    assert m.cameras[1].params[0] == 1.0 and m.images[7].tvec[0] == 1.0 and m.points3D[4].xyz[2] == 1.0
    a = float(np.float32(1.0) + np.float32(2.0 ** -23))
    assert m.images[7].lines[0, 0] == a / np.sqrt(a * a + 0.25 * 0.25)
    assert m.images[9].lines.shape == (0, 3) and np.array_equal(m.images[9].qvec, [1.0, 0, 0, 0])
    assert np.array_equal(m.points3D[4].color, [255, 0, 7])


def test_flat_problem_of_a_model_that_was_read(oracle, ref, tmp_path):
    pb = _problem(8, 300, 6, seed=77, varlen=True)
    path = str(tmp_path / "m")
    IO.write_model_text(path, IO.Model.from_filter_problem(pb), reference_precision=False)
    model = IO.read_model_text(path, reference_precision=False)
    pb2, img_ids, cam_ids, pt_ids = model.to_filter_problem()
    keep = np.diff(pb.track_start) > 0                      # AddPoint3D is skipped for empty tracks
    assert img_ids == list(range(1, 9)) and cam_ids == [1]
    assert np.array_equal(np.array(pt_ids) - 1, np.flatnonzero(keep))
    assert np.array_equal(pb2.points, pb.points[keep]) and np.array_equal(pb2.tvecs, pb.tvecs)
    assert np.array_equal(np.diff(pb2.track_start), np.diff(pb.track_start)[keep])
    assert np.array_equal(pb2.obs_image, pb.obs_image) and np.array_equal(pb2.obs_line, pb.obs_line)
    assert np.array_equal(pb2.obs_aligned, pb.obs_aligned)
    assert np.array_equal(pb2.camera_params, pb.camera_params)
    # the filter of the reference's reconstruction == the filter of the flat problem read back
    nf, od, pd, pe, _ = oracle.filter_points3d(pb, 4.0, 1.5)
    nf2, od2, pd2, pe2, _ = oracle.filter_points3d(pb2, 4.0, 1.5)
    assert nf == nf2 and np.array_equal(od, od2) and np.array_equal(pd[keep], pd2)
    # with the reference's reader precision the lines are floats: same problem to 1e-7; written
    # by the reference (six digits of poses, lines and camera parameters): to 1e-5
    pb3 = IO.read_model_text(path).to_filter_problem()[0]
    assert 0 < np.abs(pb3.obs_line - pb.obs_line).max() < 1e-6 and np.array_equal(pb3.tvecs, pb.tvecs)
    ref.model_write_text(pb, path)
    pb4 = IO.read_model_text(path).to_filter_problem()[0]
    assert np.array_equal(pb4.obs_image, pb.obs_image) and np.array_equal(pb4.points, pb.points[keep])
    assert 1e-8 < np.abs(pb4.obs_line - pb.obs_line).max() < 1e-4
    assert 1e-8 < np.abs(pb4.tvecs - pb.tvecs).max() < 1e-4


def test_reader_rejects_what_the_reference_rejects(tmp_path):
    path = str(tmp_path / "m")
    os.makedirs(path)
    files = {"cameras.txt": "1 PINHOLE 640 480 500 500 320 240\n",
             "images.txt": "1 1 0 0 0 0 0 0 1 a.jpg\n0.6 0.8 1 1 -1\n", "points3D.txt": ""}

    def write(**over):
        for name, text in {**files, **over}.items():
            with open(os.path.join(path, name), "w") as f:
                f.write(text)

    write()
    assert len(IO.read_model_text(path).images[1].lines) == 1 and len(IO.Read(path).cameras) == 1
    with pytest.raises(FileNotFoundError):                                   # Reconstruction::Read: LOG(FATAL)
        IO.Read(str(tmp_path / "nothing"))
    write(**{"cameras.txt": "1 PINHOLE 640 480 500 500 320\n"})              # CHECK(VerifyParams)
    with pytest.raises(ValueError):
        IO.read_model_text(path)
    write(**{"cameras.txt": "1 NOT_A_MODEL 640 480 500 500 320 240\n"})
    with pytest.raises(ValueError):
        IO.read_model_text(path)
    write(**{"images.txt": "1 1 0 0 0 0 0 0 1 a.jpg\n0.6 0.8 1 true -1\n"})   # CHECK(item == "0")
    with pytest.raises(ValueError):
        IO.read_model_text(path)
    write(**{"images.txt": "1 1 0 0 0 0 0 0 1 a.jpg\n0.6  0.8 1 1 -1\n"})     # getline(' '): empty item
    with pytest.raises(ValueError):
        IO.read_model_text(path)
    for token in ("1e-42", "1e39"):                                          # std::stof: out_of_range
        write(**{"images.txt": "1 1 0 0 0 0 0 0 1 a.jpg\n0.6 0.8 %s 1 -1\n" % token})
        with pytest.raises(ValueError):
            IO.read_model_text(path)
        assert len(IO.read_model_text(path, reference_precision=False).images[1].lines) == 1


def test_mapper_text_round_trip_on_a_hand_made_state(tmp_path):
    """IncrementalMapper.write_text / read_text without a GPU: the assertions of
    tests/test_gpu_mapper.py::test_mapper_controller_schedule_with_local_ba on a state set by hand."""
    from privacy_preserving_sfm_b200 import mapper as M
    scene, gt = M.make_mapper_scene(num_images=7, num_points=60, seed=3, visibility=0.6)
    rng = np.random.default_rng(0)
    m = M.IncrementalMapper.__new__(M.IncrementalMapper)
    m.scene = scene
    m.registered = [0, 2, 3, 5, 6]
    m.qvec = rng.normal(size=(7, 4))
    m.qvec /= np.linalg.norm(m.qvec, axis=1)[:, None]               # unit, as the mapper keeps them
    m.tvec = rng.normal(size=(7, 3))
    m.points = rng.normal(size=(60, 3))
    m.has_point = rng.uniform(size=60) < 0.7
    m.obs_on = scene.visible & (rng.uniform(size=scene.visible.shape) < 0.8)
    # Reconstruction::Normalize on the mapper's state: registered images and points move together
    before = IO.projection_centers(m.qvec[m.registered], m.tvec[m.registered])
    pts_before, t_unreg = m.points.copy(), m.tvec[[1, 4]].copy()
    scale = m.normalize_scene()
    after = IO.projection_centers(m.qvec[m.registered], m.tvec[m.registered])
    d0, d1 = np.linalg.norm(before[0] - before[1]), np.linalg.norm(after[0] - after[1])
    assert abs(d1 - scale * d0) < 1e-9 * d1 and np.array_equal(m.tvec[[1, 4]], t_unreg)
    has = m.has_point
    assert np.array_equal(m.points[~has], pts_before[~has])
    assert abs(np.linalg.norm(m.points[has][0] - m.points[has][1])
               - scale * np.linalg.norm(pts_before[has][0] - pts_before[has][1])) < 1e-9 * scale
    out = str(tmp_path / "model")
    m.write_text(out)
    model = M.IncrementalMapper.read_text(out)
    assert len(model["cameras"]) == 1 and model["cameras"][1][0] == "PINHOLE"
    assert sorted(model["images"]) == sorted(i + 1 for i in m.registered)
    assert len(model["points"]) == int(m.has_point.sum())
    for img_id, (q, t, cam_id, name, lines) in model["images"].items():
        i = img_id - 1
        assert np.array_equal(q, IO.normalize_quaternion(m.qvec[i])) and np.array_equal(t, m.tvec[i])
        vis = np.flatnonzero(scene.visible[i])
        assert np.array_equal(lines[:, :3], scene.lines[i, vis])            # 17 digits: exact
        has = m.obs_on[i, vis] & m.has_point[vis]
        assert np.array_equal(lines[:, 4], np.where(has, vis + 1, -1))
        assert np.array_equal(lines[:, 3], scene.aligned[vis])
    for pid, (xyz, err, track) in model["points"].items():
        assert np.array_equal(xyz, m.points[pid - 1]) and err == -1.0
        assert [i - 1 for i in track[:, 0]] == [i for i in m.registered if m.obs_on[i, pid - 1]]
        for img_id, line_idx in track:                                       # track elements point back
            assert model["images"][img_id][4][line_idx, 4] == pid


@pytest.mark.parametrize("num_cams,use_images,p", [(12, True, (0.1, 0.9)), (3, True, (0.1, 0.9)),
                                                   (12, False, (0.1, 0.9)), (9, True, (0.0, 1.0)),
                                                   (30, True, (0.25, 0.6))])
def test_normalize_is_the_references(ref, num_cams, use_images, p):
    """model_io.normalize_scene against the reference's own Reconstruction::Normalize (what the
    mapper applies after every global bundle adjustment): bit for bit."""
    pb = _problem(num_cams, 200, min(4, num_cams), seed=90 + num_cams)
    tv, pts, scale, translation = IO.normalize_scene(pb.qvecs, pb.tvecs, pb.points, 10.0, p[0], p[1],
                                                     use_images)
    tv_ref, pts_ref = ref.normalize(pb, 10.0, p[0], p[1], use_images)
    assert np.array_equal(tv.view(np.uint64), tv_ref.view(np.uint64))
    assert np.array_equal(pts.view(np.uint64), pts_ref.view(np.uint64))
    # a similarity: line residuals of the model are unchanged (up to rounding), the extent is 10
    c = IO.projection_centers(pb.qvecs, tv)
    if use_images and p == (0.0, 1.0):
        assert abs(np.linalg.norm(c.max(axis=0) - c.min(axis=0)) - 10.0) < 1e-5
    c0 = IO.projection_centers(pb.qvecs, pb.tvecs)
    assert np.abs((c0 - translation) * scale - c).max() < 1e-9 * max(1.0, scale)
    # fewer than two images: untouched
    one = IO.normalize_scene(pb.qvecs[:1], pb.tvecs[:1], pb.points)
    assert np.array_equal(one[0], pb.tvecs[:1]) and one[2] == 1.0


def test_unusual_but_valid_number_tokens_read_as_the_reference_reads_them(ref, tmp_path):
    """std::sto* accept what strtod-style parsing accepts: signs, exponents, missing digits on one
    side of the point, hexadecimal floats, trailing characters, nan / inf."""
    path = str(tmp_path / "m")
    os.makedirs(path)
    with open(os.path.join(path, "cameras.txt"), "w") as f:
        f.write("3 OPENCV 640 480 +5.0e2 0x1.f4p8 .32e3 240. 1e-2 -1E-3 0 0x0p0\n")
    with open(os.path.join(path, "images.txt"), "w") as f:
        f.write("0012 1e0 0 -0 0.0 1.5abc +2 -.5e1 3 name_with_underscore.png\n")
        f.write("6e-1 0x1.999999999999ap-1 1 1 9 -.6 +.8 2.5E0 0 -1 3 4junk 5 0 7\n")
        f.write("5 nan 0 0 1 inf 0 0 3 x.jpg\n\n")
    with open(os.path.join(path, "points3D.txt"), "w") as f:
        f.write("9 1e1 -2.E0 .5 300 -1 7 -1 12 0\n7 0 0 0 1 2 3 0x1p-3 12 2 5\n")
    m = IO.read_model_text(path)
    _same_as_reference_dump(m, ref.model_read_text(path))
    assert m.cameras[3].params[1] == 500.0 and m.images[12].tvec[0] == 1.5
    assert m.images[12].lines.shape == (3, 3) and np.isnan(m.images[5].qvec).all()
    assert np.array_equal(m.points3D[9].color, [300 & 255, 255, 7])          # static_cast<uint8_t>
    assert np.array_equal(m.points3D[7].track, [[12, 2], [5, 5]])             # a dangling id is read twice
