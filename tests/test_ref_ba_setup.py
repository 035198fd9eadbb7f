"""Bundle-adjustment ASSEMBLY: the product's adaptor against the
REFERENCE'S OWN BundleAdjuster::SetUp (CPU).

oracle/build_ref.sh compiles src/optim/bundle_adjustment.cc and the classes it reads
(src/base/{image,point3d,track,camera,camera_models}.cc, src/util/*.cc) from where they lie under
/root/reference against the stand-ins of oracle/ref/shim/, where ceres::Problem is a RECORDER: one
call of the reference's BundleAdjuster::Solve yields the residual blocks (which functor — variable
or constant pose —, which image / point / camera), the constant parameter blocks and the
parameterisations (quaternion, constant tvec components, constant intrinsics groups) the reference
would hand to Ceres.  The same colmap::Reconstruction and an identical BundleAdjustmentConfig then
go through ppsfm::BundleAdjuster<colmap::Reconstruction>::AssembleOnly (cpp/ppsfm_adaptor.h: the
flat problem of the C-ABI).  Both sides are written in one canonical text form and must be equal:
same observations with the same pose treatment and loss, same constant / variable poses with the
same gauge components, same constant points, same variable intrinsics per camera.  What Ceres
does with the problem afterwards stays unpinned.

Skipped where neither oracle/_ref/libref_ba_setup.so nor /root/reference exists."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import synthetic as S


@pytest.fixture(scope="module")
def ref():
    import oracle.reference as R
    if not R.available():
        pytest.skip("oracle/_ref/libref_ba_setup.so not built and /root/reference absent")
    return R


def _scene(num_images=8, num_points=60, obs=4, seed=3, cameras=None, unmatched=5):
    """A small reconstruction: every image also carries a few lines without a 3-D point."""
    sb = S.make_ba_scene(num_cams=num_images, num_points=num_points, obs_per_point=obs, seed=seed)
    rng = np.random.default_rng(seed)
    img, pt, ln = list(sb["obs_cam"]), list(sb["obs_pt"]), [list(l) for l in sb["obs_line"]]
    for i in range(num_images):
        for _ in range(unmatched):
            th = rng.uniform(0, 2 * np.pi)
            img.append(i)
            pt.append(-1)
            ln.append([np.cos(th), np.sin(th), rng.uniform(-0.3, 0.3)])
    img, pt, ln = np.array(img), np.array(pt), np.array(ln)
    order = np.lexsort((rng.permutation(len(img)), img))       # image-major, shuffled inside
    img, pt, ln = img[order], pt[order], ln[order]
    if cameras is None:
        cameras = ([1], [[1000.0, 1000.0, 500.0, 500.0]], np.zeros(num_images, np.int32))
    models, params, image_camera = cameras
    cp = np.zeros((len(models), 12))
    for c, p in enumerate(params):
        cp[c, :len(p)] = p
    return dict(camera_model=models, camera_params=cp, image_camera=image_camera,
                qvecs=sb["qvecs"] * 1.7,            # un-normalised on purpose: SetUp normalises
                tvecs=sb["tvecs"], points=sb["points"],
                image_line_start=np.searchsorted(img, np.arange(num_images + 1)).astype(np.int64),
                lines=ln, line_point=pt), (img, pt)


def _same(ref, scene, config, expect_solve=True):
    rc, a, b = ref.ba_setup_compare(scene, config)
    assert rc == (0 if expect_solve else 1), rc
    if a != b:
        import difflib
        raise AssertionError("\n".join(list(difflib.unified_diff(
            a.splitlines(), b.splitlines(), "reference", "product", lineterm=""))[:60]))
    return a


def test_global_bundle_configuration(ref):
    # IncrementalMapper::AdjustGlobalBundle (src/sfm/incremental_mapper.cc:893-939): every
    # registered image, first pose constant, one translation component of the second constant
    scene, _ = _scene()
    text = _same(ref, scene, dict(images=range(8), constant_poses=[0], constant_tvecs={1: [0]}))
    assert "image 1 camera 1 constant 1" in text and "image 2 camera 1 constant 0 tvec_constant_mask 1" in text
    assert text.count("obs image") == 60 * 4 and "point 1 constant 0" in text
    for loss, scale in [(1, 1.0), (2, 0.5)]:
        _same(ref, scene, dict(images=range(8), constant_poses=[0], loss_type=loss, loss_scale=scale))


def test_local_bundle_configuration(ref):
    # AdjustLocalBundle (:781-891): a few images, the points they see are variable points; their
    # observations in images outside the configuration enter through the constant-pose functor,
    # and those cameras are constant
    scene, (img, pt) = _scene(num_images=10, num_points=80, obs=5, seed=5)
    local = [2, 5, 7]
    pts = sorted(set(pt[(np.isin(img, local)) & (pt >= 0)]))
    text = _same(ref, scene, dict(images=local, constant_poses=[2], variable_points=pts,
                                  loss_type=1, loss_scale=1.0))
    assert "pose_constant 1" in text and "pose_constant 0" in text
    # without the variable points: tracks that leave the configuration make their points constant
    text = _same(ref, scene, dict(images=local, constant_tvecs={5: [0, 2]}))
    assert "constant 1" in text.split("point", 1)[1]
    # points listed as constant, and points of the configuration listed explicitly
    _same(ref, scene, dict(images=local, variable_points=pts[::2], constant_points=pts[1::2]))


def test_pose_options(ref):
    scene, _ = _scene()
    text = _same(ref, scene, dict(images=range(8), refine_extrinsics=False))
    assert "constant 0 tvec" not in text                       # every pose constant
    _same(ref, scene, dict(images=[0, 1, 2], constant_poses=[0, 1, 2]))
    _same(ref, scene, dict(images=range(8), constant_tvecs={3: [1], 4: [0, 1, 2]}))


def test_intrinsics_groups(ref):
    # ParameterizeCameras (:490-528) with several cameras and models
    cams = ([2, 4, 1], [[900.0, 500, 480, 0.05], [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003],
                        [1000.0, 1000.0, 500.0, 500.0]], np.array([0, 0, 1, 1, 2, 2, 0, 1], np.int32))
    scene, _ = _scene(cameras=cams)
    for flags in [dict(refine_focal_length=True), dict(refine_extra_params=True),
                  dict(refine_principal_point=True, refine_extra_params=True),
                  dict(refine_focal_length=True, refine_principal_point=True, refine_extra_params=True)]:
        text = _same(ref, scene, dict(images=range(8), constant_poses=[0], **flags))
        assert "variable_mask 0" not in text or flags == dict(refine_extra_params=True)  # PINHOLE has no extras
        text = _same(ref, scene, dict(images=range(8), constant_cameras=[1], **flags))
        assert "camera 2 model 4 variable_mask 0" in text


def test_camera_seen_only_from_outside_the_configuration_is_constant(ref):
    """AddPointToProblem (:476-479): a camera that enters the problem only through images outside
    the configuration is SetConstantCamera'd, whatever the refine_* flags say."""
    cams = ([2, 2], [[900.0, 500, 480, 0.05], [950.0, 510, 470, 0.02]],
            np.array([0, 0, 0, 0, 1, 1, 1, 1], np.int32))
    scene, (img, pt) = _scene(cameras=cams)
    local = [0, 1, 2]                                  # all on camera 1; camera 2 only outside
    pts = sorted(set(pt[(np.isin(img, local)) & (pt >= 0)]))
    text = _same(ref, scene, dict(images=local, variable_points=pts, refine_focal_length=True,
                                  refine_extra_params=True))
    assert "camera 2 model 2 variable_mask 0" in text and "camera 1 model 2 variable_mask 9" in text


def test_empty_problem(ref):
    # no line of the configuration has a 3-D point: problem_->NumResiduals() == 0 -> Solve false
    scene, _ = _scene()
    scene["line_point"] = np.full_like(scene["line_point"], -1)
    _same(ref, scene, dict(images=range(8)), expect_solve=False)


# ---------------------------------------------------------------------------------------------
# The Python mirror (privacy_preserving_sfm_b200/bundle_adjustment.py: BundleAdjuster._SetUp, what
# the mapper driver uses) against the same record
# ---------------------------------------------------------------------------------------------
_GROUPS = {  # camera_models.h: focal / principal point / extra parameter indices per model id
    1: ([0, 1], [2, 3], []), 2: ([0], [1, 2], [3]), 4: ([0, 1], [2, 3], [4, 5, 6, 7])}


def _python_assembly_text(scene, config):
    from privacy_preserving_sfm_b200 import bundle_adjustment as ba
    rec = ba.Reconstruction()
    nparams = {1: 4, 2: 4, 4: 8}
    for c, m in enumerate(scene["camera_model"]):
        rec.cameras[c + 1] = ba.Camera(c + 1, int(m), scene["camera_params"][c][:nparams[int(m)]])
    tracks = {}
    ils = scene["image_line_start"]
    for i in range(len(scene["image_camera"])):
        lines = []
        for k in range(ils[i], ils[i + 1]):
            p = int(scene["line_point"][k])
            lines.append(ba.FeatureLine(scene["lines"][k], False, p + 1 if p >= 0 else ba.kInvalidPoint3DId))
            if p >= 0:
                tracks.setdefault(p + 1, []).append((i + 1, k - ils[i]))
        rec.images[i + 1] = ba.Image(i + 1, int(scene["image_camera"][i]) + 1, scene["qvecs"][i],
                                     scene["tvecs"][i], lines)
    for p in range(len(scene["points"])):
        rec.points3D[p + 1] = ba.Point3D(scene["points"][p], tracks.get(p + 1, []))
    cfg = ba.BundleAdjustmentConfig()
    for i in config.get("images", []):
        cfg.AddImage(int(i) + 1)
    for i in config.get("constant_poses", []):
        cfg.SetConstantPose(int(i) + 1)
    for i, idxs in config.get("constant_tvecs", {}).items():
        cfg.SetConstantTvec(int(i) + 1, list(idxs))
    for p in config.get("variable_points", []):
        cfg.AddVariablePoint(int(p) + 1)
    for p in config.get("constant_points", []):
        cfg.AddConstantPoint(int(p) + 1)
    for c in config.get("constant_cameras", []):
        cfg.SetConstantCamera(int(c) + 1)
    opt = ba.BundleAdjustmentOptions()
    opt.loss_function_type = ba.LossFunctionType(config.get("loss_type", 0))
    opt.loss_function_scale = config.get("loss_scale", 1.0)
    for k in ("refine_focal_length", "refine_principal_point", "refine_extra_params"):
        setattr(opt, k, bool(config.get(k, False)))
    opt.refine_extrinsics = bool(config.get("refine_extrinsics", True))
    adj = ba.BundleAdjuster.__new__(ba.BundleAdjuster)       # (no GPU context for the assembly)
    adj._options, adj._config = opt, cfg
    arrays, img_ids, pt_ids, cam_ids = adj._SetUp(rec)
    out = ["solve %d" % (1 if len(arrays.obs_image) else 0)]
    loss = config.get("loss_type", 0)
    scale = 0.0 if loss == 0 else config.get("loss_scale", 1.0)
    for o in range(len(arrays.obs_image)):
        ii = arrays.obs_image[o]
        l = arrays.obs_line[o]
        out.append("obs image %d point %d line %s %s %s pose_constant %d loss %d %s" % (
            img_ids[ii], pt_ids[arrays.obs_point[o]], float(l[0]).hex(), float(l[1]).hex(),
            float(l[2]).hex(), arrays.pose_flags[ii] & 1, loss, float(scale).hex()))
    for i, iid in enumerate(img_ids):
        f = int(arrays.pose_flags[i])
        out.append("image %d camera %d constant %d tvec_constant_mask %d quaternion 1" % (
            iid, cam_ids[arrays.image_camera[i]], f & 1, 0 if f & 1 else f >> 1))
    for p, pid in enumerate(pt_ids):
        out.append("point %d constant %d" % (pid, arrays.point_const[p]))
    for c, cid in enumerate(cam_ids):
        m = int(arrays.camera_model[c])
        foc, pp, ext = _GROUPS[m]
        var = 0
        if opt.refine_focal_length:
            var |= sum(1 << k for k in foc)
        if opt.refine_principal_point:
            var |= sum(1 << k for k in pp)
        if opt.refine_extra_params:
            var |= sum(1 << k for k in ext)
        if arrays.camera_const[c]:
            var = 0
        out.append("camera %d model %d variable_mask %d" % (cid, m, var))
    return "\n".join(sorted(out)) + "\n"


def _c_hex(text):
    """C's %a prints 0x0p+0 / 0x1.8p+1 where Python's float.hex prints 0x0.0p+0 / 0x1.8000000000000p+1."""
    import re
    return re.sub(r"-?0x[0-9a-f.]+p[+-]\d+", lambda m: float.fromhex(m.group(0)).hex(), text)


def test_python_mirror_assembles_like_the_reference(ref):
    cams = ([2, 4, 1], [[900.0, 500, 480, 0.05], [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003],
                        [1000.0, 1000.0, 500.0, 500.0]], np.array([0, 0, 1, 1, 2, 2, 0, 1], np.int32))
    scene, (img, pt) = _scene(cameras=cams)
    local = [0, 1]                        # cameras 1 only; cameras 2 and 3 enter from outside
    pts = sorted(set(int(p) for p in pt[(np.isin(img, local)) & (pt >= 0)]))
    for config in [dict(images=range(8), constant_poses=[0], constant_tvecs={1: [0]}),
                   dict(images=local, variable_points=pts, loss_type=1, loss_scale=1.0,
                        refine_focal_length=True, refine_extra_params=True),
                   dict(images=[2, 5, 7], constant_tvecs={5: [0, 2]}, constant_cameras=[1],
                        refine_principal_point=True),
                   dict(images=range(8), refine_extrinsics=False, loss_type=2, loss_scale=0.5)]:
        rc, a, _ = ref.ba_setup_compare(scene, config)
        assert rc == 0
        assert _c_hex(a) == _python_assembly_text(scene, config)


# ---------------------------------------------------------------------------------------------
# RefineAbsolutePoseFromLines (row A10): what the reference hands to Ceres, against what
# ppsfm_refine_absolute_pose_from_lines_ex builds (csrc/ba_host.cu)
# ---------------------------------------------------------------------------------------------
def test_pose_refinement_problem_of_the_reference(ref):
    """src/estimators/pose.cc:96-213 run against the recording ceres::Problem: one (2; 4, 3, 3, k)
    block per inlier in index order, Cauchy loss of the given scale, every point constant, the
    quaternion parameterised, the translation free, DENSE_QR on one thread, qvec normalised in
    place — and the variable intrinsics: none by default, otherwise the focal-length and / or
    extra-parameter groups, never the principal point.  The library builds the same problem
    (loss_type 2, point_const all 1, refine_principal_point 0, the group masks of kFocalMask /
    kExtraMask), which the last assertions tie to its source."""
    import os
    import re
    import test_ref_cost as TC
    sc = S.make_abs_pose_scene(n=400, inlier_ratio=0.5, seed=41)
    rng = np.random.default_rng(41)
    mask = (rng.random(400) < 0.6).astype(np.uint8)
    q = np.array([0.9, 0.1, -0.3, 0.2]) * 2.5                  # not normalised
    src = open(os.path.join(TC.ROOT, "privacy_preserving_sfm_b200", "csrc", "ba_host.cu")).read()
    body = src[src.index("int ppsfm_refine_absolute_pose_from_lines_ex("):]
    body = body[:body.index("\n}\n")]
    assert "o.loss_type = 2;" in body and "std::vector<uint8_t> pc(np, 1);" in body
    assert "refine_principal_point" not in body                # stays at its default 0
    assert re.search(r"o->refine_principal_point = 0;", src)
    focal, extra = TC._mask_table("privacy_preserving_sfm_b200/csrc/ba_host.cu", "kFocalMask"), \
        TC._mask_table("privacy_preserving_sfm_b200/csrc/ba_host.cu", "kExtraMask")
    for model, params in TC.MODELS.items():
        for rf, re_ in [(False, False), (True, False), (False, True), (True, True)]:
            r = ref.refine_absolute_pose_setup(sc["lines"], sc["points"], mask, model, params, q,
                                               sc["t"], rf, re_, gradient_tolerance=0.5,
                                               max_num_iterations=37, loss_scale=0.7)
            assert r["residual_blocks"] == int(mask.sum()) and r["uniform_blocks"]
            assert r["points_in_inlier_order"] and r["constant_points"] == int(mask.sum())
            assert (r["loss_kind"], r["loss_scale"]) == (2, 0.7)
            assert r["quaternion_parameterization"] and r["tvec_free"]
            assert (r["linear_solver_type"], r["num_threads"]) == (1, 1)       # DENSE_QR
            assert (r["gradient_tolerance"], r["max_num_iterations"]) == (0.5, 37)
            assert np.array_equal(r["qvec"], q / np.linalg.norm(q))
            want = (focal[model] if rf else 0) | (extra[model] if re_ else 0)
            assert r["camera_variable_mask"] == want, (model, rf, re_)
    # no inlier: nothing is added, nothing parameterised
    r = ref.refine_absolute_pose_setup(sc["lines"], sc["points"], np.zeros(400, np.uint8), 1,
                                       TC.MODELS[1], q, sc["t"])
    assert r["residual_blocks"] == 0 and not r["quaternion_parameterization"]
    assert np.array_equal(r["qvec"], q)
