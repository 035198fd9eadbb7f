"""GPU parity tests of intrinsics refinement in the bundle adjustment
(BundleAdjuster::ParameterizeCameras, src/optim/bundle_adjustment.cc:490-528; camera_params as the
fourth parameter block of the line cost functors, src/base/cost_functions.h:56-58, 130-141; the
refine_focal_length / refine_extra_params branch of RefineAbsolutePoseFromLines,
src/estimators/pose.cc:149-183) against the CPU oracle, which differentiates the same model
expressions with Jets.  Floating point: tolerances are stated per assertion."""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import bundle_adjustment as ba, synthetic as S

pytestmark = pytest.mark.gpu


def _scene(num_cams=8, num_points=300, obs=5, seed=3, **kw):
    return S.make_ba_scene(num_cams=num_cams, num_points=num_points, obs_per_point=obs, seed=seed,
                           **kw)


def _gauge_flags(n):
    f = np.zeros(n, np.uint8)
    f[0], f[1] = 1, 2
    return f


def _pair(oracle, sc, models, params, **kw):
    args = (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"], sc["obs_line"],
            models, params)
    return ba.BaArrays(*args, **kw), oracle.BaArrays(*args, **kw)


def _solve_both(ctx, oracle, a, b, **kw):
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(**kw))
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(num_threads=1, **kw))
    assert ok and ok2
    return s, s2


def _compare(a, b, s, s2, tol=1e-7):
    assert (s.num_successful_steps, s.num_unsuccessful_steps) == \
        (s2.num_successful_steps, s2.num_unsuccessful_steps)
    assert s.num_effective_parameters_reduced == s2.num_effective_parameters_reduced
    assert s.num_residuals_reduced == s2.num_residuals_reduced
    assert abs(s.initial_cost - s2.initial_cost) <= 1e-10 * s2.initial_cost
    # (refining the focal length lets this cost collapse — the residual is proportional to f —
    # so the final cost can be ~1e-18 of the initial one: tolerance relative to both)
    assert abs(s.final_cost - s2.final_cost) <= 1e-7 * s2.final_cost + 1e-13 * s2.initial_cost
    # parameters: relative to their own magnitude (focal ~1e3, distortion ~1e-2)
    scale = np.maximum(np.abs(b.camera_params), 1e-3)
    assert (np.abs(a.camera_params - b.camera_params) / scale).max() < 1e-6
    assert np.abs(a.qvecs - b.qvecs).max() < tol
    assert np.abs(a.tvecs - b.tvecs).max() < tol * max(1.0, np.abs(b.tvecs).max())
    assert np.abs(a.points - b.points).max() < 10 * tol


CASES = [
    # model, params, refine flags
    (2, [1000.0, 500, 500, 0.08], dict(refine_extra_params=1)),
    (3, [1000.0, 500, 500, 0.05, -0.01], dict(refine_extra_params=1)),
    (4, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003],
     dict(refine_focal_length=1, refine_extra_params=1)),
    (4, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003],
     dict(refine_focal_length=1, refine_principal_point=1, refine_extra_params=1)),
    (1, [1000.0, 990, 500, 480], dict(refine_focal_length=1)),
    (5, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003], dict(refine_extra_params=1)),
    (6, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.004, 0.02, -0.005, 0.001],
     dict(refine_focal_length=1, refine_extra_params=1)),
    (7, [1000.0, 990, 500, 480, 0.3], dict(refine_extra_params=1)),
    (8, [900.0, 500, 480, 0.05], dict(refine_extra_params=1)),
    (9, [900.0, 500, 480, 0.05, -0.01], dict(refine_focal_length=1, refine_extra_params=1)),
    (10, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.004, -0.002, 0.001, -0.001],
     dict(refine_extra_params=1)),
]


@pytest.mark.parametrize("model,params,refine", CASES)
def test_intrinsics_refinement_matches_oracle(ctx, oracle, model, params, refine):
    sc = _scene(noise_px=2.0)
    a, b = _pair(oracle, sc, [model], [params], pose_flags=_gauge_flags(8))
    s, s2 = _solve_both(ctx, oracle, a, b, max_num_iterations=8, gradient_tolerance=1e-6, **refine)
    _compare(a, b, s, s2)
    assert s.num_successful_steps >= 1
    p0 = np.zeros(12)
    p0[:len(params)] = params
    moved = a.camera_params[0] != p0
    assert moved.any()                                  # the variable group really moved ...
    if not refine.get("refine_principal_point"):        # ... and the constant groups did not
        pp_idx = [1, 2] if model in (0, 2, 3, 8, 9) else [2, 3]
        assert not moved[pp_idx].any()
    if not refine.get("refine_focal_length"):
        assert not moved[[0] if model in (0, 2, 3, 8, 9) else [0, 1]].any()


@pytest.mark.parametrize("loss", [1, 2])
def test_intrinsics_refinement_with_robust_loss(ctx, oracle, loss):
    sc = _scene(noise_px=2.0, seed=5)
    a, b = _pair(oracle, sc, [2], [[1000.0, 500, 500, 0.08]], pose_flags=_gauge_flags(8))
    s, s2 = _solve_both(ctx, oracle, a, b, max_num_iterations=8, gradient_tolerance=1e-6,
                        loss_type=loss, loss_scale=1.5, refine_extra_params=1)
    _compare(a, b, s, s2)


def test_two_cameras_one_constant(ctx, oracle):
    """config.SetConstantCamera (bundle_adjustment.cc:497): images alternate between two cameras,
    the second is constant; a point sees both."""
    sc = _scene(noise_px=2.0, seed=9)
    icam = np.arange(8) % 2
    params = [[1000.0, 500, 500, 0.08, 0, 0, 0, 0], [1000.0, 1000, 500, 500, 0.03, 0.0, 0.0, 0.0]]
    kw = dict(image_camera=icam, pose_flags=_gauge_flags(8))
    a, b = _pair(oracle, sc, [2, 4], params, camera_const=[0, 1], **kw)
    s, s2 = _solve_both(ctx, oracle, a, b, max_num_iterations=8, gradient_tolerance=1e-6,
                        refine_extra_params=1)
    _compare(a, b, s, s2)
    assert a.camera_params[0, 3] != 0.08
    assert np.array_equal(a.camera_params[1, :8], params[1])
    # both variable: two intrinsics blocks coupled through the points
    a, b = _pair(oracle, sc, [2, 4], params, **kw)
    s, s2 = _solve_both(ctx, oracle, a, b, max_num_iterations=8, gradient_tolerance=1e-6,
                        refine_extra_params=1)
    _compare(a, b, s, s2)
    assert a.camera_params[0, 3] != 0.08 and a.camera_params[1, 4] != 0.03


def test_every_image_its_own_camera(ctx, oracle):
    """Eight intrinsics blocks; every point sees five of them."""
    sc = _scene(noise_px=2.0, seed=13)
    params = [[1000.0 + 3 * i, 500, 500, 0.05] for i in range(8)]
    a, b = _pair(oracle, sc, [2] * 8, params, image_camera=np.arange(8),
                 pose_flags=_gauge_flags(8))
    s, s2 = _solve_both(ctx, oracle, a, b, max_num_iterations=6, gradient_tolerance=1e-6,
                        refine_extra_params=1)
    _compare(a, b, s, s2)


def test_principal_point_only_does_not_move(ctx, oracle):
    """The residual is a difference of two projections: the principal point cancels, its Jacobian
    columns are zero and the parameters stay where they are."""
    sc = _scene(noise_px=2.0)
    a, b = _pair(oracle, sc, [1], [[1000.0, 1000, 490, 510]], pose_flags=_gauge_flags(8))
    # (an ordinary BA in effect: converged after ~4 steps, later ones are rounding noise)
    s, s2 = _solve_both(ctx, oracle, a, b, max_num_iterations=4, gradient_tolerance=1e-6,
                        refine_principal_point=1)
    _compare(a, b, s, s2)
    assert np.array_equal(a.camera_params[0, :4], [1000.0, 1000, 490, 510])


def test_only_the_intrinsics_are_variable(ctx, oracle):
    """Constant poses and constant points: the residual blocks stay in the problem because the
    camera block is variable (Ceres drops only all-constant blocks); no pose block at all."""
    sc = _scene(noise_px=2.0)
    n_pts = sc["points"].shape[0]
    a, b = _pair(oracle, sc, [2], [[1000.0, 500, 500, 0.08]], pose_flags=np.ones(8, np.uint8),
                 point_const=np.ones(n_pts, np.uint8))
    # (the residual is linear in k: Gauss-Newton is there after one step, LM after two; a third
    # step would be accepted or rejected on rounding noise)
    s, s2 = _solve_both(ctx, oracle, a, b, max_num_iterations=2, refine_extra_params=1)
    _compare(a, b, s, s2)
    assert s.num_successful_steps == 2
    assert s.num_effective_parameters_reduced == 1
    assert s.num_residuals_reduced == 2 * len(sc["obs_cam"])
    assert a.camera_params[0, 3] != 0.08
    assert np.array_equal(a.points, sc["points"])
    # without the flag the same problem has no variable block at all
    a, b = _pair(oracle, sc, [2], [[1000.0, 500, 500, 0.08]], pose_flags=np.ones(8, np.uint8),
                 point_const=np.ones(n_pts, np.uint8))
    s, s2 = _solve_both(ctx, oracle, a, b, max_num_iterations=8)
    assert s.num_residuals_reduced == s2.num_residuals_reduced == 0


def test_constant_cameras_leave_the_solve_unchanged(ctx):
    """refine_* set but every camera in ConstantCameras(): the default solve (the point sums use
    atomics, so two runs agree to rounding, not bit for bit)."""
    sc = _scene(noise_px=1.0)
    args = (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"], sc["obs_line"],
            [2], [[1000.0, 500, 500, 0.08]])
    a = ba.BaArrays(*args, pose_flags=_gauge_flags(8), camera_const=[1])
    b = ba.BaArrays(*args, pose_flags=_gauge_flags(8))
    kw = dict(max_num_iterations=5, gradient_tolerance=1e-6)
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(
        refine_focal_length=1, refine_extra_params=1, **kw))
    ok2, s2 = ba.solve_arrays(ctx, b, ba.default_solver_options(**kw))
    assert ok and ok2 and abs(s.final_cost - s2.final_cost) <= 1e-9 * s2.final_cost
    assert s.num_effective_parameters_reduced == s2.num_effective_parameters_reduced
    assert s.num_residuals_reduced == s2.num_residuals_reduced
    # (two runs of the same solve differ by the order of their atomic point sums: rounding level
    # at the start, amplified by the iterations)
    assert np.abs(a.qvecs - b.qvecs).max() < 1e-8 and np.abs(a.points - b.points).max() < 1e-7
    assert np.array_equal(a.camera_params, b.camera_params)


def test_resident_problem_reset_restores_the_intrinsics(ctx):
    sc = _scene(noise_px=2.0)
    a = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                    sc["obs_line"], [2], [[1000.0, 500, 500, 0.08]], pose_flags=_gauge_flags(8))
    o = ba.default_solver_options(max_num_iterations=6, refine_extra_params=1)
    rp = ba.ResidentProblem(ctx, a, o)
    ok, s1 = rp.run()
    rp.download()
    k1 = a.camera_params[0, 3]
    rp.reset()
    ok, s2 = rp.run()
    rp.download()
    rp.free()
    assert k1 != 0.08 and abs(a.camera_params[0, 3] - k1) <= 1e-6 * abs(k1)
    assert s1.initial_cost == s2.initial_cost      # the run after reset starts from k = 0.08 again
    assert abs(s1.final_cost - s2.final_cost) <= 1e-8 * s1.final_cost


def test_pose_refinement_with_focal_length(ctx, oracle):
    """RefineAbsolutePoseFromLines with refine_focal_length / refine_extra_params
    (pose.cc:149-183): one image, constant points, Cauchy loss; the principal point stays fixed.
    The oracle solves the same problem through its bundle-adjustment entry point."""
    sc = S.make_abs_pose_scene(n=2000, inlier_ratio=0.6, seed=71)
    rng = np.random.default_rng(1)
    q0 = S.rotmat_to_quat(sc["R"]) + 0.002 * rng.normal(size=4)
    t0 = sc["t"] + 0.01 * rng.normal(size=3)
    mask = sc["is_inlier"].astype(np.uint8)
    for model, prm, ff, fe in [(2, [1010.0, 500.0, 500.0, 0.01], True, True),
                               (2, [1010.0, 500.0, 500.0, 0.01], False, True),
                               (1, [1010.0, 995.0, 500.0, 500.0], True, False)]:
        cam = ba.Camera(1, model, prm)
        opt = ba.AbsolutePoseRefinementOptions()
        opt.refine_focal_length, opt.refine_extra_params = ff, fe
        q, t = q0.copy(), t0.copy()
        ok = ba.RefineAbsolutePoseFromLines(opt, mask, sc["lines"], sc["points"], q, t, cam,
                                            ctx=ctx)
        s = ba.RefineAbsolutePoseFromLines.last_summary
        idx = np.flatnonzero(mask)
        b = oracle.BaArrays([q0 / np.linalg.norm(q0)], [t0], sc["points"][idx],
                            np.zeros(len(idx), np.int32), np.arange(len(idx), dtype=np.int32),
                            sc["lines"][idx], [model], [prm],
                            point_const=np.ones(len(idx), np.uint8))
        ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(
            num_threads=1, loss_type=2, loss_scale=1.0, gradient_tolerance=1.0,
            max_num_iterations=100, function_tolerance=1e-6, parameter_tolerance=1e-8,
            refine_focal_length=int(ff), refine_extra_params=int(fe)))
        assert ok and ok2
        assert (s.num_successful_steps, s.num_unsuccessful_steps) == \
            (s2.num_successful_steps, s2.num_unsuccessful_steps)
        assert np.abs(q - b.qvecs[0]).max() < 1e-8 and np.abs(t - b.tvecs[0]).max() < 1e-8
        n = len(prm)
        assert np.abs(cam.params - b.camera_params[0, :n]).max() < 1e-6 * 1000
        assert not np.array_equal(cam.params, prm)
        pp = [1, 2] if model == 2 else [2, 3]
        assert np.array_equal(cam.params[pp], np.array(prm)[pp])   # "always fixed" (pose.cc:152)


def test_bundle_adjuster_object_with_refine_extra_params(ctx):
    sc = _scene(num_cams=6, num_points=200, obs=4, seed=41, noise_px=2.0)
    rec = ba.Reconstruction()
    rec.cameras[1] = ba.Camera(1, "SIMPLE_RADIAL", [1000.0, 500.0, 500.0, 0.08])
    rec.cameras[2] = ba.Camera(2, "SIMPLE_RADIAL", [1000.0, 500.0, 500.0, 0.08])
    for i in range(6):
        rec.images[i + 1] = ba.Image(i + 1, 1 + i % 2, sc["qvecs"][i], sc["tvecs"][i])
    for p in range(sc["points"].shape[0]):
        rec.points3D[p + 100] = ba.Point3D(sc["points"][p])
    for k in range(len(sc["obs_cam"])):
        img = rec.images[sc["obs_cam"][k] + 1]
        pid = sc["obs_pt"][k] + 100
        img.lines.append(ba.FeatureLine(sc["obs_line"][k], False, pid))
        rec.points3D[pid].track.append((img.image_id, len(img.lines) - 1))
    cfg = ba.BundleAdjustmentConfig()
    for i in range(6):
        cfg.AddImage(i + 1)
    cfg.SetConstantPose(1)
    cfg.SetConstantTvec(2, [0])
    cfg.SetConstantCamera(2)
    opt = ba.BundleAdjustmentOptions()
    opt.refine_extra_params = True
    opt.print_summary = False
    opt.solver_options.max_num_iterations = 8
    adj = ba.BundleAdjuster(opt, cfg, ctx=ctx)
    assert adj.Solve(rec)
    assert rec.cameras[1].params[3] != 0.08 and rec.cameras[1].params[0] == 1000.0
    assert np.array_equal(rec.cameras[2].params, [1000.0, 500.0, 500.0, 0.08])
    assert adj.Summary().final_cost < adj.Summary().initial_cost
