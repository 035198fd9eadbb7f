"""GPU parity tests of the line-reprojection bundle adjustment (CUDA through the C-ABI vs the CPU
oracle).  Floating-point path: tolerances are stated per assertion; the north-star bar is poses
within 1e-6 rad / 1e-6 relative translation of the reference path."""
import numpy as np
import pytest

import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import bundle_adjustment as ba, synthetic as S

pytestmark = pytest.mark.gpu

MODELS = [(0, [900.0, 500, 480]), (1, [1000.0, 990, 500, 480]), (2, [900.0, 500, 480, 0.05]),
          (3, [900.0, 500, 480, 0.05, -0.01]),
          (4, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003]),
          (5, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003]),                  # OPENCV_FISHEYE
          (6, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.004, 0.02, -0.005, 0.001]),  # FULL_OPENCV
          (7, [1000.0, 990, 500, 480, 0.3]),                                         # FOV
          (7, [1000.0, 990, 500, 480, 0.003]),                                       # FOV, small omega
          (8, [900.0, 500, 480, 0.05]),                                              # SIMPLE_RADIAL_FISHEYE
          (9, [900.0, 500, 480, 0.05, -0.01]),                                       # RADIAL_FISHEYE
          (10, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003, 0.004, -0.002, 0.001, -0.001])]  # THIN_PRISM_FISHEYE


def _scene(num_cams=8, num_points=300, obs=5, seed=3, **kw):
    return S.make_ba_scene(num_cams=num_cams, num_points=num_points, obs_per_point=obs, seed=seed,
                           **kw)


def _both(oracle, sc, model=1, params=None, pose_flags=None, point_const=None):
    params = params if params is not None else sc["cam_params"]
    args = (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"], sc["obs_line"],
            [model], [params])
    kw = dict(pose_flags=pose_flags, point_const=point_const)
    return ba.BaArrays(*args, **kw), oracle.BaArrays(*args, **kw)


def _gauge_flags(n):
    f = np.zeros(n, np.uint8)
    f[0] = 1        # image 0: constant pose   (sfm/incremental_mapper.cc:907-926)
    f[1] = 2        # image 1: constant tvec[0]
    return f


@pytest.mark.parametrize("n", [1, 5, 16, 17, 63, 64, 65, 127, 128, 130, 192, 600, 1000, 2994])
def test_dense_cholesky_solve(ctx, n):
    rng = np.random.default_rng(n)
    M = rng.normal(size=(n, n))
    A = M @ M.T + n * np.eye(n)
    b = rng.normal(size=n)
    ok, x = ba.dense_cholesky_solve(ctx, A, b)
    assert ok
    xr = np.linalg.solve(A, b)
    assert np.abs(x - xr).max() <= 1e-10 * max(1.0, np.abs(xr).max())


def test_dense_cholesky_rejects_indefinite(ctx):
    A = np.eye(70)
    A[40, 40] = -1.0
    ok, _ = ba.dense_cholesky_solve(ctx, A, np.ones(70))
    assert not ok


@pytest.mark.parametrize("model,params", MODELS)
@pytest.mark.parametrize("loss", [0, 1, 2])
def test_linearize_matches_autodiff_oracle(ctx, oracle, model, params, loss):
    sc = _scene(num_cams=5, num_points=60, obs=3, seed=11)
    a, _ = _both(oracle, sc, model, params)
    o = ba.default_solver_options(loss_type=loss, loss_scale=1.5)
    r, jc, jp, cost = ba.linearize_arrays(ctx, a, o)
    cost_ref = 0.0
    for k in range(len(sc["obs_cam"])):
        ci, pi = sc["obs_cam"][k], sc["obs_pt"][k]
        q = sc["qvecs"][ci] / np.linalg.norm(sc["qvecs"][ci])
        rr, jcr, jpr = oracle.line_cost_tangent(model, params, sc["obs_line"][k], q,
                                                sc["tvecs"][ci], sc["points"][pi])
        s = rr @ rr
        if loss == 0:
            rho0, rho1 = s, 1.0
        elif loss == 1:
            b = 1.5 ** 2
            rho0, rho1 = 2 * b * (np.sqrt(1 + s / b) - 1), 1 / np.sqrt(1 + s / b)
        else:
            b = 1.5 ** 2
            rho0, rho1 = b * np.log(1 + s / b), 1 / (1 + s / b)
        cost_ref += 0.5 * rho0
        sr = np.sqrt(rho1)
        assert np.allclose(r[k], sr * rr, rtol=1e-10, atol=1e-9)
        assert np.allclose(jc[k], sr * jcr, rtol=1e-8, atol=1e-7)   # |J| ~ 1e3
        assert np.allclose(jp[k], sr * jpr, rtol=1e-8, atol=1e-7)
    assert abs(cost - cost_ref) <= 1e-10 * cost_ref


def test_linearize_masks_constant_blocks(ctx, oracle):
    sc = _scene(num_cams=4, num_points=30, obs=3, seed=12)
    flags = np.array([1, 2 | 8, 0, 0], np.uint8)       # cam0 const, cam1 tvec[0], tvec[2] const
    pc = np.zeros(30, np.uint8)
    pc[:5] = 1
    a, _ = _both(oracle, sc, pose_flags=flags, point_const=pc)
    r, jc, jp, _ = ba.linearize_arrays(ctx, a, ba.default_solver_options())
    oc, op = sc["obs_cam"], sc["obs_pt"]
    assert np.all(jc[oc == 0] == 0)
    assert np.all(jc[oc == 1][:, :, 3] == 0) and np.all(jc[oc == 1][:, :, 5] == 0)
    assert np.any(jc[oc == 1][:, :, 4] != 0) and np.any(jc[oc == 1][:, :, :3] != 0)
    assert np.all(jp[pc[op] == 1] == 0)
    # blocks with nothing variable are dropped entirely (Ceres removes them from the program)
    dropped = (oc == 0) & (pc[op] == 1)
    assert np.all(r[dropped] == 0)
    assert np.all(np.abs(r[~dropped]).sum(axis=1) > 0)


def _compare_solutions(a, b, s, s2, tol_pose=1e-8, tol_cost=1e-9):
    assert s.termination_type == s2.termination_type
    # same accepted / rejected step sequence — up to the point where a decision is made on
    # rounding noise (cost change below 1e-11 relative), beyond which the two need not agree
    acc, acc2 = list(s.trace_accepted[:s.trace_len]), list(s2.trace_accepted[:s2.trace_len])
    for i in range(min(len(acc), len(acc2))):
        if acc[i] != acc2[i]:
            moved = max(abs(s.trace_cost[i] - s.trace_cost[i - 1]),
                        abs(s2.trace_cost[i] - s2.trace_cost[i - 1]))
            assert i > 0 and moved <= 1e-11 * s2.trace_cost[i - 1], (i, acc, acc2)
            break
    else:
        assert len(acc) == len(acc2)
        assert (s.num_successful_steps, s.num_unsuccessful_steps) == \
            (s2.num_successful_steps, s2.num_unsuccessful_steps)
    assert abs(s.initial_cost - s2.initial_cost) <= 1e-11 * s2.initial_cost
    assert abs(s.final_cost - s2.final_cost) <= tol_cost * max(s2.final_cost, 1e-300)
    # north-star bar: 1e-6 rad / 1e-6 relative translation; observed agreement is far tighter
    assert np.abs(a.qvecs - b.qvecs).max() < tol_pose
    assert np.abs(a.tvecs - b.tvecs).max() < tol_pose * max(1.0, np.abs(b.tvecs).max())
    assert np.abs(a.points - b.points).max() < tol_pose * 10


@pytest.mark.parametrize("loss,scale", [(0, 1.0), (1, 1.0), (2, 2.0)])
def test_ba_solve_matches_oracle(ctx, oracle, loss, scale):
    sc = _scene(num_cams=10, num_points=400, obs=5, seed=21)
    a, b = _both(oracle, sc, pose_flags=_gauge_flags(10))
    kw = dict(loss_type=loss, loss_scale=scale, max_num_iterations=30, gradient_tolerance=1e-2)
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(**kw))
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(num_threads=1, **kw))
    assert ok and ok2
    assert s.final_cost < 0.05 * s.initial_cost
    _compare_solutions(a, b, s, s2)
    # gauge: constant pose / constant tvec[0] really stayed put
    assert np.array_equal(a.qvecs[0], sc["qvecs"][0] / np.linalg.norm(sc["qvecs"][0])) or \
        np.allclose(a.qvecs[0], sc["qvecs"][0], atol=1e-15)
    assert np.array_equal(a.tvecs[0], sc["tvecs"][0])
    assert a.tvecs[1, 0] == sc["tvecs"][1, 0]


@pytest.mark.parametrize("model,params", MODELS)
def test_ba_solve_camera_models(ctx, oracle, model, params):
    sc = _scene(num_cams=6, num_points=150, obs=4, seed=31)
    a, b = _both(oracle, sc, model, params, pose_flags=_gauge_flags(6))
    kw = dict(max_num_iterations=15, gradient_tolerance=1e-4)   # stop at convergence: beyond it
    # accept/reject decisions are made on rounding noise and need not agree
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(**kw))
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(num_threads=1, **kw))
    assert ok and ok2
    _compare_solutions(a, b, s, s2, tol_pose=1e-7, tol_cost=1e-8)


def test_ba_local_bundle_configuration(ctx, oracle):
    """Local BA shape (sfm/incremental_mapper.cc:781-891): few variable images, observations from
    out-of-bundle images through constant poses, some constant points, SOFT_L1."""
    sc = _scene(num_cams=12, num_points=300, obs=6, seed=41)
    flags = np.ones(12, np.uint8)
    flags[[2, 3, 4, 5]] = 0
    flags[6] = 0
    flags[5] = 2                        # second-to-last: tvec[0] constant
    pc = (np.arange(300) % 7 == 0).astype(np.uint8)
    a, b = _both(oracle, sc, pose_flags=flags, point_const=pc)
    kw = dict(loss_type=1, loss_scale=1.0, max_num_iterations=25, gradient_tolerance=10.0)
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(**kw))
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(num_threads=1, **kw))
    assert ok and ok2
    _compare_solutions(a, b, s, s2)
    const_imgs = np.where(flags & 1)[0]
    assert np.array_equal(a.qvecs[const_imgs], sc["qvecs"][const_imgs])
    assert np.array_equal(a.points[pc == 1], sc["points"][pc == 1])
    assert s.num_residuals_reduced == s2.num_residuals_reduced < s.num_residuals


def test_ba_noise_free_scene_is_a_fixed_point(ctx):
    sc = _scene(num_cams=8, num_points=300, obs=5, seed=51, noise_px=0.0)
    a = ba.BaArrays(sc["qvecs_gt"], sc["tvecs_gt"], sc["points_gt"], sc["obs_cam"], sc["obs_pt"],
                    sc["obs_line"], [1], [sc["cam_params"]], pose_flags=_gauge_flags(8))
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(max_num_iterations=5,
                                                             gradient_tolerance=1e-6))
    assert ok and s.initial_cost < 1e-12 and s.num_successful_steps + s.num_unsuccessful_steps == 0
    assert np.allclose(a.points, sc["points_gt"], atol=1e-12)


def test_ba_recovers_ground_truth_without_noise(ctx):
    sc = _scene(num_cams=8, num_points=400, obs=6, seed=52, noise_px=0.0)
    a = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                    sc["obs_line"], [1], [sc["cam_params"]], pose_flags=_gauge_flags(8))
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(max_num_iterations=50,
                                                             gradient_tolerance=1e-9))
    assert ok and s.final_cost < 1e-14
    # gauge fixed by cam 0 + tvec[0] of cam 1 at their true values -> unique solution = truth
    sign = np.sign((a.qvecs * sc["qvecs_gt"]).sum(axis=1))[:, None]
    assert np.abs(a.qvecs * sign - sc["qvecs_gt"]).max() < 1e-7
    assert np.abs(a.tvecs - sc["tvecs_gt"]).max() < 1e-6
    assert np.abs(a.points - sc["points_gt"]).max() < 1e-6


def test_ba_empty_and_degenerate_problems(ctx):
    sc = _scene(num_cams=3, num_points=10, obs=2, seed=61)
    a = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"][:0], sc["obs_pt"][:0],
                    sc["obs_line"][:0], [1], [sc["cam_params"]])
    ok, _ = ba.solve_arrays(ctx, a, ba.default_solver_options())
    assert not ok                                    # zero residuals -> false (:269-271)
    # everything constant: nothing to optimise, state untouched
    a = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                    sc["obs_line"], [1], [sc["cam_params"]], pose_flags=np.ones(3, np.uint8),
                    point_const=np.ones(10, np.uint8))
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options())
    assert ok and s.num_residuals_reduced == 0
    assert np.array_equal(a.points, sc["points"]) and np.array_equal(a.tvecs, sc["tvecs"])
    with pytest.raises(pp.PpsfmError):               # line normal must be unit length
        bad = sc["obs_line"].copy()
        bad[0, :2] *= 2
        ba.solve_arrays(ctx, ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"],
                                         sc["obs_pt"], bad, [1], [sc["cam_params"]]),
                        ba.default_solver_options())


def test_refine_absolute_pose_matches_oracle(ctx, oracle):
    sc = S.make_abs_pose_scene(n=3000, inlier_ratio=0.5, seed=71)
    cam = ba.Camera(1, "PINHOLE", [1000.0, 1000.0, 500.0, 500.0])
    # start from a perturbed pose, refine on the true inliers
    rng = np.random.default_rng(1)
    q0 = S.rotmat_to_quat(sc["R"]) + 0.002 * rng.normal(size=4)
    t0 = sc["t"] + 0.01 * rng.normal(size=3)
    mask = sc["is_inlier"].astype(np.uint8)
    q, t = q0.copy(), t0.copy()
    ok = ba.RefineAbsolutePoseFromLines(ba.AbsolutePoseRefinementOptions(), mask, sc["lines"],
                                        sc["points"], q, t, cam, ctx=ctx)
    ok2, q2, t2, s2 = oracle.refine_absolute_pose(sc["lines"], sc["points"], mask, 1, cam.params,
                                                  q0, t0)
    s = ba.RefineAbsolutePoseFromLines.last_summary
    assert ok and ok2
    assert (s.num_successful_steps, s.num_unsuccessful_steps) == \
        (s2.num_successful_steps, s2.num_unsuccessful_steps)
    assert np.abs(q - q2).max() < 1e-9 and np.abs(t - t2).max() < 1e-9
    Rest = S.quat_to_rotmat(q / np.linalg.norm(q))
    ang = np.degrees(np.arccos(np.clip((np.trace(Rest.T @ sc["R"]) - 1) / 2, -1, 1)))
    assert ang < 0.02 and np.linalg.norm(t - sc["t"]) < 2e-3
    # no inliers: nothing to do, still "usable"
    q3, t3 = q0.copy(), t0.copy()
    assert ba.RefineAbsolutePoseFromLines(ba.AbsolutePoseRefinementOptions(), np.zeros_like(mask),
                                          sc["lines"], sc["points"], q3, t3, cam, ctx=ctx)
    assert np.array_equal(q3, q0) and np.array_equal(t3, t0)


def _reconstruction(sc):
    rec = ba.Reconstruction()
    rec.cameras[1] = ba.Camera(1, "PINHOLE", sc["cam_params"])
    n_img = sc["qvecs"].shape[0]
    for i in range(n_img):
        rec.images[i + 1] = ba.Image(i + 1, 1, sc["qvecs"][i], sc["tvecs"][i])
    for p in range(sc["points"].shape[0]):
        rec.points3D[p + 100] = ba.Point3D(sc["points"][p])
    for k in range(len(sc["obs_cam"])):
        img = rec.images[sc["obs_cam"][k] + 1]
        pid = sc["obs_pt"][k] + 100
        img.lines.append(ba.FeatureLine(sc["obs_line"][k], False, pid))
        rec.points3D[pid].track.append((img.image_id, len(img.lines) - 1))
    return rec


def test_bundle_adjuster_object_api(ctx, oracle):
    """BundleAdjuster(options, config).Solve(reconstruction) as IncrementalMapper::AdjustGlobalBundle
    uses it (sfm/incremental_mapper.cc:893-939)."""
    sc = _scene(num_cams=7, num_points=200, obs=4, seed=81)
    rec = _reconstruction(sc)
    config = ba.BundleAdjustmentConfig()
    for iid in rec.images:
        config.AddImage(iid)
    config.SetConstantPose(1)
    config.SetConstantTvec(2, [0])
    options = ba.BundleAdjustmentOptions()
    options.solver_options.max_num_iterations = 20
    options.solver_options.gradient_tolerance = 1e-4
    options.print_summary = False
    adj = ba.BundleAdjuster(options, config, ctx=ctx)
    assert adj.Solve(rec)
    with pytest.raises(pp.PpsfmError):
        adj.Solve(rec)                                # single use (:262)
    _, b = _both(oracle, sc, pose_flags=_gauge_flags(7))
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(num_threads=1, max_num_iterations=20,
                                                             gradient_tolerance=1e-4))
    s = adj.Summary()
    assert abs(s.final_cost - s2.final_cost) <= 1e-9 * s2.final_cost
    for i in range(7):
        assert np.abs(rec.images[i + 1].qvec - b.qvecs[i]).max() < 1e-8
        assert np.abs(rec.images[i + 1].tvec - b.tvecs[i]).max() < 1e-8
    assert np.abs(rec.points3D[100].xyz - b.points[0]).max() < 1e-7


def test_bundle_adjuster_out_of_set_observations(ctx):
    """AddPointToProblem (:437-488): variable points keep their observations from images outside
    the image set through constant poses; points with partial tracks become constant (:530-542)."""
    sc = _scene(num_cams=6, num_points=120, obs=4, seed=82)
    rec = _reconstruction(sc)
    config = ba.BundleAdjustmentConfig()
    for iid in (1, 2, 3):
        config.AddImage(iid)
    config.SetConstantPose(1)
    for pid in range(100, 160):
        config.AddVariablePoint(pid)
    options = ba.BundleAdjustmentOptions()
    options.solver_options.max_num_iterations = 10
    options.print_summary = False
    before = {iid: (img.qvec.copy(), img.tvec.copy()) for iid, img in rec.images.items()}
    pts_before = {pid: p.xyz.copy() for pid, p in rec.points3D.items()}
    adj = ba.BundleAdjuster(options, config, ctx=ctx)
    assert adj.Solve(rec)
    s = adj.Summary()
    assert s.final_cost < s.initial_cost
    for iid in (1, 4, 5, 6):                          # constant / out-of-set images untouched
        assert np.array_equal(rec.images[iid].tvec, before[iid][1])
    assert any(not np.array_equal(rec.images[i].tvec, before[i][1]) for i in (2, 3))
    moved = [pid for pid in pts_before if not np.array_equal(rec.points3D[pid].xyz, pts_before[pid])]
    assert moved and all(100 <= pid < 160 for pid in moved)   # only configured points move


def test_ba_solve_is_reproducible_run_to_run(ctx):
    """No floating-point atomics whose order matters on the default path (segmented reductions,
    ordered partial sums, a dataflow factorisation with a fixed summation order): two solves of
    the same problem return the same bits (tracks of <= 32 observations)."""
    sc = _scene(num_cams=12, num_points=3000, obs=6, seed=77)
    outs = []
    for _ in range(3):
        a, _unused = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                                 sc["obs_line"], [1], [sc["cam_params"]],
                                 pose_flags=_gauge_flags(12)), None
        ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(max_num_iterations=12,
                                                                  loss_type=1, loss_scale=1.0))
        assert ok
        outs.append((a.qvecs.copy(), a.tvecs.copy(), a.points.copy(), s.final_cost,
                     s.num_successful_steps, s.num_unsuccessful_steps))
    for o in outs[1:]:
        assert o[3] == outs[0][3] and o[4:] == outs[0][4:]
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1])
        assert np.array_equal(o[2], outs[0][2])
