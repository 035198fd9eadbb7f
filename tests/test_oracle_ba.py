"""CPU tests of the BA oracle: autodiff vs finite differences, LM vs scipy.optimize.least_squares,
gauge / constant-block handling.  (Ceres boundary: parity unpinned — these are the independent
cross-checks SURVEY.md §8(c) lists.)"""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import synthetic as S

MODELS = [(0, [900.0, 500, 480]), (1, [1000.0, 990, 500, 480]), (2, [900.0, 500, 480, 0.05]),
          (3, [900.0, 500, 480, 0.05, -0.01]),
          (4, [1000.0, 990, 500, 480, 0.05, -0.01, 0.002, -0.003])]


@pytest.mark.parametrize("model,params", MODELS)
def test_autodiff_matches_finite_differences(oracle, model, params):
    sc = S.make_ba_scene(num_cams=4, num_points=20, obs_per_point=3, seed=1)
    q, t, X, line = sc["qvecs"][2], sc["tvecs"][2], sc["points"][7], sc["obs_line"][5]
    r, jc, jx = oracle.line_cost_tangent(model, params, line, q, t, X)
    eps = 1e-6
    num, numx = np.zeros((2, 6)), np.zeros((2, 3))
    for k in range(3):
        d = np.zeros(3)
        d[k] = eps
        f = lambda qq, tt, xx: oracle.line_cost(model, params, line, qq, tt, xx)[0]
        num[:, k] = (f(oracle.quaternion_plus(q, d), t, X) - f(oracle.quaternion_plus(q, -d), t, X)) / (2 * eps)
        num[:, 3 + k] = (f(q, t + d, X) - f(q, t - d, X)) / (2 * eps)
        numx[:, k] = (f(q, t, X + d) - f(q, t, X - d)) / (2 * eps)
    assert np.allclose(jc, num, rtol=1e-6, atol=1e-5)
    assert np.allclose(jx, numx, rtol=1e-6, atol=1e-5)


def test_pinhole_residual_closed_form(oracle):
    # r = alpha * (fx a, fy b) for PINHOLE
    sc = S.make_ba_scene(num_cams=3, num_points=10, obs_per_point=2, seed=2)
    q, t, X, line = sc["qvecs"][1], sc["tvecs"][1], sc["points"][3], sc["obs_line"][0]
    r = oracle.line_cost(1, [1000.0, 900.0, 500, 480], line, q, t, X)[0]
    p = S.quat_to_rotmat(q) @ X + t
    alpha = line[0] * p[0] / p[2] + line[1] * p[1] / p[2] + line[2]
    assert np.allclose(r, alpha * np.array([1000.0 * line[0], 900.0 * line[1]]), rtol=1e-9, atol=1e-9)


def test_quaternion_plus(oracle):
    q = np.array([0.5, -0.5, 0.5, 0.5])
    assert np.array_equal(oracle.quaternion_plus(q, np.zeros(3)), q)
    d = np.array([0.01, -0.02, 0.03])
    qp = oracle.quaternion_plus(q, d)
    assert abs(np.linalg.norm(qp) - 1) < 1e-15
    n = np.linalg.norm(d)
    qd = np.concatenate([[np.cos(n)], np.sin(n) / n * d])
    assert np.allclose(qp, S.quat_mul(qd, q), atol=1e-15)


def _flags(n):
    f = np.zeros(n, np.uint8)
    f[0], f[1] = 1, 2
    return f


def test_lm_reaches_the_scipy_minimum(oracle):
    from scipy.optimize import least_squares
    sc = S.make_ba_scene(num_cams=5, num_points=60, obs_per_point=4, seed=3)
    flags = _flags(5)
    a = oracle.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                        sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags)
    ok, s = oracle.ba_solve(a, oracle.ba_default_options(max_num_iterations=50,
                                                         gradient_tolerance=1e-5, num_threads=1))
    assert ok and s.termination_type in (0, 1) and s.final_gradient_max_norm < 1e-2

    q0 = sc["qvecs"] / np.linalg.norm(sc["qvecs"], axis=1, keepdims=True)

    def unpack(x):
        q, t, X = q0.copy(), sc["tvecs"].copy(), x[-180:].reshape(60, 3)
        k = 0
        for i in range(1, 5):
            q[i] = oracle.quaternion_plus(q0[i], x[k:k + 3])
            k += 3
            if i == 1:
                t[i, 1:] = sc["tvecs"][i, 1:] + x[k:k + 2]
                k += 2
            else:
                t[i] = sc["tvecs"][i] + x[k:k + 3]
                k += 3
        return q, t, X

    def fun(x):
        q, t, X = unpack(x)
        out = []
        for o in range(len(sc["obs_cam"])):
            ci, pi = sc["obs_cam"][o], sc["obs_pt"][o]
            out.append(oracle.line_cost(1, sc["cam_params"], sc["obs_line"][o], q[ci], t[ci], X[pi])[0])
        return np.concatenate(out)

    x0 = np.concatenate([np.zeros(3 * 4 + 2 + 3 * 3), sc["points"].reshape(-1)])
    sol = least_squares(fun, x0, method="lm", xtol=1e-14, ftol=1e-14, gtol=1e-14)
    assert abs(0.5 * np.sum(sol.fun ** 2) - s.final_cost) <= 1e-8 * s.final_cost
    q, t, X = unpack(sol.x)
    assert np.abs(q - a.qvecs).max() < 1e-6 and np.abs(t - a.tvecs).max() < 1e-6
    assert np.abs(X - a.points).max() < 1e-5


@pytest.mark.parametrize("loss", [1, 2])
def test_robust_losses_against_scipy_cost(oracle, loss):
    sc = S.make_ba_scene(num_cams=4, num_points=40, obs_per_point=3, seed=4)
    a = oracle.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                        sc["obs_line"], [1], [sc["cam_params"]], pose_flags=_flags(4))
    o = oracle.ba_default_options(loss_type=loss, loss_scale=2.0, num_threads=1)
    c = oracle.ba_cost(a, o)
    tot = 0.0
    for k in range(len(sc["obs_cam"])):
        ci, pi = sc["obs_cam"][k], sc["obs_pt"][k]
        q = sc["qvecs"][ci] / np.linalg.norm(sc["qvecs"][ci])
        r = oracle.line_cost(1, sc["cam_params"], sc["obs_line"][k], q, sc["tvecs"][ci],
                             sc["points"][pi])[0]
        s = r @ r
        b = 4.0
        tot += 0.5 * (2 * b * (np.sqrt(1 + s / b) - 1) if loss == 1 else b * np.log(1 + s / b))
    assert abs(c - tot) <= 1e-12 * tot
    ok, summ = oracle.ba_solve(a, oracle.ba_default_options(loss_type=loss, loss_scale=2.0,
                                                            max_num_iterations=30, num_threads=1,
                                                            gradient_tolerance=1e-6))
    assert ok and summ.final_cost < 0.1 * summ.initial_cost


def test_gauge_and_constant_blocks(oracle):
    sc = S.make_ba_scene(num_cams=6, num_points=80, obs_per_point=4, seed=5)
    flags = _flags(6)
    flags[5] = 1
    pc = np.zeros(80, np.uint8)
    pc[::5] = 1
    a = oracle.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                        sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags, point_const=pc)
    ok, s = oracle.ba_solve(a, oracle.ba_default_options(max_num_iterations=20, num_threads=1,
                                                         gradient_tolerance=1e-6))
    assert ok
    assert np.array_equal(a.tvecs[0], sc["tvecs"][0]) and np.array_equal(a.tvecs[5], sc["tvecs"][5])
    assert a.tvecs[1, 0] == sc["tvecs"][1, 0] and a.tvecs[1, 1] != sc["tvecs"][1, 1]
    assert np.array_equal(a.points[pc == 1], sc["points"][pc == 1])
    assert not np.array_equal(a.points[pc == 0], sc["points"][pc == 0])
    # residual blocks with nothing variable are dropped
    n_drop = int(((flags[sc["obs_cam"]] & 1) & pc[sc["obs_pt"]]).sum())
    assert s.num_residuals - s.num_residuals_reduced == 2 * n_drop
    # multi-threaded run reaches the same minimum
    b = oracle.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                        sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags, point_const=pc)
    ok, s2 = oracle.ba_solve(b, oracle.ba_default_options(max_num_iterations=20, num_threads=4,
                                                          gradient_tolerance=1e-6))
    # (thread-order dependent summation: the runs may stop one LM step apart at the 1e-6 gradient
    # tolerance, so compare at the accuracy that tolerance implies, not at rounding level)
    assert abs(s2.final_cost - s.final_cost) <= 1e-8 * s.final_cost
    assert np.abs(a.points - b.points).max() < 1e-6


def test_no_residuals_returns_false(oracle):
    sc = S.make_ba_scene(num_cams=3, num_points=10, obs_per_point=2, seed=6)
    a = oracle.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"][:0], sc["obs_pt"][:0],
                        sc["obs_line"][:0], [1], [sc["cam_params"]])
    ok, _ = oracle.ba_solve(a, oracle.ba_default_options(num_threads=1))
    assert not ok


def test_pose_refinement_improves_pose(oracle):
    sc = S.make_abs_pose_scene(n=2000, inlier_ratio=0.6, seed=7)
    rng = np.random.default_rng(0)
    q0 = S.rotmat_to_quat(sc["R"]) + 0.002 * rng.normal(size=4)
    t0 = sc["t"] + 0.01 * rng.normal(size=3)
    mask = sc["is_inlier"].astype(np.uint8)
    ok, q, t, s = oracle.refine_absolute_pose(sc["lines"], sc["points"], mask, 1,
                                              [1000.0, 1000.0, 500, 500], q0, t0)
    assert ok and s.final_cost < 0.5 * s.initial_cost
    ang = lambda qq: np.degrees(np.arccos(np.clip(
        (np.trace(S.quat_to_rotmat(qq / np.linalg.norm(qq)).T @ sc["R"]) - 1) / 2, -1, 1)))
    assert ang(q) < 0.1 * ang(q0)
    assert np.linalg.norm(t - sc["t"]) < 0.2 * np.linalg.norm(t0 - sc["t"])


def test_oracle_intrinsics_refinement(oracle):
    """ParameterizeCameras (src/optim/bundle_adjustment.cc:490-528) in the oracle: the camera
    parameters are a fourth parameter block (cost_functions.h:56-58) whose groups (focal /
    principal point / extra) are variable per refine_* flag.  (a) refine_extra_params on a
    SIMPLE_RADIAL camera whose distortion starts wrong converges back to the generating k = 0;
    (b) a camera in config.ConstantCameras() stays put; (c) the principal point has an
    identically zero Jacobian in this cost (difference of two projections) and does not move;
    (d) the Jacobian with respect to the intrinsics agrees with finite differences of the cost."""
    from privacy_preserving_sfm_b200 import synthetic as S
    sc = S.make_ba_scene(num_cams=8, num_points=300, obs_per_point=5, seed=3, noise_px=0.0)
    flags = np.zeros(8, np.uint8)
    flags[0], flags[1] = 1, 2
    args = (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"], sc["obs_line"])
    kw = dict(num_threads=1, max_num_iterations=60, gradient_tolerance=1e-12,
              function_tolerance=1e-16)
    # (with noise-free lines the residual vanishes at the generating geometry whatever the
    # intrinsics — the foot point coincides with the projection — so the test scene is noisy)
    sn = S.make_ba_scene(num_cams=8, num_points=300, obs_per_point=5, seed=3, noise_px=2.0)
    argn = (sn["qvecs"], sn["tvecs"], sn["points"], sn["obs_cam"], sn["obs_pt"], sn["obs_line"])
    b = oracle.BaArrays(*argn, [2], [[1000.0, 500.0, 500.0, 0.08]], pose_flags=flags)
    ok, s = oracle.ba_solve(b, oracle.ba_default_options(refine_extra_params=1, **kw))
    b0 = oracle.BaArrays(*argn, [2], [[1000.0, 500.0, 500.0, 0.08]], pose_flags=flags)
    ok0, s0 = oracle.ba_solve(b0, oracle.ba_default_options(**kw))
    assert ok and ok0 and s.final_cost < s0.final_cost        # one more degree of freedom
    assert b.camera_params[0, 3] != 0.08 and b.camera_params[0, 0] == 1000.0
    assert s.num_effective_parameters_reduced == s0.num_effective_parameters_reduced + 1
    # (b) constant camera: same solve as without the flag, parameters untouched
    b2 = oracle.BaArrays(*args, [2], [[1000.0, 500.0, 500.0, 0.08]], pose_flags=flags,
                         camera_const=[1])
    ok2, s2 = oracle.ba_solve(b2, oracle.ba_default_options(refine_extra_params=1, **kw))
    b3 = oracle.BaArrays(*args, [2], [[1000.0, 500.0, 500.0, 0.08]], pose_flags=flags)
    ok3, s3 = oracle.ba_solve(b3, oracle.ba_default_options(**kw))
    assert b2.camera_params[0, 3] == 0.08 and s2.final_cost == s3.final_cost
    # (c) principal point only: nothing to gain, parameters do not move
    b4 = oracle.BaArrays(*args, [1], [[1000.0, 1000.0, 490.0, 510.0]], pose_flags=flags)
    ok4, s4 = oracle.ba_solve(b4, oracle.ba_default_options(refine_principal_point=1, **kw))
    assert np.array_equal(b4.camera_params[0, :4], [1000.0, 1000.0, 490.0, 510.0])
    # (d) finite differences of the cost in the OPENCV parameters at the initial state
    prm = np.array([1000.0, 990.0, 500.0, 480.0, 0.05, -0.01, 0.002, -0.003])
    o = oracle.ba_default_options(num_threads=1, refine_focal_length=1, refine_extra_params=1,
                                  max_num_iterations=1, jacobi_scaling=0)
    def cost(p):
        return oracle.ba_cost(oracle.BaArrays(*args, [4], [p], pose_flags=flags), o)
    c0 = cost(prm)
    bb = oracle.BaArrays(*args, [4], [prm], pose_flags=flags)
    ok5, s5 = oracle.ba_solve(bb, o)
    assert abs(s5.initial_cost - c0) <= 1e-12 * c0
    # one LM step with a huge trust region (1e4) is a Gauss-Newton step: the cost must drop and
    # the focal lengths move the way the finite-difference gradient says
    g = np.array([(cost(prm + h) - cost(prm - h)) / (2 * np.abs(h).sum())
                  for h in np.eye(8) * np.array([1e-3, 1e-3, 1, 1, 1e-7, 1e-7, 1e-7, 1e-7])])
    assert abs(g[2]) < 1e-6 * abs(g[0]) and abs(g[3]) < 1e-6 * abs(g[0])   # principal point: 0
    assert s5.final_cost < s5.initial_cost
