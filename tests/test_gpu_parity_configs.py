"""Parity at the BASELINE.json configuration sizes: the CUDA path through the
C-ABI against the CPU oracle on exactly the workloads bench.py times.

  config 2  absolute-pose P6L RANSAC, 50 000 correspondences x 10 000 hypotheses, seed 0:
            bit-exact report, winner, mask and generator state; plus a differential of the
            count-only scoring kernel (float -> FP64 -> reference cascade) against the exact
            residual kernel over EVERY model of that run (~1.9e9 pairs) and a near-threshold fuzz.
  config 3  line-reprojection BA, 100 cameras / 30 000 points / 300 000 observations, TRIVIAL and
            SOFT_L1, the mapper's solver options: cost 1e-9, poses 1e-6 rad / 1e-6 relative
            translation (the north-star bar; observed agreement is tighter), points 1e-6
            relative, same accept / reject trace.
  config 4  500 cameras / 200 000 points / 2 000 000 observations: three LM iterations.
"""
import numpy as np
import pytest

import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import RANSACOptions, bundle_adjustment as ba, synthetic as S

pytestmark = pytest.mark.gpu

N_CORR, N_HYP, MAX_ERROR = 50000, 10000, 12.0 / 1000.0


def _config2_scene():
    # bench.py: make_scene()
    return S.make_abs_pose_scene(n=N_CORR, inlier_ratio=0.30, noise_px=1.0, focal=1000.0,
                                 aligned_fraction=0.30, seed=S.SCENE_SEED)


def _config2_options():
    return RANSACOptions(max_error=MAX_ERROR, min_inlier_ratio=0.25, confidence=0.99999,
                         dyn_num_trials_multiplier=3.0, min_num_trials=N_HYP,
                         max_num_trials=N_HYP)


def test_config2_ransac_matches_oracle(ctx, oracle):
    sc, o = _config2_scene(), _config2_options()
    ctx.set_prng_seed(0)
    oracle.set_prng_seed(0)
    rep, mask = ctx.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], o)
    oo = oracle.make_options(o.max_error, o.min_inlier_ratio, o.confidence,
                             o.dyn_num_trials_multiplier, o.min_num_trials, o.max_num_trials)
    oref, omask = oracle.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], oo)
    assert rep.success == oref.success == 1
    assert rep.num_trials == oref.num_trials == N_HYP
    assert (rep.best_trial, rep.best_model_idx) == (oref.best_trial, oref.best_model_idx)
    assert rep.num_inliers == oref.num_inliers
    assert rep.residual_sum == oref.residual_sum           # index-order sum, bit-exact
    assert list(rep.model) == list(oref.model)             # pose bit-exact (bar: 1e-12)
    assert rep.num_models_scored == oref.num_models_scored
    assert np.array_equal(mask, omask)
    assert ctx.prng_peek() == oracle.prng_peek()
    # the resident entry point (what bench.py's `value` times) gives the same report
    corr = ctx.upload(sc["lines"], sc["aligned"], sc["points"])
    ctx.set_prng_seed(0)
    rep2, mask2 = ctx.ransac_p6l_resident(corr, o)
    corr.free()
    assert (rep2.num_trials, rep2.best_trial, rep2.best_model_idx, rep2.num_inliers) == \
        (rep.num_trials, rep.best_trial, rep.best_model_idx, rep.num_inliers)
    assert list(rep2.model) == list(rep.model) and np.array_equal(mask2, mask)
    # and it is the generating pose up to what a minimal sample of six noisy lines gives
    R = np.array(rep.model[:9]).reshape(3, 3).T
    ang = np.degrees(np.arccos(np.clip((np.trace(R.T @ sc["R"]) - 1) / 2, -1, 1)))
    assert ang < 0.5 and np.linalg.norm(np.array(rep.model[9:]) - sc["t"]) < 2e-2


def _all_models_of_config2(ctx, sc):
    ctx.set_prng_seed(0)
    table = ctx.sample_table(N_CORR, N_HYP)
    models, nm = ctx.p6l_solve_batch(sc["lines"], sc["aligned"], sc["points"], table)
    flat = np.concatenate([models[t, :nm[t]] for t in range(N_HYP) if nm[t] > 0])
    return flat, nm


def test_config2_score_filter_differential_over_all_models(ctx, oracle):
    """Every model the config-2 call scores (~37.7 k models x 50 k correspondences = 1.9e9 pairs):
    the count-only kernel of the RANSAC loop must equal the exact residual kernel's counts, which
    follow utils.cc:64-88 operation by operation (and equal the CPU oracle on a sample)."""
    sc = _config2_scene()
    flat, nm = _all_models_of_config2(ctx, sc)
    assert flat.shape[0] > 30000 and int(nm.sum()) == flat.shape[0]
    thr = MAX_ERROR * MAX_ERROR
    got = ctx.score_models(sc["lines"], sc["points"], flat, thr)
    _, want, _ = ctx.line_residuals(sc["lines"], sc["points"], flat, thr, want_residuals=False)
    assert np.array_equal(got.astype(np.uint64), want)
    rng = np.random.default_rng(1)
    for m in rng.choice(flat.shape[0], 24, replace=False):
        r = oracle.line_residuals(sc["lines"], sc["points"], flat[m])
        assert int((r <= thr).sum()) == int(got[m])
    # the best count of the differential is the call's winner
    assert int(got.max()) > 0.28 * N_CORR


def test_score_filter_near_threshold_fuzz(ctx, oracle):
    """Thresholds drawn from the residuals themselves: for 160 (model, correspondence) pairs of
    the config-2 set the threshold is put exactly on the pair's residual, one ulp below and one
    ulp above, so that the filter stages' bands straddle real data at every magnitude the set
    holds; counts over all 50 000 correspondences must equal the exact kernel's."""
    sc = _config2_scene()
    rng = np.random.default_rng(2)
    models = np.stack([S.model_from_pose(sc["R"], sc["t"])] +
                      [S.model_from_pose(S.random_rotation(rng), rng.uniform(-1, 1, 3))
                       for _ in range(15)])
    # perturbed copies of the true pose: many residuals close to the threshold
    for k in range(1, 8):
        models[k] = S.model_from_pose(sc["R"], sc["t"] + rng.normal(0, 2e-3 * k, 3))
    res, _, _ = ctx.line_residuals(sc["lines"], sc["points"], models, 1.0)
    thrs = []
    for _ in range(160):
        m, i = rng.integers(models.shape[0]), rng.integers(N_CORR)
        r = res[m, i]
        if np.isfinite(r) and r < 1e300:
            thrs += [r, np.nextafter(r, 0.0), np.nextafter(r, np.inf)]
    # and around the call's own threshold
    t0 = MAX_ERROR * MAX_ERROR
    thrs += [t0, np.nextafter(t0, 0.0), np.nextafter(t0, 1.0)]
    for thr in thrs[::3][:60] + thrs[1::3][:60] + thrs[2::3][:60]:
        got = ctx.score_models(sc["lines"], sc["points"], models, thr)
        want = (res <= thr).sum(axis=1)
        assert np.array_equal(got.astype(np.int64), want.astype(np.int64)), thr
    r_cpu = oracle.line_residuals(sc["lines"], sc["points"], models[3])
    assert np.array_equal(r_cpu, res[3])


# ------------------------------------------------------------------------------------------------
def _ba_problem(oracle, cams, points, seed=None):
    sc = S.make_ba_scene(num_cams=cams, num_points=points, obs_per_point=10,
                         seed=S.SCENE_SEED if seed is None else seed)
    flags = np.zeros(cams, np.uint8)
    flags[0], flags[1] = 1, 2     # src/sfm/incremental_mapper.cc:907-926
    args = (sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"], sc["obs_line"],
            [1], [sc["cam_params"]])
    return sc, ba.BaArrays(*args, pose_flags=flags), oracle.BaArrays(*args, pose_flags=flags)


def _rotation_angle(q1, q2):
    """angle (rad) of the relative rotation between unit quaternions (rows)"""
    d = np.abs(np.sum(q1 * q2, axis=1)).clip(0, 1)
    return 2 * np.arccos(d)


def _assert_ba_parity(a, b, s, s2, same_trace=True):
    assert abs(s.initial_cost - s2.initial_cost) <= 1e-11 * s2.initial_cost
    assert abs(s.final_cost - s2.final_cost) <= 1e-9 * s2.final_cost
    if same_trace:
        n = s2.trace_len
        assert s.trace_len == n
        assert list(s.trace_accepted[:n]) == list(s2.trace_accepted[:n])
        assert (s.num_successful_steps, s.num_unsuccessful_steps) == \
            (s2.num_successful_steps, s2.num_unsuccessful_steps)
        assert s.termination_type == s2.termination_type
        for i in range(n):
            assert abs(s.trace_cost[i] - s2.trace_cost[i]) <= 1e-9 * s2.trace_cost[i], i
    qa = a.qvecs / np.linalg.norm(a.qvecs, axis=1, keepdims=True)
    qb = b.qvecs / np.linalg.norm(b.qvecs, axis=1, keepdims=True)
    assert _rotation_angle(qa, qb).max() <= 1e-6                      # rad
    tn = np.linalg.norm(b.tvecs, axis=1)
    assert (np.linalg.norm(a.tvecs - b.tvecs, axis=1) / tn).max() <= 1e-6
    scale = np.abs(b.points).max()
    assert np.abs(a.points - b.points).max() <= 1e-6 * scale


@pytest.mark.parametrize("loss,grad_tol", [(0, 1.0), (1, 1.0), (0, 1e-4)])
def test_config3_ba_matches_oracle(ctx, oracle, loss, grad_tol):
    """BASELINE config 3 with the mapper's solver options (global BA: TRIVIAL; local BA: SOFT_L1
    scale 1; 50 iterations, function / parameter tolerance 0, gradient tolerance 1.0 —
    src/controllers/incremental_mapper.cc:215-243); the third case tightens the gradient
    tolerance so that the comparison is made at a deeply converged minimum."""
    sc, a, b = _ba_problem(oracle, 100, 30000)
    kw = dict(loss_type=loss, loss_scale=1.0, max_num_iterations=50, gradient_tolerance=grad_tol,
              function_tolerance=0.0, parameter_tolerance=0.0)
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(**kw))
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(num_threads=-1, **kw))
    assert ok and ok2
    assert s2.termination_type == 0 and s.final_cost < 0.05 * s.initial_cost
    _assert_ba_parity(a, b, s, s2)


def test_config4_ba_first_iterations_match_oracle(ctx, oracle):
    """BASELINE config 4 (500 cameras / 200 k points / 2 M observations), the problem bench.py
    times: three LM iterations against the oracle's dense-Schur LM."""
    _, a, b = _ba_problem(oracle, 500, 200000)
    kw = dict(loss_type=0, max_num_iterations=3, gradient_tolerance=0.0, function_tolerance=0.0,
              parameter_tolerance=0.0)
    ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(**kw))
    ok2, s2 = oracle.ba_solve(b, oracle.ba_default_options(num_threads=-1, **kw))
    assert ok and ok2
    _assert_ba_parity(a, b, s, s2)
