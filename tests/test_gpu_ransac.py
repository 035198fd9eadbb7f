"""GPU parity tests of the absolute-pose RANSAC path: CUDA (through the C-ABI) vs the CPU oracle.

Bar: bit-exact residuals, supports, inlier masks, winning (trial, model) indices, models and PRNG
state; these are integer / IEEE-double-without-contraction computations.
"""
import numpy as np
import pytest

from privacy_preserving_sfm_b200 import RANSACOptions, synthetic as S
import privacy_preserving_sfm_b200 as pp

pytestmark = pytest.mark.gpu


def _oracle_opts(O, o):
    return O.make_options(o.max_error, o.min_inlier_ratio, o.confidence,
                          o.dyn_num_trials_multiplier, o.min_num_trials, o.max_num_trials)


def test_sample_table_matches_oracle(ctx, oracle):
    for n, h in [(6, 10), (7, 33), (1000, 500), (50000, 2000)]:
        ctx.set_prng_seed(0)
        oracle.set_prng_seed(0)
        a = ctx.sample_table(n, h)
        b = oracle.sample_table(n, h)
        assert np.array_equal(a, b)
        assert ctx.prng_peek() == oracle.prng_peek()


def test_line_residuals_bit_exact(ctx, oracle):
    sc = S.make_abs_pose_scene(n=4099, seed=3)
    rng = np.random.default_rng(5)
    models = [S.model_from_pose(sc["R"], sc["t"])]
    for _ in range(6):
        models.append(S.model_from_pose(S.random_rotation(rng), rng.uniform(-1, 1, 3)))
    models = np.array(models)
    thr = (12 / 1000.0) ** 2
    res, cnt, sm = ctx.line_residuals(sc["lines"], sc["points"], models, thr)
    for k in range(len(models)):
        r = oracle.line_residuals(sc["lines"], sc["points"], models[k])
        assert np.array_equal(res[k], r)           # bit-exact, incl. DBL_MAX behind the camera
        c, s = oracle.inlier_support(r, thr)
        assert int(cnt[k]) == c
        assert sm[k] == s                           # index-order sum, bit-exact


def test_line_residuals_empty_and_ragged(ctx, oracle):
    sc = S.make_abs_pose_scene(n=37, seed=4)
    m = S.model_from_pose(sc["R"], sc["t"])[None, :]
    res, cnt, sm = ctx.line_residuals(sc["lines"][:0], sc["points"][:0], m, 1e-4)
    assert res.shape == (1, 0) and cnt[0] == 0 and sm[0] == 0.0
    for n in (1, 31, 32, 33, 37):
        res, cnt, sm = ctx.line_residuals(sc["lines"][:n], sc["points"][:n], m, 1e-4)
        r = oracle.line_residuals(sc["lines"][:n], sc["points"][:n], m[0])
        assert np.array_equal(res[0], r)
        assert (int(cnt[0]), sm[0]) == oracle.inlier_support(r, 1e-4)


@pytest.mark.parametrize("octet", [0, 2])
def test_p6l_solve_batch_bit_exact(ctx, oracle, monkeypatch, octet):
    """Both solve kernels — one thread per hypothesis, and eight lanes per hypothesis
    (p6l_octet.cuh, the kernel at the head of every RANSAC call) — against the oracle."""
    monkeypatch.setenv("PPSFM_SOLVE_OCTET", str(octet))
    sc = S.make_abs_pose_scene(n=3000, seed=7)
    ctx.set_prng_seed(1)
    table = ctx.sample_table(3000, 600)
    models, nm = ctx.p6l_solve_batch(sc["lines"], sc["aligned"], sc["points"], table)
    tot = 0
    for t in range(table.shape[0]):
        idx = table[t]
        ref = oracle.p6l_estimate(sc["lines"][idx], sc["aligned"][idx], sc["points"][idx])
        assert nm[t] == len(ref), f"trial {t}"
        assert np.array_equal(models[t, :nm[t]], ref), f"trial {t}"
        tot += nm[t]
    assert tot > 1000


def test_p6l_solve_kernels_agree_on_degenerate_samples(ctx, monkeypatch):
    """Samples that take the rare branches (all lines aligned; a singular translation block ->
    fixed mixing matrix; repeated correspondences -> near-singular elimination and the affine
    change of variables): the two solve kernels must still agree bit for bit."""
    sc = S.make_abs_pose_scene(n=400, seed=17)
    lines, pts, al = sc["lines"].copy(), sc["points"].copy(), sc["aligned"].copy()
    lines[:3] = lines[0]                      # det(l0 l1 l2) = 0
    pts[10:16] = pts[10]                      # six identical points
    lines[20:26, :2] = [1.0, 0.0]             # parallel lines
    al[30:36] = 1                             # all aligned -> no model
    rng = np.random.default_rng(3)
    table = np.stack([rng.permutation(400)[:6] for _ in range(300)]).astype(np.uint32)
    table[0] = [0, 1, 2, 50, 51, 52]
    table[1] = [10, 11, 12, 13, 14, 15]
    table[2] = [20, 21, 22, 23, 24, 25]
    table[3] = [30, 31, 32, 33, 34, 35]
    table[4] = [0, 1, 2, 10, 11, 12]
    table[5] = [7, 7, 7, 7, 7, 7]
    outs = []
    for octet in (0, 2):
        monkeypatch.setenv("PPSFM_SOLVE_OCTET", str(octet))
        models, nm = ctx.p6l_solve_batch(lines, al, pts, table)
        outs.append((models.copy(), nm.copy()))
    assert np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][1][3] == 0
    for t in range(len(table)):
        a, b = outs[0][0][t, :outs[0][1][t]], outs[1][0][t, :outs[1][1][t]]
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), t   # NaNs compare by bits


def test_p6l_recovers_generating_pose(ctx):
    probs = S.make_p6l_minimal_problems(200, seed=11)
    est = pp.P6LEstimator(ctx)
    errs = []
    for p in probs:
        sols = est.Estimate(p["lines"], p["points"])
        gt = np.hstack([p["R"], p["t"][:, None]])
        errs.append(min(np.abs(s - gt).max() for s in sols))
    assert np.median(errs) < 1e-10 and max(errs) < 1e-6


def test_p6l_all_aligned_returns_no_model(ctx):
    p = S.make_p6l_minimal_problems(1, seed=2)[0]
    est = pp.P6LEstimator(ctx)
    assert est.Estimate((p["lines"], np.ones(6, np.uint8)), p["points"]) == []


@pytest.mark.parametrize("n,opts,cases", [
    (2000, dict(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999, min_num_trials=100,
                max_num_trials=10000), [(0.3, 1), (0.7, 2), (0.05, 3)]),   # mapper settings
    (5000, dict(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999, min_num_trials=3000,
                max_num_trials=3000), [(0.3, 1), (0.9, 2)]),               # fixed trial count
    (1500, dict(max_error=0.012, min_inlier_ratio=0.1, confidence=0.99, min_num_trials=0,
                max_num_trials=2**64 - 1), [(0.5, 1), (0.7, 2), (1.0, 3)]),  # defaults, early abort
    (800, dict(max_error=0.012, min_inlier_ratio=0.3, confidence=0.9999, min_num_trials=0,
               max_num_trials=2**64 - 1), [(0.35, 4), (0.6, 5)]),          # several waves
])
def test_ransac_matches_oracle(ctx, oracle, n, opts, cases):
    for inlier_ratio, seed in cases:
        sc = S.make_abs_pose_scene(n=n, inlier_ratio=inlier_ratio, seed=seed)
        o = RANSACOptions(**opts)
        ctx.set_prng_seed(0)
        oracle.set_prng_seed(0)
        rep, mask = ctx.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], o)
        oref, omask = oracle.ransac_p6l(sc["lines"], sc["aligned"], sc["points"],
                                        _oracle_opts(oracle, o))
        tag = f"n={n} ratio={inlier_ratio}"
        assert rep.success == oref.success, tag
        assert rep.num_trials == oref.num_trials, tag
        assert rep.num_inliers == oref.num_inliers, tag
        assert rep.residual_sum == oref.residual_sum, tag
        assert (rep.best_trial, rep.best_model_idx) == (oref.best_trial, oref.best_model_idx), tag
        assert list(rep.model) == list(oref.model), tag
        assert rep.num_models_scored == oref.num_models_scored, tag
        if rep.success:
            assert np.array_equal(mask, omask), tag
        assert ctx.prng_peek() == oracle.prng_peek(), tag   # same number of PRNG draws


def test_ransac_many_tied_candidates_stay_cheap(ctx, oracle):
    """The mapper's registration calls: few hundred correspondences, almost all inliers, so most
    good models TIE at the full inlier count and a wave holds thousands of candidates for the
    index-order support pass.  They are gathered and evaluated in a few launches (this call took
    108 launches / 9.5 ms when candidates were copied one by one in batches of 32), and the result
    is the oracle's."""
    sc = S.make_abs_pose_scene(n=500, inlier_ratio=0.95, noise_px=0.5, aligned_fraction=0.4, seed=5)
    o = RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                      min_num_trials=100, max_num_trials=10000)
    ctx.set_prng_seed(0)
    oracle.set_prng_seed(0)
    rep, mask = ctx.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], o)
    tm = ctx.ransac_timing()
    oref, omask = oracle.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], _oracle_opts(oracle, o))
    assert rep.success and oref.success and rep.num_trials == oref.num_trials
    assert (rep.best_trial, rep.best_model_idx) == (oref.best_trial, oref.best_model_idx)
    assert rep.residual_sum == oref.residual_sum and np.array_equal(mask, omask)
    assert tm.kernel_launches <= 16, tm.kernel_launches


@pytest.mark.parametrize("first,growth,chunks,prune_min", [
    (10, 2, 4, 2048), (3, 100, 4, 2048), (4, 3, 4, 2048), (3, 100, 1, 2048),
    (10, 2, 4, 128), (3, 100, 4, 128), (5, 2, 4, 256)])
def test_ransac_wave_pipeline_is_invisible(ctx, oracle, monkeypatch, first, growth, chunks,
                                           prune_min):
    """The trial loop runs as a pipeline of waves, issued ahead of the replay where the loop is
    certain to get there.  Whatever the partition: same report, mask and generator state as the
    serial loop — including an adaptive abort inside a wave while later waves are already sampled
    and in flight (the generator is rewound to the aborting wave's snapshot).  Waves after the
    first are scored in two phases with exact pruning (models that cannot reach the best count of
    the earlier waves skip the second phase); prune_min lowers the size threshold of that path so
    that these small sets go through it."""
    monkeypatch.setenv("PPSFM_RANSAC_PRUNE_MIN", str(prune_min))
    monkeypatch.setenv("PPSFM_RANSAC_FIRST", str(first))
    monkeypatch.setenv("PPSFM_RANSAC_GROWTH", str(growth))
    monkeypatch.setenv("PPSFM_RANSAC_CHUNKS", str(chunks))
    for n, ratio, min_trials, max_trials, seed in [
            (3000, 0.45, 2048, 10000, 1), (3000, 0.5, 2048, 10000, 2), (2500, 0.4, 1200, 10000, 3),
            (3000, 0.6, 5000, 10000, 4), (2000, 0.35, 0, 10000, 5), (2000, 0.3, 4000, 4000, 6),
            (12000, 0.3, 2500, 2500, 7), (2001, 0.4, 1500, 10000, 8), (777, 0.5, 1100, 3000, 9)]:
        sc = S.make_abs_pose_scene(n=n, inlier_ratio=ratio, seed=seed)
        o = RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                          min_num_trials=min_trials, max_num_trials=max_trials)
        ctx.set_prng_seed(0)
        oracle.set_prng_seed(0)
        rep, mask = ctx.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], o)
        oref, omask = oracle.ransac_p6l(sc["lines"], sc["aligned"], sc["points"],
                                        _oracle_opts(oracle, o))
        tag = f"n={n} ratio={ratio} min={min_trials}"
        assert (rep.success, rep.num_trials, rep.num_inliers) == \
            (oref.success, oref.num_trials, oref.num_inliers), tag
        assert rep.residual_sum == oref.residual_sum, tag
        assert (rep.best_trial, rep.best_model_idx) == (oref.best_trial, oref.best_model_idx), tag
        assert list(rep.model) == list(oref.model), tag
        assert rep.num_models_scored == oref.num_models_scored, tag
        assert np.array_equal(mask, omask), tag
        assert ctx.prng_peek() == oracle.prng_peek(), tag


def test_ransac_too_few_samples(ctx):
    sc = S.make_abs_pose_scene(n=5, seed=9)
    rep, _ = ctx.ransac_p6l(sc["lines"], sc["aligned"], sc["points"],
                            RANSACOptions(max_error=0.01))
    assert rep.success == 0 and rep.num_trials == 0


def test_estimate_absolute_pose_from_lines(ctx, oracle):
    sc = S.make_abs_pose_scene(n=4000, inlier_ratio=0.4, seed=21)
    o = RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                      min_num_trials=100, max_num_trials=10000)
    ctx.set_prng_seed(0)
    oracle.set_prng_seed(0)
    ok, q, t, ninl, mask = pp.EstimateAbsolutePoseFromLines(
        o, (sc["lines"], sc["aligned"]), sc["points"], ctx=ctx)
    ok2, q2, t2, ninl2, mask2, _ = oracle.estimate_absolute_pose_from_lines(
        sc["lines"], sc["aligned"], sc["points"], _oracle_opts(oracle, o))
    assert ok and ok2
    assert np.array_equal(q, q2) and np.array_equal(t, t2)
    assert ninl == ninl2 and np.array_equal(mask, mask2)
    # and it is the right pose: rotation angle error < 0.2 deg, translation < 1e-2
    Rgt = sc["R"]
    Rest = S.quat_to_rotmat(q / np.linalg.norm(q))
    ang = np.degrees(np.arccos(np.clip((np.trace(Rest.T @ Rgt) - 1) / 2, -1, 1)))
    assert ang < 0.2 and np.linalg.norm(t - sc["t"]) < 1e-2


def test_estimate_rejects_mostly_aligned_inliers(ctx, oracle):
    sc = S.make_abs_pose_scene(n=2000, inlier_ratio=0.5, aligned_fraction=0.97, seed=5)
    o = RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                      min_num_trials=100, max_num_trials=2000)
    ctx.set_prng_seed(0)
    oracle.set_prng_seed(0)
    ok, *_ = pp.EstimateAbsolutePoseFromLines(o, (sc["lines"], sc["aligned"]), sc["points"],
                                              ctx=ctx)
    ok2, *_ = oracle.estimate_absolute_pose_from_lines(sc["lines"], sc["aligned"], sc["points"],
                                                       _oracle_opts(oracle, o))
    assert ok == ok2 and not ok      # pose.cc:71-83


def test_options_check(ctx):
    with pytest.raises(pp.PpsfmError):
        pp.RANSAC_P6L(RANSACOptions(max_error=0.0), ctx)
    with pytest.raises(pp.PpsfmError):
        pp.RANSAC_P6L(RANSACOptions(max_error=1.0, min_num_trials=5, max_num_trials=4), ctx)


def _filter_case(rng, n=4096, k=300):
    R = np.stack([S.random_rotation(rng) for _ in range(k)])
    t = rng.uniform(-1, 1, (k, 3))
    models = np.concatenate([R.transpose(0, 2, 1).reshape(k, 9), t], axis=1)   # column-major 3x4
    X = rng.uniform(-3, 3, (n, 3))
    th = rng.uniform(0, 2 * np.pi, n)
    lines = np.stack([np.cos(th), np.sin(th), rng.uniform(-1, 1, n)], axis=1)
    # points exactly on / next to the camera plane of model 0 (pz = 0, +-eps, +-tiny)
    r3, t3 = R[0][2], t[0][2]
    base = X[:64] - np.outer((X[:64] @ r3 + t3), r3)            # pz ~ 0 up to rounding
    X[:64] = base
    X[64:96] = base[:32] + np.outer(2.0 ** -np.arange(20, 52), r3)
    X[96:128] = base[:32] - np.outer(2.0 ** -np.arange(20, 52), r3)
    return models, lines, X


def _check_counts(ctx, oracle, lines, X, models, tag, oracle_models=(0, 1, 7)):
    res, _, _ = ctx.line_residuals(lines, X, models, 1e-4)
    # thresholds that hit residuals exactly: r_max^2 = a residual of model 1, and its neighbours
    finite = np.sort(res[1][np.isfinite(res[1]) & (res[1] < 1.0)])
    thrs = [1e-4, 0.0, 1e-300, 1e300, np.inf]
    if len(finite) > 3:
        thrs += [finite[len(finite) // 2], np.nextafter(finite[len(finite) // 2], 0.0),
                 np.nextafter(finite[len(finite) // 3], 1.0)]
    for thr in thrs:
        _, want, _ = ctx.line_residuals(lines, X, models, thr, want_residuals=False)
        got = ctx.score_models(lines, X, models, thr)
        assert np.array_equal(got.astype(np.uint64), want), (tag, thr)
        # and against the CPU oracle for a few models
        for m in oracle_models:
            r_cpu = oracle.line_residuals(lines, X, models[m])
            assert int((r_cpu <= thr).sum()) == int(got[m]), (tag, thr, m)


def test_score_filter_edge_cases(ctx, oracle):
    """The count-only scoring kernel decides almost every pair with division-free filters (a float
    stage, then an FP64 stage) and falls back to the reference arithmetic inside their error
    bands.  Aim at the bands: residuals that equal the threshold exactly (and its neighbours),
    points on / next to the camera plane, magnitudes at and beyond the float stage's range, odd
    and tiny set sizes, NaN / inf inputs — counts must stay bit-identical to the reference."""
    rng = np.random.default_rng(77)
    models, lines, X = _filter_case(rng)
    _check_counts(ctx, oracle, lines, X, models, "clean")            # both filter stages active
    for n_odd in (1, 2, 3, 255, 257, 1001):                           # ragged ends of the pairing
        _check_counts(ctx, oracle, lines[:n_odd], X[:n_odd], models, "n=%d" % n_odd, (1,))
    # magnitudes: the scene scaled up to the float stage's limit and beyond, and down into and
    # below the float subnormal range (residuals are invariant under a common scale of X and t)
    for e in (20, 36, 39, 45, 200, -20, -40, -60, -200):
        sc = 2.0 ** e
        m2 = models.copy()
        m2[:, 9:] *= sc
        _check_counts(ctx, oracle, lines, X * sc, m2, "scale 2^%d" % e, (1,))
    # a set with non-finite and extreme entries: every filter stage must stand down
    X[128] *= 1e150
    X[129] *= 1e-150
    X[130, 0] = np.nan
    X[131, 1] = np.inf
    lines[132, 2] = np.nan
    lines[133] *= 1e100
    _check_counts(ctx, oracle, lines, X, models, "dirty")
    # models with non-finite / huge / zero entries next to ordinary ones
    m3 = models.copy()
    m3[3, 4] = np.nan
    m3[4, 11] = np.inf
    m3[5] *= 1e60
    m3[6] = 0.0
    m3[8] *= 1e-60
    models_c, lines_c, X_c = _filter_case(np.random.default_rng(78))
    _check_counts(ctx, oracle, lines_c, X_c, m3, "odd models", (1, 5, 6))
