"""Latency of registration-shaped RANSAC calls (the mapper's settings) at several sizes and inlier
ratios; high ratios make most good models tie at the full inlier count."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import synthetic as S
from privacy_preserving_sfm_b200.estimators import EstimateAbsolutePoseFromLines
ctx = pp.Context(0)
for n, ratio in [(500, 0.95), (2000, 0.95), (2000, 0.5), (50000, 0.9), (50000, 0.99), (20000, 1.0)]:
    sc = S.make_abs_pose_scene(n=n, inlier_ratio=ratio, noise_px=0.5, focal=1000.0, aligned_fraction=0.4, seed=5)
    opt = pp.RANSACOptions(max_error=12.0 / 1000.0, min_inlier_ratio=0.25, confidence=0.99999,
                           min_num_trials=100, max_num_trials=10000)
    for rep in range(3):
        t0 = time.perf_counter()
        ok, q, t, ninl, mask = EstimateAbsolutePoseFromLines(opt, (sc["lines"], sc["aligned"]), sc["points"], ctx=ctx)
        dt = time.perf_counter() - t0
    tm = ctx.ransac_timing()
    print(n, ratio, "ok", ok, "inliers", ninl, "ms %.3f" % (1e3 * dt), {k: getattr(tm, k) for k, _ in tm._fields_})
