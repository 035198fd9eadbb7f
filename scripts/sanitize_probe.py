#!/usr/bin/env python
"""Small invocations of the hand-synchronised kernels for compute-sanitizer
(memcheck / racecheck / synccheck): the scoring kernel (bulk-copy + mbarrier pipeline, pruned
two-phase path), the octet solve kernel, the dataflow Cholesky (walker + helpers, flag polling)
and the tensor-core Schur gather inside a small BA solve, with and without intrinsics refinement.
  compute-sanitizer --tool racecheck python scripts/sanitize_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import privacy_preserving_sfm_b200 as pp                                    # noqa: E402
from privacy_preserving_sfm_b200 import bundle_adjustment as ba, synthetic as S  # noqa: E402

os.environ["PPSFM_RANSAC_PRUNE_MIN"] = "128"
ctx = pp.Context(0)
sc = S.make_abs_pose_scene(n=3001, inlier_ratio=0.45, seed=1)
o = pp.RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                     min_num_trials=1500, max_num_trials=3000)
ctx.set_prng_seed(0)
rep, mask = ctx.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], o)
print("ransac:", rep.num_trials, rep.num_inliers, flush=True)

rng = np.random.default_rng(0)
for n in (6, 130, 600):
    M = rng.normal(size=(n, n))
    A = M @ M.T + n * np.eye(n)
    b = rng.normal(size=n)
    ok, x = ba.dense_cholesky_solve(ctx, A, b)
    print("cholesky", n, ok, float(np.abs(A @ x - b).max()), flush=True)

sb = S.make_ba_scene(num_cams=12, num_points=500, obs_per_point=5, seed=5)
flags = np.zeros(12, np.uint8)
flags[0], flags[1] = 1, 2
a = ba.BaArrays(sb["qvecs"], sb["tvecs"], sb["points"], sb["obs_cam"], sb["obs_pt"],
                sb["obs_line"], [1], [sb["cam_params"]], pose_flags=flags)
ok, s = ba.solve_arrays(ctx, a, ba.default_solver_options(max_num_iterations=3,
                                                          gradient_tolerance=1e-6))
print("ba:", ok, s.initial_cost, s.final_cost, flush=True)
# intrinsics refinement (two cameras, both variable): the kernels of ba_intrinsics.cu
sn = S.make_ba_scene(num_cams=12, num_points=500, obs_per_point=5, seed=5, noise_px=2.0)
ai = ba.BaArrays(sn["qvecs"], sn["tvecs"], sn["points"], sn["obs_cam"], sn["obs_pt"],
                 sn["obs_line"], [2, 2], [[1000.0, 500, 500, 0.08], [1000.0, 500, 500, 0.05]],
                 image_camera=np.arange(12) % 2, pose_flags=flags)
ok, s = ba.solve_arrays(ctx, ai, ba.default_solver_options(max_num_iterations=3,
                                                           refine_extra_params=1))
print("ba intrinsics:", ok, s.initial_cost, s.final_cost, ai.camera_params[:, 3], flush=True)
ctx.close()
