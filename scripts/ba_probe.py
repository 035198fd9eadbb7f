"""Runs a few LM iterations of the config-4 bundle adjustment (500 cams / 200k points / 2M obs);
used under ncu to profile the BA kernels.  argv: [iterations] [cams] [points]"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import bundle_adjustment as ba
from privacy_preserving_sfm_b200 import synthetic as S

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cams = int(sys.argv[2]) if len(sys.argv) > 2 else 500
pts = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
ctx = pp.Context(0)
sc = S.make_ba_scene(num_cams=cams, num_points=pts, obs_per_point=10, seed=S.SCENE_SEED)
flags = np.zeros(cams, np.uint8)
flags[0], flags[1] = 1, 2
arrays = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                     sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags)
opts = ba.default_solver_options(loss_type=0, max_num_iterations=iters, gradient_tolerance=0.0,
                                 function_tolerance=0.0, parameter_tolerance=0.0)
prob = ba.ResidentProblem(ctx, arrays, opts)
for rep in range(2):
    prob.reset()
    t0 = time.perf_counter()
    ok, s = prob.run()
    dt = time.perf_counter() - t0
    print(f"run {rep}: {s.num_iterations} iterations in {dt * 1e3:.2f} ms, cost "
          f"{s.initial_cost:.6g} -> {s.final_cost:.6g}; schur {s.schur_time_s * 1e3:.2f} ms, "
          f"chol {s.cholesky_time_s * 1e3:.2f} ms, backsub {s.backsub_time_s * 1e3:.2f} ms, "
          f"jac {s.jacobian_time_s * 1e3:.2f} ms / {s.jacobian_launches}")
prob.free()
