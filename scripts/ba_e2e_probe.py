"""End-to-end BA solve (host arrays in / out) at config 4 with phase timing (PPSFM_BA_TIMING=1)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import bundle_adjustment as ba
from privacy_preserving_sfm_b200 import synthetic as S
ctx = pp.Context(0)
sc = S.make_ba_scene(num_cams=500, num_points=200000, obs_per_point=10, seed=S.SCENE_SEED)
flags = np.zeros(500, np.uint8); flags[0], flags[1] = 1, 2
opts = ba.default_solver_options(loss_type=0, max_num_iterations=10, gradient_tolerance=0.0,
                                 function_tolerance=0.0, parameter_tolerance=0.0)
for rep in range(3):
    a = ba.BaArrays(sc["qvecs"], sc["tvecs"], sc["points"], sc["obs_cam"], sc["obs_pt"],
                    sc["obs_line"], [1], [sc["cam_params"]], pose_flags=flags)
    t0 = time.perf_counter()
    ok, s = ba.solve_arrays(ctx, a, opts)
    print(f"rep {rep}: {1e3*(time.perf_counter()-t0):.1f} ms, {s.num_iterations} it, cost {s.final_cost:.6g}")
