#!/usr/bin/env python
"""A few RANSAC calls of H trials in one wave (for ncu captures of the solve kernels).
  python scripts/solve_probe.py H"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import privacy_preserving_sfm_b200 as pp                      # noqa: E402
from privacy_preserving_sfm_b200 import synthetic as S       # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
os.environ["PPSFM_RANSAC_CHUNKS"] = "1"
ctx = pp.Context(0)
sc = S.make_abs_pose_scene(n=20000, inlier_ratio=0.30, seed=S.SCENE_SEED)
o = pp.RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                     min_num_trials=H, max_num_trials=H)
corr = ctx.upload(sc["lines"], sc["aligned"], sc["points"])
for it in range(3):
    ctx.set_prng_seed(it)
    ctx.ransac_p6l_resident(corr, o)
