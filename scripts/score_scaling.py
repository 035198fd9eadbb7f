#!/usr/bin/env python
"""Score-kernel time against the number of hypotheses in one wave (bench scene, no pipelining)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PPSFM_RANSAC_CHUNKS"] = "1"
import privacy_preserving_sfm_b200 as pp                      # noqa: E402
from privacy_preserving_sfm_b200 import synthetic as S       # noqa: E402

N_CORR = 50000
ctx = pp.Context(0)
sc = S.make_abs_pose_scene(n=N_CORR, inlier_ratio=0.30, noise_px=1.0, focal=1000.0,
                           aligned_fraction=0.30, seed=S.SCENE_SEED)
corr = ctx.upload(sc["lines"], sc["aligned"], sc["points"])
mask = np.zeros(N_CORR, dtype=np.uint8)
for segs in [None, "8", "16", "32"]:
    os.environ.pop("PPSFM_SCORE_SEGS", None)
    if segs:
        os.environ["PPSFM_SCORE_SEGS"] = segs
    for H in (1024, 2048, 3072, 6144, 10000, 20000):
        opts = pp.RANSACOptions(max_error=12.0 / 1000.0, min_inlier_ratio=0.25, confidence=0.99999,
                                dyn_num_trials_multiplier=3.0, min_num_trials=H, max_num_trials=H)
        ms, sv = [], []
        for it in range(5):
            ctx.set_prng_seed(0)
            rep, _ = ctx.ransac_p6l_resident(corr, opts, mask_out=mask)
            if it >= 2:
                tm = ctx.ransac_timing()
                ms.append(tm.score_ms)
                sv.append(tm.solve_ms)
        print(f"segs {segs} H {H}: score {np.mean(ms):.3f} ms  solve {np.mean(sv):.3f} ms  "
              f"models {rep.num_models_scored}  ns/model {1e6 * np.mean(ms) / rep.num_models_scored:.1f}",
              flush=True)
