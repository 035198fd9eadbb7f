#!/usr/bin/env python
"""A/B timing of the score-kernel variants (PPSFM_SCORE_VARIANT) on the bench workload.

  python scripts/score_probe.py [variants...]        (run on the GPU box)
Prints score-kernel ms per step (CUDA events inside the library) and checks that every variant
returns the same RANSAC report as variant 0."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import privacy_preserving_sfm_b200 as pp                      # noqa: E402
from privacy_preserving_sfm_b200 import synthetic as S       # noqa: E402

N_CORR, N_HYP = 50000, 10000
variants = [v for v in sys.argv[1:]] or ["0", "4", "10", "11", "12", "13", "14", "15", "16", "17",
                                         "13:24", "13:32", "13:48", "15:32", "15:48", "11:32"]
ctx = pp.Context(0)
sc = S.make_abs_pose_scene(n=N_CORR, inlier_ratio=0.30, noise_px=1.0, focal=1000.0,
                           aligned_fraction=0.30, seed=S.SCENE_SEED)
opts = pp.RANSACOptions(max_error=12.0 / 1000.0, min_inlier_ratio=0.25, confidence=0.99999,
                        dyn_num_trials_multiplier=3.0, min_num_trials=N_HYP, max_num_trials=N_HYP)
corr = ctx.upload(sc["lines"], sc["aligned"], sc["points"])
mask = np.zeros(N_CORR, dtype=np.uint8)
ref = None
for rnd in range(2):
    for v in variants:
        os.environ["PPSFM_SCORE_VARIANT"] = v.split(":")[0]     # "variant[:segments]"
        os.environ.pop("PPSFM_SCORE_SEGS", None)
        if ":" in v:
            os.environ["PPSFM_SCORE_SEGS"] = v.split(":")[1]
        ms = []
        for it in range(6):
            ctx.bench_l2_flush()
            ctx.set_prng_seed(0)
            rep, _ = ctx.ransac_p6l_resident(corr, opts, mask_out=mask)
            if it >= 2:
                ms.append(ctx.ransac_timing().score_ms)
        key = (int(rep.num_inliers), int(rep.num_trials), int(mask.sum()))
        if ref is None:
            ref = key
        print(f"round {rnd} variant {v}: score {np.mean(ms):.3f} ms (min {np.min(ms):.3f})  "
              f"{'same' if key == ref else 'DIFFERENT ' + str(key) + ' vs ' + str(ref)}", flush=True)
