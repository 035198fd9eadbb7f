#!/usr/bin/env python
"""GPU time of the first wave's solve kernel for several wave sizes (PPSFM_RANSAC_TRACE lines).
  PPSFM_SOLVE_OCTET=0|1 PPSFM_OCTETS_PER_WARP=1|2|4 python scripts/solve_latency.py"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np
import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import synthetic as S
H = int(sys.argv[1])
ctx = pp.Context(0)
sc = S.make_abs_pose_scene(n=20000, inlier_ratio=0.30, seed=S.SCENE_SEED)
o = pp.RANSACOptions(max_error=0.012, min_inlier_ratio=0.25, confidence=0.99999,
                     min_num_trials=H, max_num_trials=H)
os.environ["PPSFM_RANSAC_CHUNKS"] = "1"
corr = ctx.upload(sc["lines"], sc["aligned"], sc["points"])
for it in range(6):
    if it >= 3:
        os.environ["PPSFM_RANSAC_TRACE"] = "1"
    ctx.set_prng_seed(it)
    ctx.ransac_p6l_resident(corr, o)
'''
for H in (256, 1024, 3584, 10000):
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT, str(H)], capture_output=True, text=True)
    ts = [float(b) - float(a) for a, b in re.findall(r"gpu solve ([0-9.]+)\.\.([0-9.]+)", r.stderr)]
    print("H=%5d solve ms: %s" % (H, " ".join("%.3f" % t for t in ts)), flush=True)
