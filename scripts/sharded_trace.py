#!/usr/bin/env python
"""Timeline of ONE sharded RANSAC call (torchrun, one rank per GPU): PPSFM_RANSAC_TRACE=1 prints
host and GPU times of every wave to stderr on rank 0.
  python -m torch.distributed.run --nproc-per-node N scripts/sharded_trace.py [chunks]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import privacy_preserving_sfm_b200 as pp                      # noqa: E402
from privacy_preserving_sfm_b200 import synthetic as S       # noqa: E402

if len(sys.argv) > 1:
    os.environ["PPSFM_RANSAC_CHUNKS"] = sys.argv[1]
dist.init_process_group("nccl")
rank = dist.get_rank()
torch.cuda.set_device(rank)
N_CORR, N_HYP = 50000, 10000
ctx = pp.Context(rank)
ctx.comm_init_from_torch(dist)
sc = S.make_abs_pose_scene(n=N_CORR, inlier_ratio=0.30, noise_px=1.0, focal=1000.0,
                           aligned_fraction=0.30, seed=S.SCENE_SEED)
opts = pp.RANSACOptions(max_error=12.0 / 1000.0, min_inlier_ratio=0.25, confidence=0.99999,
                        dyn_num_trials_multiplier=3.0, min_num_trials=N_HYP, max_num_trials=N_HYP)
corr = ctx.upload(sc["lines"], sc["aligned"], sc["points"])
mask = np.zeros(N_CORR, dtype=np.uint8)
import time
for it in range(8):
    if it == 7 and rank == 0:
        os.environ["PPSFM_RANSAC_TRACE"] = "1"
    ctx.set_prng_seed(0)
    dist.barrier()
    t0 = time.perf_counter()
    rep, _ = ctx.ransac_p6l_resident_sharded(corr, opts, mask_out=mask)
    if rank == 0:
        print("call %d: %.3f ms" % (it, 1e3 * (time.perf_counter() - t0)), file=sys.stderr)
dist.destroy_process_group()
