#!/usr/bin/env python
"""Writes tests/golden/config2_reference.json: BASELINE.json configs[1] (50 000 lifted-line
correspondences x 10 000 hypotheses, PRNG seed 0, the scene bench.py times) run through the
REFERENCE'S OWN RANSAC loop, P6L solver, re3q3 and scoring sources as compiled here by
oracle/build_ref.sh (oracle/_ref/libref_p6l.so; Eigen and glog replaced by the stand-ins of
oracle/ref/shim/).  About 15-30 s of one core.

  python tests/golden/make_config2_reference.py

tests/test_golden.py requires the oracle, tests/test_gpu_golden.py the CUDA path, to reproduce
the file bit for bit (doubles as hex strings, the 50 000-entry inlier mask by SHA-256)."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from privacy_preserving_sfm_b200 import synthetic as S         # noqa: E402

SCENE = dict(n=50000, inlier_ratio=0.30, noise_px=1.0, focal=1000.0, aligned_fraction=0.30,
             seed=S.SCENE_SEED)
OPTIONS = (12.0 / 1000.0, 0.25, 0.99999, 3.0, 10000, 10000)


def run(impl):
    """impl: oracle.reference (the reference build) or oracle (the restatement) — same surface."""
    import oracle as O
    sc = S.make_abs_pose_scene(**SCENE)
    impl.set_prng_seed(0)
    rep, mask = impl.ransac_p6l(sc["lines"], sc["aligned"], sc["points"], O.make_options(*OPTIONS))
    return {
        "success": int(rep.success), "num_trials": int(rep.num_trials),
        "num_inliers": int(rep.num_inliers), "residual_sum": float(rep.residual_sum).hex(),
        "model": [float(x).hex() for x in rep.model],
        "mask_sha256": hashlib.sha256(np.asarray(mask, np.uint8).tobytes()).hexdigest(),
        "prng_peek_after": int(impl.prng_peek()),
    }


if __name__ == "__main__":
    import oracle.reference as R
    out = {"generated_by": "tests/golden/make_config2_reference.py (oracle/_ref/libref_p6l.so: the "
                           "reference's sources compiled against the Eigen / glog stand-ins)",
           "scene": SCENE, "options": list(OPTIONS), "prng_seed": 0, "expect": run(R)}
    path = os.path.join(HERE, "config2_reference.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, out["expect"]["num_inliers"], "inliers")
