"""Times the reduced-camera-system solver alone (n = 2994 like config 4); used under ncu."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import bundle_adjustment as ba

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2994
ctx = pp.Context(0)
rng = np.random.default_rng(0)
M = rng.normal(size=(n, 64))
A = M @ M.T + n * np.eye(n)
b = rng.normal(size=n)
for _ in range(2):
    t0 = time.perf_counter()
    ok, x = ba.dense_cholesky_solve(ctx, A, b)
    print("call s", time.perf_counter() - t0, ok, np.abs(A @ x - b).max())
