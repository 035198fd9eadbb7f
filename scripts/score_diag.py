import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import privacy_preserving_sfm_b200 as pp
from privacy_preserving_sfm_b200 import synthetic as S
ctx = pp.Context(0)
rng = np.random.default_rng(77)
n, k = 4096, 300
R = np.stack([S.random_rotation(rng) for _ in range(k)])
t = rng.uniform(-1, 1, (k, 3))
models = np.concatenate([R.transpose(0, 2, 1).reshape(k, 9), t], axis=1)
X = rng.uniform(-3, 3, (n, 3))
th = rng.uniform(0, 2 * np.pi, n)
lines = np.stack([np.cos(th), np.sin(th), rng.uniform(-1, 1, n)], axis=1)
r3, t3 = R[0][2], t[0][2]
base = X[:64] - np.outer((X[:64] @ r3 + t3), r3)
X[:64] = base
X[64:96] = base[:32] + np.outer(2.0 ** -np.arange(20, 52), r3)
X[96:128] = base[:32] - np.outer(2.0 ** -np.arange(20, 52), r3)
Xc, lc = X.copy(), lines.copy()     # clean copy: the filter is active on it
X[128] *= 1e150
X[129] *= 1e-150
X[130, 0] = np.nan
X[131, 1] = np.inf
lines[132, 2] = np.nan
lines[133] *= 1e100
for name, (L, XX) in {"dirty": (lines, X), "clean": (lc, Xc)}.items():
    res, _, _ = ctx.line_residuals(L, XX, models, 1e-4)
    finite = np.sort(res[1][np.isfinite(res[1]) & (res[1] < 1.0)])
    for thr in [1e-4, finite[len(finite) // 2], np.nextafter(finite[len(finite) // 2], 0.0),
                np.nextafter(finite[len(finite) // 3], 1.0), 0.0, 1e-300, 1e300, np.inf]:
        _, want, _ = ctx.line_residuals(L, XX, models, thr, want_residuals=False)
        for v in (0, 1, 4):
            os.environ["PPSFM_SCORE_VARIANT"] = str(v)
            got = ctx.score_models(L, XX, models, thr).astype(np.uint64)
            bad = np.nonzero(got != want)[0]
            print(name, "thr", thr, "variant", v, "mismatches", len(bad),
                  [(int(b), int(got[b]), int(want[b])) for b in bad[:6]], flush=True)
