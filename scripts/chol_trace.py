"""Critical-path breakdown of the dataflow Cholesky from a PPSFM_CHOL_TRACE dump (walker CTA).
Diagonal-tile slots: 1 column start, 2 factorisation done (Lpack packed), 4 sub-diagonal tile in
shared memory, 5 triangular solve done, 6 pre-accumulated next diagonal tile in shared memory,
7 its update done.  Helper tiles: 0 task start, 3 k-loop done."""
import sys
import numpy as np
rows = np.loadtxt(sys.argv[1], dtype=np.int64, ndmin=2)
ev = {(int(r[0]), int(r[1])): r[2:].astype(np.float64) for r in rows}
ncols = max(j for (_, j) in ev) + 1
t0 = ev[(0, 0)][1]
tot = np.zeros(5)
print("col  start  | potrf  wait-sub  trsm  wait-diag  syrk | column")
for j in range(ncols):
    d = ev[(j, j)]
    us = lambda a, b: (d[a] - d[b]) / 1e3 if d[a] > 0 and d[b] > 0 else float("nan")
    parts = [us(2, 1), us(4, 2), us(5, 4), us(6, 5), us(7, 6)]
    tot += np.nan_to_num(parts)
    if j < 6 or j % 8 == 0 or j == ncols - 1:
        print(f"{j:3d} {(d[1]-t0)/1e3:7.1f} | " + "  ".join(f"{p:6.1f}" for p in parts) + f" | {us(7, 1):6.1f}")
print("sums (us): potrf %.0f, wait-sub %.0f, trsm %.0f, wait-diag %.0f, syrk %.0f" % tuple(tot))
last = max(v[7] for v in ev.values())
print("walker total us", (max(max(v[2], v[5], v[7]) for (i, j), v in ev.items() if i == j) - t0) / 1e3)
