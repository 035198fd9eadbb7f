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

# inside the diagonal factorisation (slots 8..14), averaged over the columns
names = ["first 8x8 factor", "panel 0", "update 0 + factor 1", "rounds 1-3", "rounds 4-6",
         "8x8 inverses", "16x16 inverses"]
acc = np.zeros(7)
cnt = 0
for j in range(ncols - 1):
    d = ev[(j, j)]
    if len(d) < 15 or d[8] <= 0 or d[14] <= 0:
        continue
    marks = [d[1], d[8], d[9], d[10], d[11], d[12], d[13], d[14]]
    acc += np.diff(marks) / 1e3
    cnt += 1
if cnt:
    print("inside potrf (us, mean of %d columns): " % cnt +
          ", ".join(f"{n} {v:.2f}" for n, v in zip(names, acc / cnt)) +
          f"; packing after it {np.mean([(ev[(j, j)][2] - ev[(j, j)][14]) / 1e3 for j in range(ncols - 1)]):.2f}")


