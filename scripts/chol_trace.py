"""Critical-path breakdown of the dataflow Cholesky from a PPSFM_CHOL_TRACE dump.
slots: 0 task start, 1 k-loop done, 2 published, 3 task end; diagonal tiles also 4 start of the
factorisation, 5/7/9/11 after sub-block rounds 2/4/6/8, 13 block inverses, 14 Lpack stored"""
import sys
import numpy as np
rows = np.loadtxt(sys.argv[1], dtype=np.int64, ndmin=2)
ev = {(int(r[0]), int(r[1])): r[2:].astype(np.float64) for r in rows}
t0 = min(v[0] for v in ev.values() if v[0] > 0)
ncols = max(j for (_, j) in ev) + 1
print("total us", (max(v[2] for v in ev.values()) - t0) / 1e3)
tot = np.zeros(3)
for j in range(ncols):
    d = ev[(j, j)]
    us = lambda a, b: (d[a] - d[b]) / 1e3
    line = (f"{j:3d} kdone {(d[1]-t0)/1e3:8.1f} potrf {us(2,1):6.1f} [pad {us(4,1):4.1f} r01 {us(5,4):4.1f} "
            f"r23 {us(7,5):4.1f} r45 {us(9,7):4.1f} r67 {us(11,9):4.1f} inv {us(13,11):4.1f} "
            f"st {us(14,13):4.1f} pub {us(2,14):4.1f}] tail {us(3,2):5.1f}")
    if (j + 1, j) in ev:
        x = ev[(j + 1, j)]
        line += f" | trsm after diag pub {(x[2]-d[2])/1e3:5.1f}"
        tot[1] += (x[2] - d[2]) / 1e3
        if (j + 1, j + 1) in ev:
            n = ev[(j + 1, j + 1)]
            line += f" | next kdone after trsm pub {(n[1]-x[2])/1e3:5.1f}"
            tot[2] += (n[1] - x[2]) / 1e3
    tot[0] += us(2, 1)
    if j < 8 or j % 8 == 0:
        print(line)
print("sums (us): potrf %.0f, trsm %.0f, next-k %.0f" % tuple(tot))
