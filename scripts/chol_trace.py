"""Critical-path breakdown of the dataflow Cholesky from a PPSFM_CHOL_TRACE dump.
slots: 0 task start, 1 k-loop done, 2 published, 3 task end; diagonal tiles also 4 padded,
5/7/9/11 factor16 of sub-block 0..3 done, 6/8/10 panel+update done, 13 inverses, 14 stores"""
import sys
import numpy as np
rows = np.loadtxt(sys.argv[1], dtype=np.int64)
ev = {(int(r[0]), int(r[1])): r[2:].astype(np.float64) for r in rows}
t0 = min(v[0] for v in ev.values() if v[0] > 0)
ncols = max(j for (_, j) in ev) + 1
print("total us", (max(v[2] for v in ev.values()) - t0) / 1e3)
for j in range(ncols):
    d = ev[(j, j)]
    us = lambda a, b: (d[a] - d[b]) / 1e3
    line = (f"{j:3d} kdone {(d[1]-t0)/1e3:8.1f} potrf {us(2,1):6.1f} [pad {us(4,1):4.1f} f0 {us(5,4):4.1f} "
            f"pu0 {us(6,5):4.1f} f1 {us(7,6):4.1f} pu1 {us(8,7):4.1f} f2 {us(9,8):4.1f} pu2 {us(10,9):4.1f} "
            f"f3 {us(11,10):4.1f} inv {us(13,11):4.1f} st {us(14,13):4.1f} pub {us(2,14):4.1f}] tail {us(3,2):5.1f}")
    if (j + 1, j) in ev:
        x = ev[(j + 1, j)]
        line += f" | trsm after diag pub {(x[2]-d[2])/1e3:5.1f}"
        if (j + 1, j + 1) in ev:
            n = ev[(j + 1, j + 1)]
            line += f" | next kdone after trsm pub {(n[1]-x[2])/1e3:5.1f}"
    print(line)
