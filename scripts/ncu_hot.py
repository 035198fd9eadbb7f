#!/usr/bin/env python
"""Top stall sites of one kernel from an ncu report's source page (SASS view).

  python scripts/ncu_hot.py report.ncu-rep kernel-regex [launch-index] [top-n]
"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                      "regex:" + kre, "--launch-skip-before-match", "0", "-c", str(skip + 1)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several kernels may be concatenated: split at "Kernel Name" rows, keep the last
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s = starts[min(skip, len(starts) - 1)]
e = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(rows)
hdr = rows[s + 1]
body = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in body)
print(rows[s][1][:100], "| samples", tot, "| SASS lines", len(body))
agg = {k: sum(int(r[ci[k]]) for r in body) for k in stalls}
print("stall mix:", ", ".join(f"{k[6:]} {100 * v / max(1, tot):.1f}%" for k, v in
                              sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]]))[:topn]
for i in sorted(order):
    r = body[i]
    n = int(r[ci["# Samples"]])
    top = sorted(((int(r[ci[k]]), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {100 * n / max(1, tot):5.1f}%  {r[ci['Source']].strip()[:70]:70s} "
          f"{top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
