#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv
  python scripts/ncu_summary.py raw gpurun_out/prof.ncu-rep [metric-substring ...]
"""
import collections
import csv
import subprocess
import sys

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
       "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
       "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
       "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        k = row["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':48s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}  grid block")
    for k, (c, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:48]:48s} {c:8d} {t:12.1f} {t / c:10.1f} {100 * t / tot:6.1f}%  {g} {b}")


def raw(path, extra):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== {name.split('(')[0]}")
        for i, h in enumerate(hdr):
            if h in KEY or any(e in h for e in extra):
                print(f"  {h:70s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3:])
