#!/usr/bin/env python
"""Markdown table of the bench lines at N = 1, 2, 4, 8 (profiles/r02_bench_n*.json) with the
scaling efficiencies computed from the per-N values.  python scripts/bench_table.py [prefix]"""
import json
import os
import sys

prefix = sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_bench_n"
rows = {}
for n in (1, 2, 4, 8):
    p = "%s%d.json" % (prefix, n)
    if os.path.exists(p):
        rows[n] = json.loads(open(p).read().strip().splitlines()[-1])
ns = sorted(rows)


def line(name, fn, fmt="%.3g"):
    vals = []
    for n in ns:
        try:
            v = fn(rows[n])
            vals.append(fmt % v if v is not None else "—")
        except (KeyError, TypeError):
            vals.append("—")
    print("| %s | %s |" % (name, " | ".join(vals)))


print("| | " + " | ".join("N = %d" % n for n in ns) + " |")
print("|---|" + "---|" * len(ns))
v1 = rows[1]["value"]
line("RANSAC hypotheses/s, resident, one call batch per GPU (weak)", lambda d: d["value"] / 1e6, "%.2f M")
line("… efficiency vs N × (N = 1)", lambda d: d["value"] / (d["n_gpus"] * v1), "%.2f")
line("RANSAC hypotheses/s end to end (host buffers)", lambda d: d["e2e"]["value"] / 1e6, "%.2f M")
line("ONE 10 k-hypothesis call sharded over the GPUs: ms per call", lambda d: d["ransac_sharded_call"]["ms_per_call"], "%.2f")
s1 = rows[1]["ransac_sharded_call"]["ms_per_call"]
line("… speed-up over N = 1", lambda d: s1 / d["ransac_sharded_call"]["ms_per_call"], "%.2f×")
line("… all-reduce + pruning-bound update, ms per call", lambda d: d["ransac_sharded_call"]["allreduce_ms_per_call"], "%.3f")
line("… identical to the single-GPU call", lambda d: d["ransac_sharded_call"]["parity_vs_single_gpu_call"], "%s")
b1 = rows[1]["ba"]["value"]
line("BA config 4, LM iterations/s, resident (strong)", lambda d: d["ba"]["value"], "%.0f")
line("… ms per iteration", lambda d: d["ba"]["ms_per_iteration"], "%.2f")
line("… speed-up over N = 1", lambda d: d["ba"]["value"] / b1, "%.2f×")
for k, name in (("jacobian_build", "Jacobian build"), ("reduced_system_incl_allreduce", "reduced system (+ all-reduce)"),
                ("cholesky_solve", "Cholesky + triangular solves"), ("backsubstitution_and_candidate_cost", "back-substitution + candidate cost")):
    line("… " + name + ", ms", lambda d, k=k: d["ba"]["phase_ms_per_iteration"][k], "%.3f")
line("… sharded solve == single-GPU solve", lambda d: d["ba"]["parity_vs_single_gpu"]["ok"] if d["n_gpus"] > 1 else None, "%s")
line("BA end to end (10-iteration solve from pinned host arrays), it/s", lambda d: d["ba"]["e2e"]["value"], "%.0f")
o1 = rows[1]["ba"]["value"] * rows[1]["ba"]["config"]["observations"]
line("BA weak scaling (200 k points per GPU): observation·iterations/s", lambda d: (d["ba_weak_scaling"]["observation_iterations_per_s"] if d["n_gpus"] > 1 else o1) / 1e9, "%.2f G")
line("… efficiency vs N × (N = 1)", lambda d: (d["ba_weak_scaling"]["observation_iterations_per_s"] if d["n_gpus"] > 1 else o1) / (d["n_gpus"] * o1), "%.2f")
line("… ms per iteration", lambda d: d["ba_weak_scaling"]["ms_per_iteration"] if d["n_gpus"] > 1 else d["ba"]["ms_per_iteration"], "%.2f")
