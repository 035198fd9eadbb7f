#!/usr/bin/env python3
"""Golden vectors of the re3q3 resultant, computed from the REFERENCE'S OWN SOURCE TEXT.

Runs in the build container only (reads /root/reference/lib/re3q3/re3q3/re3q3.h): the
assignments of lines 84-150 (a11 ... a313, t2 ... t20, c(0) ... c(8)) and the A(x) / Cramer
expressions of lines 177-188 are executed as Python statements — Python floats are IEEE
doubles and Python's `+ - * /` have C's precedence and left-to-right associativity, no FMA —
so the numbers below are what an FMA-free build of the reference computes for these lines.
tests/test_golden.py requires the oracle (and, through the oracle, the CUDA solver) to
reproduce them bit for bit.

Usage: python tests/golden/make_re3q3_golden.py   -> tests/golden/re3q3_resultant_vectors.json
"""
import json
import os
import re

import numpy as np

REF = "/root/reference/lib/re3q3/re3q3/re3q3.h"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "re3q3_resultant_vectors.json")


class Mat:
    """P(i,j) / c(k) / A(i,j) call syntax on nested lists."""

    def __init__(self, rows, cols=None):
        self.v = [[0.0] * (cols or 1) for _ in range(rows)]

    def __call__(self, i, j=0):
        return self.v[i][j]


def main():
    src = open(REF).read().split("\n")
    stmts = []
    for line in src[83:150]:
        m = re.match(r"\s*(?:double\s+)?([a-z]\w*|c\((\d)\))\s*=\s*(.*);\s*$", line)
        if not m:
            continue
        target = "cvals[%s]" % m.group(2) if m.group(2) else m.group(1)
        stmts.append("%s = %s" % (target, m.group(3)))
    first = next(i for i, l in enumerate(src) if l.strip().startswith("A << a11"))
    rows = " ".join(src[first:first + 3])
    entries = [e.strip() for e in rows[rows.index("<<") + 2:rows.rindex(";")].split(",")]
    y_rhs = next(l for l in src if l.strip().startswith("(*solutions)(1, root_cnt)"))
    z_rhs = next(l for l in src if l.strip().startswith("(*solutions)(2, root_cnt)"))
    y_rhs = y_rhs[y_rhs.index("=") + 1:y_rhs.rindex(";")]
    z_rhs = z_rhs[z_rhs.index("=") + 1:z_rhs.rindex(";")]
    program = compile("\n".join(stmts), "re3q3.h:84-150", "exec")

    names = (["a1%d" % k for k in range(1, 11)] + ["a2%d" % k for k in range(1, 11)] +
             ["a3%d" % k for k in range(1, 14)])
    rng = np.random.default_rng(20201017)
    cases = []
    for case in range(24):
        scale = [1.0, 1.0, 1e-3, 1e3, 1.0, 7.0][case % 6]
        Pv = rng.standard_normal((3, 7)) * scale
        if case % 5 == 4:
            Pv[rng.integers(3), rng.integers(7)] = 0.0
        P = Mat(3, 7)
        P.v = [[float(x) for x in row] for row in Pv]
        env = {"P": P, "cvals": [0.0] * 9}
        exec(program, env)
        a = [env[n] for n in names]
        roots = []
        for x in rng.standard_normal(3) * 2.0:
            xs1 = float(x)
            env2 = dict(env, xs1=xs1, xs2=xs1 * xs1)
            env2["xs3"] = xs1 * env2["xs2"]
            env2["xs4"] = xs1 * env2["xs3"]
            A = Mat(3, 3)
            A.v = [[eval(entries[3 * r + c], env2) for c in range(3)] for r in range(3)]
            env2["A"] = A
            roots.append({"x": xs1.hex(), "y": float(eval(y_rhs, env2)).hex(),
                          "z": float(eval(z_rhs, env2)).hex()})
        cases.append({"P": [[x.hex() for x in row] for row in P.v],
                      "a": [float(x).hex() for x in a],
                      "c": [float(x).hex() for x in env["cvals"]],
                      "roots": roots})
    with open(OUT, "w") as f:
        json.dump({"source": "lib/re3q3/re3q3/re3q3.h:84-150,177-188 evaluated in IEEE double",
                   "cases": cases}, f, indent=0)
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
