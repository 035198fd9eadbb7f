#!/usr/bin/env python3
"""Compile the hidden-variable resultant of re3q3 into order-preserving three-address code.

Input (read at GENERATION time only, in the build container):
    /root/reference/lib/re3q3/re3q3/re3q3.h
      :84-137   the 33 resultant coefficients a11 ... a313 (and the temporaries t2 ... t20)
      :142-150  c(0) ... c(8) = det M(x)
      :177-188  A(x) for a root x and the two Cramer quotients for y, z
Output: the same arithmetic as straight-line three-address code in static-single-assignment form
    oracle/re3q3_resultant.inc                              (CPU oracle)
    privacy_preserving_sfm_b200/csrc/re3q3_resultant.inc    (device code)

The generator parses every right-hand side with C's precedence and associativity (products and
sums left to right, unary minus binds tighter than '*', parentheses honoured) and emits one
statement per binary operation in exactly that evaluation order, so every intermediate value
rounds as in the reference's FMA-free build.  Two value-preserving simplifications only:
  * common sub-expressions are computed once (same operation on the same operands gives the
    same IEEE result; '+' and '*' are commutative bit for bit);
  * a negated product / sum is a negated operand (IEEE negation is exact).
Nothing is re-associated or distributed.

Usage: python scripts/gen_re3q3_resultant.py [--check]
"""
import os
import re
import sys

REF = "/root/reference/lib/re3q3/re3q3/re3q3.h"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUTS = [os.path.join(ROOT, "oracle", "re3q3_resultant.inc"),
        os.path.join(ROOT, "privacy_preserving_sfm_b200", "csrc", "re3q3_resultant.inc")]

TOKEN = re.compile(r"\s*(?:(\d+\.\d*|\d+)|([A-Za-z_]\w*(?:\(\d(?:,\d)?\))?)|(.))")


def tokenize(s):
    out = []
    pos = 0
    while pos < len(s):
        m = TOKEN.match(s, pos)
        if not m:
            break
        pos = m.end()
        if m.group(1):
            out.append(("num", m.group(1)))
        elif m.group(2):
            out.append(("id", m.group(2)))
        elif m.group(3).strip():
            out.append(("op", m.group(3)))
    return out


class Parser:
    """expr := term (('+'|'-') term)* ; term := unary ('*' unary)* ; unary := '-' unary | atom"""

    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else (None, None)

    def take(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def expr(self):
        node = self.term()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.take()[1]
            node = (op, node, self.term())
        return node

    def term(self):
        node = self.unary()
        while self.peek() == ("op", "*"):
            self.take()
            node = ("*", node, self.unary())
        return node

    def unary(self):
        if self.peek() == ("op", "-"):
            self.take()
            return ("neg", self.unary())
        return self.atom()

    def atom(self):
        kind, val = self.take()
        if kind == "num":
            return ("num", float(val))
        if kind == "id":
            return ("id", val)
        assert (kind, val) == ("op", "("), (kind, val)
        node = self.expr()
        assert self.take() == ("op", ")")
        return node


class Emitter:
    """SSA three-address code; values are (name, negated) pairs so that negation costs nothing."""

    def __init__(self):
        self.lines = []
        self.cse = {}
        self.env = {}
        self.n = 0

    def leaf(self, name):
        m = re.fullmatch(r"P\((\d),(\d)\)", name)
        if m:
            return ("P[%s][%s]" % m.groups(), False)
        if name in self.env:
            return self.env[name]
        raise KeyError(name)

    @staticmethod
    def text(v):
        return ("-" if v[1] else "") + v[0]

    def binop(self, op, a, b):
        # normalise signs: every temporary holds a non-negated operation result
        if op == "*":
            neg = a[1] != b[1]
            x, y = sorted([a[0], b[0]])
            key = ("*", x, y)
            expr = "%s * %s" % (x, y)
        else:
            if op == "-":
                b = (b[0], not b[1])
            # a + b with signs
            if a[1] and b[1]:
                neg, key, expr = True, ("+",) + tuple(sorted([a[0], b[0]])), None
                x, y = sorted([a[0], b[0]])
                expr = "%s + %s" % (x, y)
            elif not a[1] and not b[1]:
                neg = False
                x, y = sorted([a[0], b[0]])
                key, expr = ("+", x, y), "%s + %s" % (x, y)
            elif not a[1] and b[1]:
                neg, key, expr = False, ("-", a[0], b[0]), "%s - %s" % (a[0], b[0])
            else:  # -a + b = b - a (exactly)
                neg, key, expr = False, ("-", b[0], a[0]), "%s - %s" % (b[0], a[0])
        if key not in self.cse:
            name = "v%d" % self.n
            self.n += 1
            self.lines.append("const Real %s = %s;" % (name, expr))
            self.cse[key] = name
        return (self.cse[key], neg)

    def gen(self, node):
        kind = node[0]
        if kind == "id":
            return self.leaf(node[1])
        if kind == "num":
            lit = repr(node[1])
            return (lit, False)
        if kind == "neg":
            v = self.gen(node[1])
            return (v[0], not v[1])
        a = self.gen(node[1])
        b = self.gen(node[2])
        return self.binop(kind, a, b)


def parse_assignment(line):
    m = re.match(r"\s*(?:double\s+)?([A-Za-z_]\w*(?:\(\d\))?)\s*=\s*(.*);\s*$", line)
    assert m, line
    return m.group(1), Parser(tokenize(m.group(2))).expr()


A_NAMES = (["a1%d" % k for k in range(1, 11)] + ["a2%d" % k for k in range(1, 11)] +
           ["a3%d" % k for k in range(1, 14)])


def generate():
    src = open(REF).read().split("\n")
    body = [l for l in src[83:150] if "=" in l and not l.strip().startswith("//")
            and "Eigen::" not in l]
    em = Emitter()
    out = []
    out.append("// GENERATED by scripts/gen_re3q3_resultant.py from lib/re3q3/re3q3/re3q3.h:84-150 and")
    out.append("// :177-188 of the reference -- do not edit.  Every statement is one IEEE operation of the")
    out.append("// reference's expressions, in the reference's evaluation order (C precedence, left to")
    out.append("// right), with common sub-expressions computed once; nothing is re-associated.")
    out.append("// a[0..9] = a11..a110, a[10..19] = a21..a210, a[20..32] = a31..a313; c[k] = c(k).")
    out.append("template <typename Real>")
    out.append("RE3Q3_FN void re3q3_resultant(const Real (&P)[3][7], Real (&a)[33], Real (&c)[9]) {")
    start = len(em.lines)
    for line in body:
        name, ast = parse_assignment(line)
        v = em.gen(ast)
        m = re.fullmatch(r"c\((\d)\)", name)
        if m:
            em.lines.append("c[%s] = %s;" % (m.group(1), em.text(v)))
        elif name in A_NAMES:
            em.lines.append("a[%d] = %s;" % (A_NAMES.index(name), em.text(v)))
            em.env[name] = ("a[%d]" % A_NAMES.index(name), False)
        else:  # t2 ... t20
            em.env[name] = v
    out += ["  " + l for l in em.lines[start:]]
    out.append("}")
    out.append("")
    # ---- back-substitution for one real root (re3q3.h:177-188)
    em2 = Emitter()
    for i, nm in enumerate(A_NAMES):
        em2.env[nm] = ("a[%d]" % i, False)
    for nm in ("xs1", "xs2", "xs3", "xs4"):
        em2.env[nm] = (nm, False)
    first = next(i for i, l in enumerate(src) if l.strip().startswith("A << a11"))
    rows = " ".join(src[first:first + 3])
    rows = rows[rows.index("<<") + 2:rows.rindex(";")]
    entries = [e.strip() for e in rows.split(",")]
    assert len(entries) == 9, entries
    out.append("// y, z for the root x = xs1 from the first two rows of M(x) (2x2 Cramer).")
    out.append("template <typename Real>")
    out.append("RE3Q3_FN void re3q3_backsubstitute(const Real (&a)[33], const Real xs1, Real* y, Real* z) {")
    out.append("  const Real xs2 = xs1 * xs1;")
    out.append("  const Real xs3 = xs1 * xs2;")
    names = {}
    for r in range(2):          # only A(0,*) and A(1,*) are read by :182-183
        for cidx in range(3):
            v = em2.gen(Parser(tokenize(entries[3 * r + cidx])).expr())
            names["A(%d,%d)" % (r, cidx)] = v
    for k, v in names.items():
        em2.env[k] = v
    for target, key in (("*y", "(*solutions)(1, root_cnt)"), ("*z", "(*solutions)(2, root_cnt)")):
        line = next(l for l in src if l.strip().startswith(key))
        rhs = line[line.index("=") + 1:line.rindex(";")]
        rhs = re.sub(r"A\((\d),(\d)\)", lambda m: "A%s%s" % m.groups(), rhs)
        for k, v in list(names.items()):
            em2.env["A%s%s" % (k[2], k[4])] = v
        # quotient of two differences: parse numerator / denominator separately
        num, den = rhs.split("/")
        vn = em2.gen(Parser(tokenize(num)).expr())
        vd = em2.gen(Parser(tokenize(den)).expr())
        em2.lines.append("%s = %s / %s;" % (target, em2.text(vn), em2.text(vd)))
    out += ["  " + l for l in em2.lines]
    out.append("}")
    return "\n".join(out) + "\n"


def main():
    text = generate()
    if "--check" in sys.argv:
        bad = [p for p in OUTS if not os.path.exists(p) or open(p).read() != text]
        if bad:
            print("stale:", bad)
            sys.exit(1)
        print("up to date")
        return
    for p in OUTS:
        with open(p, "w") as f:
            f.write(text)
        print("wrote", p, len(text.split("\n")), "lines")


if __name__ == "__main__":
    main()
