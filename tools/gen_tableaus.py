#!/usr/bin/env python3
"""Generate the Butcher-tableau data headers from the reference's coefficient file.

Reads  /root/reference/src/ButcherTableaus.f90  (never copied into this repo) and
evaluates every `parameter` array exactly the way a Fortran compiler would in
working precision real64:

  * each decimal literal is rounded to the nearest double (Python float()),
  * `a/b` ratios are one correctly-rounded double division,
  * `NAME(lo:hi)` slices and `B - BH` array expressions are element-wise doubles,
  * `(i, i=a,b)` implied-do lists are expanded,
  * `reshape(src, [r,c])` is column-major and TRUNCATES surplus source data, as
    Fortran does.  This is what reproduces reference quirk Q2:
    Verner65E_d is declared with 6 columns of data but reshaped to [8,5], so
    the theta^6 column is silently dropped (src/ButcherTableaus.f90:975,1095-1152).

Output: C arrays with hex-float literals (bit-exact), 0-based, written to
  oracle/oracle_tableaus.h           (oracle; same numbers, separate file so the
                                      product never includes anything under oracle/)

Self-checks: row sums of a equal c, sum(b)=1, the FSAL
row equals b, sum_j d_ij = b_i for Verner98R; non-zero counts per method.

Run:  python tools/gen_tableaus.py [--ref /root/reference]
"""
import argparse
import os
import re
import sys

METHODS = ["DOP54", "DOP853", "Verner98R", "Verner65E"]


def load_statements(path):
    """Strip comments, join continuation lines, return list of logical statements."""
    stmts, cur = [], ""
    with open(path) as f:
        for raw in f:
            line = raw.split("!")[0].rstrip()  # no '!' inside strings in this file
            if not line.strip():
                continue
            s = line.strip()
            if s.startswith("&"):
                s = s[1:].lstrip()
            if s.endswith("&"):
                cur += s[:-1] + " "
                continue
            cur += s
            stmts.append(cur)
            cur = ""
    return stmts


_num = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eEdD][+-]?\d+)?(_[wW][pP])?$")


def split_top(s, sep=","):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


class Evaluator:
    def __init__(self):
        self.sym = {}  # name(lower) -> python scalar (int/float/bool) or list

    def scalar(self, tok):
        tok = tok.strip()
        if _num.match(tok):
            t = re.sub(r"_[wW][pP]$", "", tok).replace("d", "e").replace("D", "e")
            if re.match(r"^[+-]?\d+$", t):
                return int(t)
            return float(t)
        m = re.match(r"^(.+?)/(.+)$", tok)
        if m and _num.match(m.group(1).strip()) and _num.match(m.group(2).strip()):
            return float(self.scalar(m.group(1))) / float(self.scalar(m.group(2)))
        low = tok.lower()
        if low in self.sym and not isinstance(self.sym[low], list):
            return self.sym[low]
        m = re.match(r"^size\((\w+)\)$", tok, re.I)
        if m:
            return len(self.sym[m.group(1).lower()])
        raise ValueError("cannot evaluate scalar: %r" % tok)

    def array_items(self, body):
        """body = contents between [ and ]."""
        vals = []
        for item in split_top(body):
            m = re.match(r"^\(\s*(\w+)\s*,\s*(\w+)\s*=\s*(\d+)\s*,\s*(\d+)\s*\)$", item)
            if m:  # implied do (i, i=a,b)
                assert m.group(1) == m.group(2)
                vals.extend(range(int(m.group(3)), int(m.group(4)) + 1))
                continue
            m = re.match(r"^(\w+)\((\d+):(\d+)\)$", item)
            if m:  # slice of a previously defined 1-based array
                src = self.sym[m.group(1).lower()]
                vals.extend(src[int(m.group(2)) - 1:int(m.group(3))])
                continue
            vals.append(self.scalar(item))
        return vals

    def rhs(self, expr):
        expr = expr.strip()
        m = re.match(r"^reshape\(\s*\[(.*)\]\s*,\s*\[(.*)\]\s*\)$", expr, re.I | re.S)
        if m:
            data = self.array_items(m.group(1))
            shape = [int(self.scalar(t)) for t in split_top(m.group(2))]
            need = shape[0] * shape[1]
            if len(data) < need:
                raise ValueError("reshape source too small")
            return {"data": data[:need], "shape": shape, "surplus": len(data) - need}
        if expr.startswith("[") and expr.endswith("]"):
            return self.array_items(expr[1:-1])
        m = re.match(r"^(\w+)\s*-\s*(\w+)$", expr)
        if m and m.group(1).lower() in self.sym:
            a, b = self.sym[m.group(1).lower()], self.sym[m.group(2).lower()]
            return [x - y for x, y in zip(a, b)]
        if expr.upper() in (".TRUE.", ".FALSE."):
            return expr.upper() == ".TRUE."
        return self.scalar(expr)

    def feed(self, stmt):
        m = re.match(r"^(real\(WP\)|integer|logical)\s*,\s*parameter\s*::\s*(\w+)\s*(\([^=]*\))?\s*=\s*(.*)$",
                     stmt, re.I | re.S)
        if not m:
            return
        name = m.group(2).lower()
        try:
            self.sym[name] = self.rhs(m.group(4))
        except (ValueError, KeyError) as e:
            if any(name.startswith(p.lower()) for p in METHODS):
                raise
            # parameters unrelated to the four tableaus (ERK_MAX_A etc.)
            sys.stderr.write("skip %s: %s\n" % (name, e))


def hexf(x):
    return float(x).hex()


def emit_array(name, vals, ctype="double"):
    body = ",\n    ".join(", ".join(hexf(v) if ctype == "double" else str(v) for v in vals[i:i + 3 if ctype == "double" else i + 16])
                          for i in range(0, len(vals), 3 if ctype == "double" else 16))
    return "FLINT_TABLEAU_CONST %s %s[%d] = {\n    %s\n};\n" % (ctype, name, len(vals), body)


def load_symbols(ref):
    """Evaluate every tableau parameter of the reference file; returns {lowercase name: value}."""
    ev = Evaluator()
    for st in load_statements(os.path.join(ref, "src", "ButcherTableaus.f90")):
        ev.feed(st)
    return ev.sym


def build(ref):
    S = load_symbols(ref)
    out = []
    report = []
    for M in METHODS:
        p = M.lower()
        s, sint = S[p + "_s"], S[p + "_sint"]
        a, b, c, e = S[p + "_a"], S[p + "_b"], S[p + "_c"], S[p + "_e"]
        d = S[p + "_d"]
        dinz = S[p + "_di_nz"]
        assert len(a) == s * (s - 1) // 2, (M, len(a))
        assert len(c) == s - 1 and len(b) == sint and len(e) == sint
        # ---- self checks -------------------------------------------------
        worst = 0.0
        for i in range(2, s + 1):
            row = a[(i - 1) * (i - 2) // 2:(i - 1) * (i - 2) // 2 + i - 1]
            worst = max(worst, abs(sum(row) - c[i - 2]))
        assert worst < 1e-13, (M, worst)
        assert abs(sum(b) - 1.0) < 1e-12, (M, sum(b))
        if S[p + "_fsal"]:
            row = a[(sint - 1) * (sint - 2) // 2:(sint - 1) * (sint - 2) // 2 + sint - 1]
            assert row == b[:sint - 1], M
        nz_int = sum(1 for i in range(2, sint + 1) for j in range(1, i)
                     if a[(i - 1) * (i - 2) // 2 + j - 1] != 0.0)
        nz_ext = sum(1 for i in range(sint + 1, s + 1) for j in range(1, i)
                     if a[(i - 1) * (i - 2) // 2 + j - 1] != 0.0)
        report.append("%-10s s=%d sint=%d nnz(a int)=%d nnz(a ext)=%d nnz(b)=%d nnz(e)=%d rowsum err=%.1e d%s surplus=%d" % (
            M, s, sint, nz_int, nz_ext, sum(1 for x in b if x), sum(1 for x in e if x), worst,
            d["shape"], d["surplus"]))
        # ---- emit --------------------------------------------------------
        U = M.upper()
        out.append("/* ---- %s (src/ButcherTableaus.f90) ---- */\n" % M)
        for key in ("q", "p", "phat", "pstar", "s", "sint"):
            out.append("#define FLINT_%s_%s %d\n" % (U, key.upper(), S[p + "_" + key]))
        out.append("#define FLINT_%s_FSAL %d\n" % (U, 1 if S[p + "_fsal"] else 0))
        for key in ("beta", "beta_mult"):
            if p + "_" + key in S:
                out.append("#define FLINT_%s_%s %s\n" % (U, key.upper(), hexf(S[p + "_" + key])))
        out.append(emit_array("FLINT_%s_A" % U, a))
        out.append(emit_array("FLINT_%s_B" % U, b))
        out.append("/* c[i-2] is c_i for stage i = 2..s */\n")
        out.append(emit_array("FLINT_%s_C" % U, c))
        out.append(emit_array("FLINT_%s_E" % U, e))
        if p + "_bhh" in S:
            out.append(emit_array("FLINT_%s_BHH" % U, S[p + "_bhh"]))
        out.append(emit_array("FLINT_%s_DINZ" % U, dinz, "int"))
        out.append("#define FLINT_%s_D_ROWS %d\n#define FLINT_%s_D_COLS %d\n" % (U, d["shape"][0], U, d["shape"][1]))
        out.append("/* column-major as in Fortran: D[row + D_ROWS*col]; row indexes DINZ */\n")
        out.append(emit_array("FLINT_%s_D" % U, d["data"]))
        out.append("\n")
    # Verner98R dense consistency: sum_j d_ij = b_i (theta=1 reproduces the step)
    d = S["verner98r_d"]
    rows, cols = d["shape"]
    b = S["verner98r_b"]
    dinz = S["verner98r_di_nz"]
    worst = 0.0
    for r in range(rows):
        st = dinz[r]
        bi = b[st - 1] if st <= len(b) else 0.0
        worst = max(worst, abs(sum(d["data"][r + rows * cc] for cc in range(cols)) - bi))
    assert worst < 1e-11, worst
    report.append("Verner98R sum_j d_ij - b_i max = %.2e" % worst)
    # quirk Q2 evidence
    d = S["verner65e_d"]
    rows, cols = d["shape"]
    q2 = sum(d["data"][0 + rows * cc] for cc in range(cols)) - S["verner65e_b"][0]
    report.append("Verner65E sum_j d_1j - b_1 = %.6f (quirk Q2: theta^6 column truncated by reshape)" % q2)
    return "".join(out), report


HEADER = """/* GENERATED by tools/gen_tableaus.py from the reference's src/ButcherTableaus.f90.
 * DO NOT EDIT.  Numerical constants only (hex floats are the exact real64 values a
 * Fortran compiler produces for the reference's `parameter` expressions).
 * Layout: A is the packed lower triangle a21,a31,a32,... (index (i-1)(i-2)/2 + j - 1
 * for 1-based stage i, column j; src/ERKStepInt.f90:162).
 */
#ifndef %(guard)s
#define %(guard)s
#ifndef FLINT_TABLEAU_CONST
#define FLINT_TABLEAU_CONST static const
#endif

"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--root", default=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    args = ap.parse_args()
    body, report = build(args.ref)
    # the product's coefficients are generated into erk_methods_gen.cuh by tools/gen_kernels.py (which imports this parser)
    for rel, guard in (("oracle/oracle_tableaus.h", "FLINT_ORACLE_TABLEAUS_H"),):
        path = os.path.join(args.root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(HEADER % {"guard": guard})
            f.write(body)
            f.write("#endif\n")
        print("wrote", path)
    print("\n".join(report))


if __name__ == "__main__":
    main()
