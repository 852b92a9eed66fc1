#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the ERK path into small committed fixtures.

Sources (read-only, /root/reference/tests): the checked-in result files the reference's benchmark
programs wrote.  /root/reference does not exist on the GPU box, so the tests read
tests/golden/reference_goldens.json instead; this script is how that file was made.

    python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = {"source": "princemahajan/FLINT tests/*.txt (ifort builds); see SURVEY.md 8c for which cells are rounding-robust"}

wp = {}
for name in ("DOP54", "DOP853", "Verner65E", "Verner98R"):
    for suffix in ("", "_QP"):
        rows = open(os.path.join(ref, "tests", "wp_%s%s.txt" % (name, suffix))).read().strip().split("\n")[1:]
        wp[name + suffix] = [{"rtol_exp": i + 1, "fcalls": int(r.split()[1]), "err_sparse": float(r.split()[3]),
                              "err_dense": float(r.split()[5])} for i, r in enumerate(rows)]
out["work_precision"] = wp

lines = open(os.path.join(ref, "tests", "results_CR3BP.txt")).read().split("\n")
cr = {"sparse": {}, "dense": {}, "events": []}
section = None
for ln in lines:
    t = ln.split()
    if ln.strip().startswith("A."):
        section = "sparse"
    elif ln.strip().startswith("B."):
        section = "dense"
    elif len(t) == 7 and t[0] in ("DOP54", "DOP853", "Verner65E", "Verner98R"):
        cr[section][t[0]] = {"time_s": float(t[1]), "closing_err": float(t[2]), "jacobi_err": float(t[3]),
                             "fcalls": int(t[4]), "accepted": int(t[5]), "rejected": int(t[6])}
    elif len(t) >= 2 and t[0] == "2" and "E" in ln:
        s = ln.replace("E+00-", "E+00 -").replace("E-14 ", "E-14 ").split()
        cr["events"].append({"id": int(s[0]), "x": float(s[1]), "y1": float(s[2])})
out["results_CR3BP"] = cr

lz = {"sparse": {}, "dense": {}}
section = None
for ln in open(os.path.join(ref, "tests", "results_Lorenz.txt")).read().split("\n"):
    t = ln.split()
    if ln.strip().startswith("A."):
        section = "sparse"
    elif ln.strip().startswith("B."):
        section = "dense"
    elif len(t) == 6 and t[0] in ("DOP54", "DOP853", "Verner65E", "Verner98R"):
        lz[section][t[0]] = {"fcalls": int(t[2]), "accepted": int(t[3]), "rejected": int(t[4]), "total": int(t[5])}
out["results_Lorenz"] = lz

tb = {}
for ln in open(os.path.join(ref, "tests", "perf_TBP.txt")).read().split("\n"):
    t = ln.split()
    if len(t) == 7 and t[0] in ("DOP54", "DOP853", "Verner65E", "Verner98R"):
        tb[t[0]] = {"fcalls": int(t[4]), "accepted": int(t[5]), "rejected": int(t[6])}
out["perf_TBP"] = tb

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
json.dump(out, open(path, "w"), indent=1)
print("wrote", path)
