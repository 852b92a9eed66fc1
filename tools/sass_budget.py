#!/usr/bin/env python3
"""Per-attempt instruction budget of a kernel by SOURCE CATEGORY: joins ncu's per-SASS-instruction execution counts
(source page of a --set full --import-source on report) with the line / inline-chain information of the same binary
(nvdisasm --print-line-info-inline on the object the run used).

    python tools/sass_budget.py REP.ncu-rep OBJECT.o KERNEL_SUBSTRING [--kernel-index K] [--per N_WARP_ATTEMPTS]

Every executed warp instruction is attributed to a category from its inline chain (innermost frame first):
    stage sums            MethodXXX::step: a_ij k_j sums, y0 + h*acc                       (erk_methods_gen.cuh, stage blocks)
    error estimate        MethodXXX::step: error sums, Sc, norm                             (erk_methods_gen.cuh, after the last stage)
    RHS arithmetic        SYS::accel / assemble: the right-hand side's own operations        (systems.cuh)
    RHS div/sqrt          the branch-free IEEE division / square-root sequences of the RHS   (arith.cuh, called from systems.cuh)
    RHS range tests       fast-path range tests of the RHS operands                          (arith.cuh small*/moderate*)
    error-norm div/sqrt   divisions / square root of the error estimate                      (arith.cuh, called from step)
    controller pow        det_pow_dev                                                        (arith.cuh)
    controller / commit   StepSzHairer, accept / reject, commit_tail, last-step clamp         (erk_kernel.cuh)
    events                event functions + DetectEvent on the hot path                      (systems.cuh events, erk_kernel.cuh detect_event)
    state I/O             state_load / state_store / store_results / refill                  (erk_kernel.cuh)
    other
and split into fp64-pipe instructions (DADD/DMUL/DFMA/DSETP/...) and the rest.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, ksub = sys.argv[1:4]
kidx = int(sys.argv[sys.argv.index("--kernel-index") + 1]) if "--kernel-index" in sys.argv else 0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# ---------------------------------------------------------------- function line ranges of the sources
def functions(path):
    """[(first line, name)] of function-like definitions, by a crude scan (enough to name the frames of an inline chain)."""
    out = []
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+|inline\s+|__host__\s+|__device__\s+|__forceinline__\s+|__noinline__\s+|constexpr\s+)*"
                     r"[\w:<>,\s\*&]*?\b(\w+)\s*\([^;{]*$|^\s*(?:static\s+)?__device__[^;(]*\b(\w+)\s*\(")
    for n, line in enumerate(open(path), 1):
        if "__device__" in line or "__global__" in line:
            m = re.search(r"\b(\w+)\s*\(", line.split("__device__")[-1] if "__device__" in line else line)
            if m and m.group(1) not in ("__launch_bounds__", "if", "for", "while", "switch", "sizeof"):
                out.append((n, m.group(1)))
    return out


SRC = {}
for fn in ("arith.cuh", "systems.cuh", "erk_kernel.cuh", "erk_methods_gen.cuh", "det_pow.h", "interp_eval.cuh"):
    p = os.path.join(ROOT, "flint_b200", "csrc", fn)
    if os.path.exists(p):
        SRC[fn] = functions(p)


def func_at(fn, line):
    name = "?"
    for n, f in SRC.get(fn, ()):
        if n <= line:
            name = f
        else:
            break
    return name


def gen_region(line):
    """erk_methods_gen.cuh: 'stage' while inside a stage block, 'error' from the error sums on (per method)."""
    txt = GEN_LINES
    # walk back to the nearest marker
    for n in range(min(line, len(txt)) - 1, -1, -1):
        t = txt[n]
        if "/* stage " in t:
            return "stage"
        if "error" in t.lower() and ("/*" in t or "//" in t):
            return "error"
        if "static __device__" in t:
            return "stage"
    return "stage"


GEN_LINES = open(os.path.join(ROOT, "flint_b200", "csrc", "erk_methods_gen.cuh")).read().splitlines()


def category(chain):
    """chain: [(file, line)] innermost first."""
    frames = [(os.path.basename(f), l, func_at(os.path.basename(f), l)) for f, l in chain]
    names = [fr[2] for fr in frames]
    files = [fr[0] for fr in frames]
    inner_file, inner_line, inner_fn = frames[0]
    def within(fn_names):
        return any(n in fn_names for n in names)
    in_rhs = within(("accel_impl", "accel", "accel_lazy", "assemble", "eval_rhs", "rhs_eval", "accel_call", "cr3bp_quotients_plain", "twobody_fac_plain"))
    in_step = "erk_methods_gen.cuh" in files
    if within(("det_pow_dev", "det_pow_constants", "flint_det_pow")) or "det_pow.h" in files:
        return "controller pow"
    if within(("events", "detect_event")):
        return "events"
    if within(("state_load", "state_store", "store_results", "log_dense_step", "log_dense_event_step")):
        return "state I/O"
    if inner_file == "arith.cuh":
        if re.match(r"small\d+|moderate|moderate_sq|in_range", inner_fn):
            return "RHS range tests" if in_rhs else "error-norm div/sqrt"
        if re.search(r"sqrt|rcp|div", inner_fn):
            return "RHS div/sqrt" if in_rhs else ("error-norm div/sqrt" if in_step else "controller / commit")
    if in_rhs:
        return "RHS arithmetic"
    if in_step:
        gl = [l for f, l, _ in frames if f == "erk_methods_gen.cuh"][0]
        return "stage sums" if gen_region(gl) == "stage" else "error estimate"
    if within(("commit_tail",)):
        return "controller / commit"
    if inner_file == "erk_kernel.cuh" or "erk_kernel.cuh" in files:
        return "controller / commit"
    return "other"


# ---------------------------------------------------------------- disassembly with line info
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
    cubins = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")]
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", cubins[0]], capture_output=True, text=True).stdout.splitlines()
start = None
for n, l in enumerate(dis):
    if l.startswith(".text.") and ksub in l:
        start = n
        break
if start is None:
    sys.exit("kernel not found in the object: " + ksub)
addr_chain = {}
chain, fresh = [], True
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        addr_chain[int(m.group(1), 16)] = (list(chain), m.group(2).strip())
        fresh = True

# ---------------------------------------------------------------- executed counts from the ncu report
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
blk = rows[starts[kidx]:starts[kidx + 1]]
hdr, data = blk[1], blk[2:]
ia, isrc, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
base = int(data[0][ia], 16)
FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")
tot = collections.Counter()
mism = 0
ops_by_cat = collections.defaultdict(collections.Counter)
for r in data:
    a = int(r[ia], 16) - base
    ex = int(r[iex])
    if not ex:
        continue
    text = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
    op = text.split()[0].split(".")[0]
    ch = addr_chain.get(a)
    if ch is None or not ch[0]:
        cat = "other"
    else:
        if ch[1].split()[-0:1] and op not in ch[1]:
            mism += 1
        cat = category(ch[0])
    kind = "fp64" if op in FP64 else "other"
    tot[(cat, kind)] += ex
    ops_by_cat[cat][op] += ex
total = sum(tot.values())
per = float(sys.argv[sys.argv.index("--per") + 1]) if "--per" in sys.argv else None
if per is None:
    # warp-attempts: the most common per-instruction count inside the stage code
    cnt = collections.Counter(int(r[iex]) for r in data if int(r[iex]) > 0)
    per = float(cnt.most_common(1)[0][0])
print("kernel:", blk[0][1][:110])
print("warp instructions executed: %d; per-attempt unit = %.0f warp-attempts (most frequent per-instruction count)%s" % (
    total, per, "; %d address/opcode mismatches (binary differs from the report?)" % mism if mism else ""))
cats = sorted({c for c, _ in tot}, key=lambda c: -(tot[(c, "fp64")] + tot[(c, "other")]))
print("\n| category | fp64 instr / attempt | other instr / attempt | share of issue slots (fp64 x 2) | top opcodes |")
print("|---|---:|---:|---:|---|")
slots = sum(v * (2 if k == "fp64" else 1) for (c, k), v in tot.items())
sf = so = 0.0
for c in cats:
    f, o = tot[(c, "fp64")] / per, tot[(c, "other")] / per
    sf += f; so += o
    top = ", ".join("%s %.0f" % (k, v / per) for k, v in ops_by_cat[c].most_common(5))
    print("| %s | %.0f | %.0f | %.1f %% | %s |" % (c, f, o, 100 * (2 * tot[(c, "fp64")] + tot[(c, "other")]) / slots, top))
print("| **total** | **%.0f** | **%.0f** | 100 %% | |" % (sf, so))
if "--lines" in sys.argv:          # per source line of one category:  --lines "controller / commit"
    want = sys.argv[sys.argv.index("--lines") + 1]
    by_line = collections.defaultdict(collections.Counter)
    for r in data:
        a = int(r[ia], 16) - base
        ex = int(r[iex])
        ch = addr_chain.get(a)
        if not ex or ch is None or not ch[0] or category(ch[0]) != want:
            continue
        text = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip())
        f, l = ch[0][0]
        by_line[(os.path.basename(f), l)][text.split()[0].split(".")[0]] += ex
    print("\n%s, per source line (instructions per attempt):" % want)
    for (f, l), c in sorted(by_line.items(), key=lambda kv: -sum(kv[1].values()))[:45]:
        src = open(os.path.join(ROOT, "flint_b200", "csrc", f)).read().splitlines()[l - 1].strip()[:90] if os.path.exists(os.path.join(ROOT, "flint_b200", "csrc", f)) else ""
        print("  %5.1f  %s:%d  %s   | %s" % (sum(c.values()) / per, f, l, ", ".join("%s %.1f" % (k, v / per) for k, v in c.most_common(4)), src))
