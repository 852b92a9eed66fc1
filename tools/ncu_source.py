#!/usr/bin/env python3
"""Dynamic SASS statistics from an ncu report's source page.
   python tools/ncu_source.py rep.ncu-rep [kernel-index] [--regions] [--top N] [--dump LO HI] [--mem]"""
import csv, sys, re, collections, subprocess
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[kidx]:starts[kidx + 1]]
print("kernel:", blk[0][1][:120])
hdr, data = blk[1], blk[2:]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(data[0][ia], 16)
recs = []
for r in data:
    a = int(r[ia], 16) - base; t = r[isrc].strip()
    recs.append((a, t, int(r[iex]), int(r[ismp]), r))
def opname(t):
    t2 = re.sub(r'^@!?U?P\d+\s+', '', t); op = t2.split()[0]
    if op.startswith('IMAD.MOV') or (op == 'IMAD.U32' and 'RZ, RZ' in t2): return 'IMAD.MOV'
    return op.split('.')[0]
h = collections.Counter(); hs = collections.Counter()
for a, t, ex, sm, r in recs: h[opname(t)] += ex; hs[opname(t)] += sm
tot = sum(h.values()); stot = sum(hs.values())
print("total warp inst", tot, "samples", stot)
for k, v in h.most_common(24): print('  %-10s %14d %5.1f%%   samples %5.1f%%' % (k, v, 100 * v / tot, 100 * hs[k] / stot))
if '--regions' in sys.argv:
    b = collections.Counter(); bs = collections.Counter()
    for a, t, ex, sm, r in recs: b[a // 2048] += ex; bs[a // 2048] += sm
    for k in sorted(b):
        if b[k] * 1000 > tot: print(hex(k * 2048), '%14d %5.1f%%  samples %5.1f%%' % (b[k], 100 * b[k] / tot, 100 * bs[k] / stot))
if '--top' in sys.argv:
    n = int(sys.argv[sys.argv.index('--top') + 1])
    for a, t, ex, sm, r in sorted(recs, key=lambda x: -x[3])[:n]:
        st = sorted(((int(r[i]), hn) for i, hn in stall_cols), reverse=True)[:2]
        print('%6x %-60s ex %6dM smp %6d  %s' % (a, t[:60], ex // 1000000, sm, st))
if '--mem' in sys.argv:
    for a, t, ex, sm, r in recs:
        if re.search(r'\b(LDL|STL)', t) and ex * 2000 > tot: print('%6x %-50s ex %6dM smp %6d' % (a, t[:50], ex // 1000000, sm))
if '--dump' in sys.argv:
    i = sys.argv.index('--dump'); lo, hi = int(sys.argv[i + 1], 16), int(sys.argv[i + 2], 16)
    for a, t, ex, sm, r in recs:
        if lo <= a < hi: print('%6x %-64s %6dM %6d' % (a, t[:64], ex // 1000000, sm))
if '--div' in sys.argv:
    ith = hdr.index("Thread Instructions Executed")
    lowi = lows = 0
    for a, t, ex, sm, r in recs:
        th = int(r[ith])
        if ex and th / ex < 8: lowi += ex; lows += sm
    print("instructions executed with < 8 active lanes: %.1f%% of warp instructions, %.1f%% of samples" % (100 * lowi / tot, 100 * lows / stot))
if '--hot' in sys.argv:
    mx = sorted(ex for a, t, ex, sm, r in recs)[-200]       # a typical per-attempt count
    hot = [(a, t, ex, sm) for a, t, ex, sm, r in recs if ex >= 0.25 * mx]
    print("per-attempt count ~%d; hot instructions (>= 25%% of it): %d = %.1f KB; 128B lines touched: %d = %.1f KB" % (
        mx, len(hot), len(hot) * 16 / 1024, len({a // 128 for a, _, _, _ in hot}), len({a // 128 for a, _, _, _ in hot}) / 8))
    hh = collections.Counter(opname(t) for a, t, ex, sm in hot)
    print("  hot static mix:", ", ".join("%s %d" % kv for kv in hh.most_common(16)))
    b = collections.Counter(a // 1024 for a, _, _, _ in hot)
    print("  hot KB map:", " ".join("%x:%d" % (k * 1024, v) for k, v in sorted(b.items())))
