#!/usr/bin/env python3
"""Summarise ncu captures into profiles/: launch list shares + key metrics + stall mix of the integrate kernel.
    python tools/ncu_summary.py <launches.csv> <full.ncu-rep> <out.md> [traffic_key]"""
import csv, json, os, subprocess, sys, collections

launch_csv, rep, out_md = sys.argv[1:4]
tkey = sys.argv[4] if len(sys.argv) > 4 else None
L = []
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10 and r[0] != "ID"]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")[:90]
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1; d[1] += float(r[-1]) / 1e6
tot = sum(v[1] for v in agg.values())
L.append("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised: compare shares)\n")
L.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    L.append("| `%s` | %d | %.2f | %.1f %% |" % (k, v[0], v[1], 100 * v[1] / tot))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, data = rr[0], rr[1], rr[2]
g = lambda k: data[hdr.index(k)] if k in hdr else "n/a"
L.append("\n## Integrate kernel, `ncu --set full --clock-control none` (%s)\n" % g("Kernel Name")[:120])
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
L.append("| metric | value | unit |\n|---|---:|---|")
for k in keys:
    if k in hdr:
        L.append("| %s | %s | %s |" % (k, g(k), units[hdr.index(k)]))
L.append("\n### Warp stall mix (warps stalled per issue-active cycle)\n\n| reason | ratio |\n|---|---:|")
st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(data[i]))
      for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for n, v in sorted(st, key=lambda x: -x[1])[:9]:
    L.append("| %s | %.3f |" % (n, v))
open(out_md, "w").write("\n".join(L) + "\n")
print("\n".join(L))
if tkey:
    def tobytes(k):
        v, u = float(g(k)), units[hdr.index(k)].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    tp = os.path.join(os.path.dirname(out_md), "traffic.json")
    t = json.load(open(tp)) if os.path.exists(tp) else {}
    t[tkey] = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    t[tkey + "_source"] = os.path.basename(rep)
    json.dump(t, open(tp, "w"), indent=1)
