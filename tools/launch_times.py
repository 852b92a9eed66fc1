#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  python tools/launch_times.py launches.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].replace("void ", "").replace("flint::", "")[:100]
    d = agg.setdefault(name, [0, 0.0]); d[0] += 1; d[1] += float(r[-1]) / 1e6
tot = sum(v[1] for v in agg.values())
for k, v in agg.items(): print("%-100s %3d launches %9.3f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
