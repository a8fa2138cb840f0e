#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    k = re.sub(r"\(.*", "", d["Kernel Name"])[:60]
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = d["Metric Unit"]
    ms = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    a = agg.setdefault(k, [0, 0.0, []])
    a[0] += 1
    a[1] += ms
    a[2].append(round(ms, 1))
tot = sum(a[1] for a in agg.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
print("%-60s %5s %10s %6s  per-launch ms" % ("kernel", "n", "total ms", "share"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-60s %5d %10.2f %5.1f%%  %s" % (k, a[0], a[1], 100 * a[1] / tot, a[2][:8] if a[0] <= 8 else ""))
print("total %.2f ms over %d launches" % (tot, sum(a[0] for a in agg.values())))
