#!/usr/bin/env python
"""Aggregate an ncu launch list (gpu__time_duration.sum CSV) per kernel: count and total ms.
usage: tools/prep_launches.py gpurun_out/launches_X.csv"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[h]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in rows[h + 1:]:
    if len(r) == len(hdr):
        k = r[kn].split("(")[0][-44:]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print("%-46s n=%5d total=%.3f ms" % (k, n, t / 1e6))
