#!/usr/bin/env python
"""DRAM bytes per launch of the three hot-path kernels from an `ncu --set full` capture.
usage: tools/ncu_traffic.py <rep> <workload> [profiles/r2_traffic.json]   (merges into the JSON file)"""
import csv
import json
import os
import subprocess
import sys

rep, workload = sys.argv[1:3]
out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "profiles", "r2_traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
acc = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    key = "estep" if "estep" in name else ("emit" if "emit_kernel" in name else ("quantise" if "quantise" in name else None))
    if key is None:
        continue
    tot = 0.0
    for col in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(col)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
        tot += float(r[i]) * scale
    acc.setdefault(key, []).append(tot)
table = json.load(open(out)) if os.path.exists(out) else {}
table[workload] = {k: int(sum(v) / len(v)) for k, v in acc.items()}
table[workload]["source"] = os.path.basename(rep)
json.dump(table, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(table[workload]))
