#!/usr/bin/env python
"""Top SASS instructions by stall samples with their dominant stall reasons.
usage: tools/ncu_sass.py <rep> <kernel-regex> [top]"""
import csv, subprocess, sys
rep, kre = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r == hdr:
        break
    body.append(r)
si, src = hdr.index("# Samples"), hdr.index("Source")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
agg = {}
for r in body:
    for i, h in stall_cols:
        agg[h] = agg.get(h, 0) + int(r[i] or 0)
print("total samples", tot, {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
for i in order:
    r = body[i]
    st = sorted(((int(r[c] or 0), h) for c, h in stall_cols), reverse=True)[:3]
    print("%5.2f%% #%-5d %-70s %s" % (100.0 * int(r[si] or 0) / tot, i, r[src][:70], " ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
