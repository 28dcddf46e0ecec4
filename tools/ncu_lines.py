#!/usr/bin/env python
"""Attribute ncu warp-stall samples of one kernel to CUDA source lines.
usage: tools/ncu_lines.py <rep> <kernel-regex> <cubin> <mangled-substring> [top]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info (same order)."""
import csv
import re
import subprocess
import sys

rep, kre, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
# first kernel instance only
body = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        break
    body.append(r)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines = []
cur = None
infn = False
for l in dis.splitlines():
    if l.startswith(".text."):
        infn = mangled in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
si, src = hdr.index("# Samples"), hdr.index("Source")
ie = hdr.index("Instructions Executed")
wf, ex = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive")
print("sass rows", len(body), "disasm instrs", len(lines))
agg = {}
tot = 0
for r, ln in zip(body, lines):
    s = int(r[si] or 0)
    tot += s
    a = agg.setdefault(ln, [0, 0, 0, 0, 0])
    a[0] += s
    a[1] += int(r[ie] or 0)
    a[2] += int(r[wf] or 0)
    a[3] += int(r[ex] or 0)
    a[4] += 1
srcfile = {}
sortkey = 1 if len(sys.argv) > 6 and sys.argv[6] == "exec" else 0
for (ln, a) in sorted(agg.items(), key=lambda kv: -kv[1][sortkey])[:top]:
    text = ""
    if ln:
        try:
            if ln[0] not in srcfile:
                srcfile[ln[0]] = open("/root/repo/phylo_hmrf_b200/csrc/" + ln[0]).read().splitlines()
            text = srcfile[ln[0]][ln[1] - 1].strip()[:90]
        except Exception:
            text = "?"
    print("%5.1f%% %-22s sass=%-4d exec=%-11d shwf=%-10d exc=%-10d %s" % (100.0 * a[0] / max(tot, 1), str(ln), a[4], a[1], a[2], a[3], text))
