#!/usr/bin/env python
"""Opcode counts of the hot kernels in the shipped library (cuobjdump -sass): which machine instructions the
emission, quantise and E-step kernels are made of.  usage: tools/sass_counts.py [path/to/libphmrf.so]"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "phylo_hmrf_b200", "lib", "libphmrf.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = [("emit_kernel<9>", r"emit_kernelILi9ELb1ELi4E"), ("quantise_kernel", r"quantise_kernel"),
        ("estep_bulk_kernel<9,4,12,grid,6>", r"estep_bulk_kernelILi9ELi4ELi12ELb1ELi6E"),
        ("estep_bulk_kernel<5,3,12,grid,4>", r"estep_bulk_kernelILi5ELi3ELi12ELb1ELi4E"),
        ("estep_bulk_kernel<4,2,12,explicit,2>", r"estep_bulk_kernelILi4ELi2ELi12ELb0ELi2E")]
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "UBLKPF", "SYNCS", "USETMAXREG", "LDS", "STS", "LDG", "STG", "MUFU",
        "BAR", "WARPSYNC", "STL", "LDL"]
cur, counts, arch = None, {}, None
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = None
        for name, pat in WANT:
            if re.search(pat, m.group(1)):
                cur = name
                counts.setdefault(cur, collections.Counter())
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            counts[cur][m.group(1)] += 1
print("library:", os.path.basename(lib), " arch:", arch)
for name, _ in WANT:
    c = counts.get(name)
    if not c:
        print("%-40s (not found)" % name)
        continue
    print("%-40s total=%d  " % (name, sum(c.values())) + "  ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k]))
