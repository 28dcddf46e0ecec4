"""Few-warp issue-rate probes (cycles per instruction per warp); run on the GPU box."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phylo_hmrf_b200 import engine  # noqa: E402

NAMES = {7: "8 DFMA warps alone (ILP 8): cycles/DFMA", 8: "4 DFMA warps alone (ILP 8): cycles/DFMA",
         14: "12 DFMA warps alone (ILP 8): cycles/DFMA", 9: "4 DMMA warps alone (28 acc): cycles/DMMA",
         10: "4 DMMA + 8 DFMA (ILP 8): cycles/DFMA", 11: "4 DMMA + 8 DFMA (ILP 8): cycles/DMMA",
         12: "4 DMMA + 8 DFMA (ILP 4): cycles/DFMA", 13: "4 DMMA + 8 DFMA (ILP 4): cycles/DMMA"}
for w, name in NAMES.items():
    print(w, name, "%.2f" % engine.probe(w))
