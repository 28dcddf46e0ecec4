#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG=V ...]   -> phylo_hmrf_b200/lib_var/NAME/libphmrf.so
# Experiment builds of the bulk-copy phase-B kernel: only the bench shapes are instantiated
# (PHMRF_B3_FAST_BUILD), every other object comes from the regular build.
set -e
NAME=$1; shift
PKG=phylo_hmrf_b200; OUT=$PKG/lib_var/$NAME; mkdir -p $OUT
for f in kernels_b3 kernels_b3_d58 kernels_b3_d9c; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas=-v -DPHMRF_B3_FAST_BUILD "$@" \
     -c $PKG/csrc/$f.cu -o $OUT/$f.o > $OUT/$f.log 2>&1 &
done
wait
OBJS=""
for f in api kernels_a kernels_b kernels_grid kernels_prep probe; do OBJS="$OBJS $PKG/build/$f.cu.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libphmrf.so $OBJS $OUT/kernels_b3.o $OUT/kernels_b3_d58.o $OUT/kernels_b3_d9c.o
python - $OUT <<'PY'
import re, sys, glob
for f in sorted(glob.glob(sys.argv[1] + "/*.log")):
    t = open(f).read()
    if "error" in t: print(t[:2000])
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", t):
        k = re.search(r"estep_bulk_kernelILi(\d+)ELi(\d+)ELi(\d+)ELb([01])ELi(\d+)E", m.group(1))
        if k: print("D=%s NK8=%s P=%s GRID=%s KR=%s" % k.groups(), "spill st/ld", m.group(3), m.group(4), "regs", m.group(5))
PY
