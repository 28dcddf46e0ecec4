#!/bin/bash
# usage: tools/prof.sh <tag> <workload> [kernel-regex] [skip] [count]
# One ncu launch list + one --set full capture of the named kernels (1 GPU only).
TAG=$1; WL=$2; KRE=${3:-"emit_kernel|estep_kernel|quantise_unary"}; SKIP=${4:-9}; CNT=${5:-3}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $CNT -f -o gpurun_out/prof_$TAG \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu --e2e-steps 1 >> gpurun_out/bench_under_ncu_$TAG.log 2>&1
ls -la gpurun_out/ | grep $TAG
