#!/bin/bash
# usage (on the GPU box): TAG=x WLS="mid_d9_k30 cfg3_chr1_50kb" tools/run_variants.sh v0 v1 ...
# A/B timing of experiment builds (tools/build_variant.sh): each variant's library is copied over
# lib/libphmrf.so and the bench is run; the shipped library is restored at the end.
cp phylo_hmrf_b200/lib/libphmrf.so /tmp/libphmrf_shipped.so
for v in "$@"; do
  cp phylo_hmrf_b200/lib_var/$v/libphmrf.so phylo_hmrf_b200/lib/libphmrf.so
  for wl in $WLS; do
    python bench.py --steps 10 --warmup 3 --no-cpu --workload $wl --e2e-steps 1 > gpurun_out/${TAG}_${v}_$wl.json 2> gpurun_out/${TAG}_${v}_$wl.err
    python -c "
import json; d=json.load(open('gpurun_out/${TAG}_${v}_$wl.json')); print('$v $wl B=%.3f step=%.3f' % (d['roofline']['phase_ms']['B_estep'], d['ms_per_step']))"
  done
done
cp /tmp/libphmrf_shipped.so phylo_hmrf_b200/lib/libphmrf.so
