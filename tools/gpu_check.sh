#!/bin/bash
# usage (under gpurun): TAG=x tools/gpu_check.sh [workload ...]   -- GPU test suite, then one bench line per workload
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in "${@:-cfg5_chr1_10kb_band}"; do
  python bench.py --steps 10 --warmup 3 --no-cpu --workload $wl --e2e-steps 2 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err
  tail -2 gpurun_out/${TAG}_$wl.err
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_$wl.json').read().splitlines()[-1]); print('$wl', {k: round(v, 3) for k, v in d['roofline']['phase_ms'].items() if k != 'note'}, 'step', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['step']['frac_of_slower_roof'], 3), 'e2e', round(d['e2e']['ms_per_step'], 1))"
done
