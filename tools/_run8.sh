python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for wl in cfg5_chr1_10kb_band cfg3_chr1_50kb mid_d9_k30; do
python bench.py --steps 10 --warmup 3 --no-cpu --workload $wl --e2e-steps 1 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_$wl.json')); print('$wl', d['roofline']['phase_ms'], d['ms_per_step'])"
done
