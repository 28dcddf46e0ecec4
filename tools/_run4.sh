python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py tests/test_gpu_grid_region.py -m gpu -x -q 2>&1 | tail -5
for v in 12 8; do
PHMRF_BULK_P=$v python bench.py --steps 10 --warmup 3 --no-cpu --workload mid_d9_k30 --e2e-steps 1 > gpurun_out/r2f_mid_$v.json 2> gpurun_out/r2f_mid_$v.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_mid_$v.json')); print('mid P=$v', d['roofline']['phase_ms'])"
PHMRF_BULK_P=$v python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2f_band_$v.json 2> gpurun_out/r2f_band_$v.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_band_$v.json')); print('band P=$v', d['roofline']['phase_ms'], d['ms_per_step'])"
done
PHMRF_BULK_P=12 python bench.py --steps 10 --warmup 3 --no-cpu --workload cfg3_chr1_50kb --e2e-steps 1 > gpurun_out/r2f_cfg3.json 2> gpurun_out/r2f_cfg3.err; python -c "
import json; d=json.load(open('gpurun_out/r2f_cfg3.json')); print('cfg3', d['roofline']['phase_ms'], d['ms_per_step'])"
