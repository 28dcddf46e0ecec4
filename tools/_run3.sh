python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py tests/test_gpu_grid_region.py tests/test_gpu_fullsize_band.py -m gpu -x -q 2>&1 | tail -5
for v in 1 0; do
PHMRF_GRID_IMPLICIT=$v python bench.py --steps 10 --warmup 3 --no-cpu --workload mid_d9_k30 --e2e-steps 1 > gpurun_out/r2d_mid_$v.json 2> gpurun_out/r2d_mid_$v.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_mid_$v.json')); print('mid implicit=$v', d['roofline']['phase_ms'])"
done
python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2d_band.json 2> gpurun_out/r2d_band.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_band.json')); print('band', d['roofline']['phase_ms'], d['ms_per_step'])"
