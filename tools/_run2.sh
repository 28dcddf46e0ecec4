python -m pytest tests/test_gpu_parity.py tests/test_gpu_paths.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --workload mid_d9_k30 --e2e-steps 1 > gpurun_out/r2b_mid.json 2> gpurun_out/r2b_mid.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_mid.json')); print('mid', d['roofline']['phase_ms'])"
python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2b_band.json 2> gpurun_out/r2b_band.err; python -c "
import json; d=json.load(open('gpurun_out/r2b_band.json')); print('band', d['roofline']['phase_ms'], d['ms_per_step'])"
