set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 --no-cpu --workload mid_d9_k30 > gpurun_out/r2a_mid_bulk.json 2> gpurun_out/r2a_mid_bulk.err; tail -c 1500 gpurun_out/r2a_mid_bulk.json
PHMRF_ESTEP_KERNEL=r1 python bench.py --steps 10 --warmup 3 --no-cpu --workload mid_d9_k30 > gpurun_out/r2a_mid_r1.json 2> gpurun_out/r2a_mid_r1.err; tail -c 1500 gpurun_out/r2a_mid_r1.json
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2a_band_bulk.json 2> gpurun_out/r2a_band_bulk.err; tail -c 1800 gpurun_out/r2a_band_bulk.json
