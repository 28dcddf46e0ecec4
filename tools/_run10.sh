python tools/pcie_probe.py
nvidia-smi topo -m 2>/dev/null | head -8
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
for L in 3 1; do
python bench.py --steps 5 --warmup 3 --e2e-steps 3 --no-cpu --lanes $L > gpurun_out/${TAG}_cfg5_l$L.json 2> gpurun_out/${TAG}_cfg5_l$L.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_cfg5_l$L.json')); print('lanes $L e2e', d['e2e']['ms_per_step'], d['config']['cpu_affinity'])"
done
