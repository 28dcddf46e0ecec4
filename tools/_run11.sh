python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_n2.json 2> gpurun_out/${TAG}_n2.err
tail -2 gpurun_out/${TAG}_n2.err
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_n2.json')); print('n2', d['roofline']['phase_ms'], d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], d['config']['cpu_affinity'])"
nvidia-smi topo -m | head -6
