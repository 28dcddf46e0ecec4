#!/usr/bin/env python
"""Host<->device copy rates of this box from pinned memory (GB/s): H2D, D2H, both at once.
The floor of bench.py's end-to-end leg is bytes / these rates.

    python tools/pcie_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \\
           tools/pcie_probe.py                                   # all GPUs copying at the same time: the box's total
"""
import json
import os

import torch

world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()          # every GPU copies during the same interval
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    gbs = reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    if world > 1:
        t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())  # sum over the GPUs
    return gbs


def both():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


out = {"gpus_copying_at_once": world,
       "h2d_gbs": timed(lambda: d1.copy_(h1, non_blocking=True)),
       "d2h_gbs": timed(lambda: h2.copy_(d2, non_blocking=True)),
       "bidir_each_gbs": timed(both)}
if world > 1:
    out["note"] = "sums over the GPUs; bidir_each_gbs is the sum of ONE direction while both are busy"
    if dist.get_rank() == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
else:
    print(json.dumps(out))
