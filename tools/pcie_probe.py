#!/usr/bin/env python
"""Host<->device copy rates of this box from pinned memory (GB/s): H2D, D2H, both at once.
The floor of bench.py's end-to-end leg is bytes / these rates."""
import json
import torch
n = 1 << 30
h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return reps * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


def both():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


out = {"h2d_gbs": timed(lambda: d1.copy_(h1, non_blocking=True)),
       "d2h_gbs": timed(lambda: h2.copy_(d2, non_blocking=True)),
       "bidir_each_gbs": timed(both)}
print(json.dumps(out))
