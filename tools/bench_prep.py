"""Time the per-region preprocessing through the public call, host buffers in and out,
next to the NumPy oracle on a smaller window.  Run on the GPU box:  python tools/bench_prep.py [W] [d]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import prep_oracle as po  # noqa: E402
from phylo_hmrf_b200 import utility  # noqa: E402


def region(seed, W, d, fill):
    rng = np.random.default_rng(seed)
    ii, jj = np.triu_indices(W)
    keep = rng.random(len(ii)) < fill
    ii, jj = ii[keep], jj[keep]
    val = np.log1p((40.0 / (1.0 + (jj - ii)))[:, None] * rng.gamma(2.0, 0.5, size=(len(ii), d)))
    return val, np.stack([ii, jj], axis=1).astype(np.int64)


def main():
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    val, pos = region(0, W, d, 0.6)
    utility._region_image(val[:1000], pos[:1000], 1, 0, 5, 50, False, 0)  # warm-up (context, module load)
    t0 = time.perf_counter()
    data1, _, _, _ = utility._region_image(val, pos, 1, 0, 5, 50, False, 0)
    t_gpu = time.perf_counter() - t0
    Wc = 96
    vc, pc = region(1, Wc, d, 0.6)
    t0 = time.perf_counter()
    po.image_pipeline_diag(vc, pc, filter_mode=0, niter=5, kappa=50, gamma=0.1)
    t_cpu = time.perf_counter() - t0
    out = {"stage": "per-region preprocessing (image, hole fill, 5 diffusion steps, flatten)", "W": W, "d": d,
           "pixels_per_species": W * W, "nodes": int(data1.shape[0]),
           "gpu_call_s": t_gpu, "gpu_pixels_per_s": W * W * d / t_gpu,
           "cpu_oracle": {"W": Wc, "s": t_cpu, "pixels_per_s": Wc * Wc * d / t_cpu, "cores": 1,
                          "kind": "port (NumPy loops, utility.py:603-630 is a Python double loop in the reference too)"}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
