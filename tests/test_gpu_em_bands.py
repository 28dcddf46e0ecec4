"""Row bands in the product path on hardware: `phyloHMRF.fit_accumulate_test` under `torchrun` with two
ranks (NCCL) cuts the one region into two bands -- emission per band, max|logp| max-reduced, integer
unary sent to the owner, GCO there, label windows sent back, E-step per band, one all-reduce of the
statistics -- and must reproduce the single-GPU run of the same problem.  Needs two GPUs."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
B, D, K, M_ITER = 96, 5, 8, 8


@pytest.mark.gpu
def test_two_gpu_banded_em_matches_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    worker = os.path.join(HERE, "mgpu_em_worker.py")
    args = [str(tmp_path), str(B), str(D), str(K), str(M_ITER)]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    subprocess.run([sys.executable, worker] + args, check=True, env=env, timeout=600)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), worker] + args, check=True, env=env,
                   timeout=900)
    one = np.load(str(tmp_path / "w1_rank0.npz"))
    assert len(one["cost_vec"]) >= 6
    assert int(one["n_grid"]) == int(one["n_pieces"]) == 1    # the edge list is the grid's: built on the device
    for rank in (0, 1):
        two = np.load(str(tmp_path / ("w2_rank%d.npz" % rank)))
        assert int(two["n_bands"]) == 1                       # each rank held one band of the region ...
        assert int(two["n_grid"]) == int(two["n_pieces"]) == 1   # ... built on the device from the grid geometry
        np.testing.assert_allclose(two["cost_vec"], one["cost_vec"], rtol=1e-10)
        np.testing.assert_array_equal(two["t_labels"], one["t_labels"])
        np.testing.assert_array_equal(two["labels_local"], one["labels_local"])
        np.testing.assert_allclose(two["means"], one["means"], rtol=1e-10)
