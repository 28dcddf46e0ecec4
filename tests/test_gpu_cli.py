"""`python -m phylo_hmrf_b200.cli --reload 1 ...`: the reference's run() (phylo_hmrf.py:1570-1749)
on cached inputs, end to end on the GPU path, writing the reference's `.mat` result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cli_reload_run_writes_mat(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import scipy.io
    from phylo_hmrf_b200 import cli, synth, utility
    (tmp_path / "edge.1.txt").write_text("0\t1\n1\t2\n1\t3\n3\t4\n4\t5\n4\t6\n3\t7\n")
    (tmp_path / "branch_length.1.txt").write_text("\t".join(["1.0"] * 7) + "\n")
    d, B = 4, 30
    regs = [synth.make_band(21 + r, B, d) for r in range(2)]
    samples = np.concatenate([g["X_own"] for g in regs])
    len_vec, els, s = [], np.empty(2, dtype=object), 0
    for r, g in enumerate(regs):
        n = g["n_own"]
        len_vec.append([n, s, s + n, B, B, 0, 0, r, 1, 21 + r])
        els[r] = utility.edge_weightlist_grid3_undirected_unsym(g["X_own"], g["x"] * B + g["y"], B, '', 8)
        s += n
    np.save(str(tmp_path / "data.50Kb.observed.1.npy"), samples)
    np.save(str(tmp_path / "edgelist.50Kb.observed.1.npy"), els, allow_pickle=True)
    np.savetxt(str(tmp_path / "lenvec.50Kb.observed.1.txt"), np.asarray(len_vec), fmt='%d', delimiter='\t')
    opts = cli.parse_args(["-n", "3", "-r", "1", "--reload", "1", "-p", str(tmp_path), "--output", str(tmp_path),
                           "--miter", "7", "-g", "3", "--beta1", "0.1"])
    mdict = cli.run(opts)
    out = scipy.io.loadmat(str(tmp_path / "estimate_ou_1_1.00_3.mat"))
    assert out["state_vec"].size == len(samples)
    assert out["params_vec1"].shape == (3, 23) and out["cost_vec"].shape[1] == 4
    assert np.array_equal(out["len_vec"], np.asarray(len_vec))
    assert np.isfinite(mdict["cost_vec"]).all()


def test_cli_under_torchrun_on_two_gpus(tmp_path):
    """The same command line under `torchrun` with two ranks (NCCL): each rank takes LOCAL_RANK as its device and
    keeps resident only its own region, initialisation and M-step run on rank 0 and are broadcast, rank 0 alone
    writes the `.mat` file.  (The OU M-step draws random starts, so the numbers are not compared with a
    single-process run; tests/test_gpu_em_bands.py does that with a deterministic M-step.)"""
    import os
    import socket
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import scipy.io
    from phylo_hmrf_b200 import synth, utility
    (tmp_path / "edge.1.txt").write_text("0\t1\n1\t2\n1\t3\n3\t4\n4\t5\n4\t6\n3\t7\n")
    d, B = 4, 30
    regs = [synth.make_band(21 + r, B, d) for r in range(2)]
    samples = np.concatenate([g["X_own"] for g in regs])
    len_vec, els, s = [], np.empty(2, dtype=object), 0
    for r, g in enumerate(regs):
        n = g["n_own"]
        len_vec.append([n, s, s + n, B, B, 0, 0, r, 1, 21 + r])
        els[r] = utility.edge_weightlist_grid3_undirected_unsym(g["X_own"], g["x"] * B + g["y"], B, '', 8)
        s += n
    np.save(str(tmp_path / "data.50Kb.observed.1.npy"), samples)
    np.save(str(tmp_path / "edgelist.50Kb.observed.1.npy"), els, allow_pickle=True)
    np.savetxt(str(tmp_path / "lenvec.50Kb.observed.1.txt"), np.asarray(len_vec), fmt='%d', delimiter='\t')
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "phylo_hmrf_b200.cli",
                    "-n", "3", "-r", "1", "--reload", "1", "-p", str(tmp_path), "--output", str(tmp_path),
                    "--miter", "7", "-g", "3", "--beta1", "0.1"], check=True, env=env, cwd=str(tmp_path), timeout=900)
    out = scipy.io.loadmat(str(tmp_path / "estimate_ou_1_1.00_3.mat"))
    assert out["state_vec"].size == len(samples) and out["params_vec1"].shape == (3, 23)
    assert np.isfinite(out["cost_vec"]).all() and out["cost_vec"].shape[1] == 4
