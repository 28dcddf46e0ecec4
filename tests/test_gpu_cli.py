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
