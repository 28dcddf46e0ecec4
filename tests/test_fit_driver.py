"""the fork-free EM driver against the reference's own `fit_accumulate_test`
(base.py:301-455), which tests/golden/make_golden.py executed on the same scripted stand-ins
(tests/golden/fit_script.py).  CPU-only: the per-region work is scripted, so this pins the
iteration / cost-aggregation / convergence / best-iteration logic and nothing else."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import fit_script as fs  # noqa: E402

from phylo_hmrf_b200.hmrf import phyloHMRF  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "fit_driver.npz"))


class Driver(fs.ScriptedModel, phyloHMRF):
    """Scripted per-region work + the re-hosted driver (no device needed: __init__ bypassed)."""


@pytest.mark.parametrize("name", sorted(fs.SCENARIOS))
@pytest.mark.parametrize("threads", [1, 2])
def test_driver_matches_reference(name, threads):
    m_iter, thr, _ = fs.SCENARIOS[name]
    m = object.__new__(Driver)
    m.script(name)
    res = m.fit_accumulate_test(np.zeros((fs.N, fs.D)), fs.LEN_VEC, thr, "test", m_iter, n_threads=threads)
    params_vec, params_vec1, plist, it1, it2, cost_vec, t_labels = res
    assert m.iteration == int(GOLD[name + "_n_iter"])
    assert [it1, it2] == GOLD[name + "_it"].tolist()
    np.testing.assert_array_equal(cost_vec, GOLD[name + "_cost_vec"])
    np.testing.assert_array_equal(params_vec, GOLD[name + "_params_vec"])
    np.testing.assert_array_equal(params_vec1, GOLD[name + "_params_vec1"])
    np.testing.assert_array_equal(plist, GOLD[name + "_plist"])
    np.testing.assert_array_equal(t_labels, GOLD[name + "_t_labels"])
    np.testing.assert_array_equal(m.labels_local, GOLD[name + "_labels_local"])
    np.testing.assert_array_equal(m.params_vec1, GOLD[name + "_final_params_vec1"])
    np.testing.assert_array_equal(m.finalized_with, GOLD[name + "_finalized_with"])


def test_scenarios_cover_every_exit():
    its = {n: int(GOLD[n + "_n_iter"]) for n in fs.SCENARIOS}
    assert 6 < its["converges"] < 40                 # relative-change stop (base.py:421-422)
    assert its["runs_out"] == 6                       # m_iter exhausted
    assert its["stalls_after_best"] < 80              # >50 iterations since the best cost (:427-428)


def test_missing_hooks_fail_loudly():
    m = object.__new__(phyloHMRF)
    with pytest.raises(NotImplementedError):
        m._init(np.zeros((2, 2)))
    with pytest.raises(NotImplementedError):
        m._do_mstep({})
