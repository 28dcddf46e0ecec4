"""OU tree algebra and M-step objective against fixtures produced by the
reference's own methods (tests/golden/make_golden.py::make_ou_cases, phylo_hmrf.py:715-1325)."""
import os

import numpy as np
import pytest

from phylo_hmrf_b200 import ou

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ou_cases.npz"))
TREES = ["example", "caterpillar5"]


def _stats(p):
    return {'post': G[p + "post"], 'obs': G[p + "obs"], 'obs*obs.T': G[p + "obsobsT"]}


@pytest.mark.parametrize("name", TREES)
def test_tree_structure_matches_reference(name):
    p = name + "_"
    t = ou.OUTree(G[p + "edges"])
    assert np.array_equal(t.leaf_vec, G[p + "leaf_vec"])
    assert np.array_equal(t.A1, G[p + "A1"]) and np.array_equal(t.A2, G[p + "A2"])
    assert np.array_equal(np.asarray(t.pair_list), G[p + "pair_list"])
    assert np.array_equal(t.parent, G[p + "parent"])
    assert np.array_equal(np.concatenate(t.path_vec), G[p + "path_flat"])
    assert [len(x) for x in t.path_vec] == G[p + "path_len"].tolist()
    assert [t.leaf_list[int(l)] for l in t.leaf_vec] == G[p + "leaf_rank"].tolist()
    assert t.n_params == G[p + "params"].shape[1]
    if name == "example":  # the shipped tree: 8 nodes, leaves 2,5,6,7, 23 parameters per state
        assert t.node_num == 8 and t.leaf_vec.tolist() == [2, 5, 6, 7] and t.n_params == 23


@pytest.mark.parametrize("name", TREES)
def test_moments_match_ou_param_varied_constraint(name):
    p = name + "_"
    t = ou.OUTree(G[p + "edges"])
    _, mu, cov = t.moments(G[p + "params"])
    np.testing.assert_array_equal(mu, G[p + "means"])
    np.testing.assert_allclose(cov + 1e-3 * np.eye(t.n_leaves), G[p + "covars"], rtol=1e-15, atol=0)
    # one vector at a time gives the same numbers as the batch
    _, mu0, cov0 = t.moments(G[p + "params"][2])
    np.testing.assert_array_equal(mu0, mu[2])
    np.testing.assert_array_equal(cov0, cov[2])


@pytest.mark.parametrize("name", TREES)
def test_mstep_objective_matches_reference(name):
    p = name + "_"
    t = ou.OUTree(G[p + "edges"])
    stats, params = _stats(p), G[p + "params"]
    for c in range(len(params)):
        lik, values, V = ou.mstep_objective(t, params[c], c, stats, 5000, 1.0, 1e-3, G[p + "init_ou_params"][c])
        np.testing.assert_allclose(lik, G[p + "liks"][c], rtol=1e-12)
        np.testing.assert_array_equal(values, G[p + "values"][c])
        np.testing.assert_allclose(V, G[p + "cv_mtx"][c], rtol=1e-15, atol=0)
    batch = ou.mstep_objective_batch(t, params, stats, 5000, 1.0, 1e-3)
    np.testing.assert_allclose(batch, G[p + "liks"], rtol=1e-10)
    # out-of-box parameters with a NaN fall back to the initial estimate (:1043-1048)
    lik_nan, _, _ = ou.mstep_objective(t, G[p + "nanp"], 0, stats, 5000, 1.0, 1e-3, G[p + "init_ou_params"][0])
    np.testing.assert_allclose(lik_nan, G[p + "lik_nan"], rtol=1e-12)
    assert [t.check_params(params[0]), t.check_params(G[p + "bad"]), t.check_params(G[p + "nanp"])] == \
        G[p + "check"].tolist() == [1, -1, -2]


@pytest.mark.parametrize("name", TREES)
def test_single_objective_and_init_guess_match_reference(name):
    p = name + "_"
    t = ou.OUTree(G[p + "edges"])
    obs = G[p + "single_obs"]
    got = [ou.single_objective(t, G[p + "params"][c], obs, 1e-3) for c in (0, 2, 4)]
    np.testing.assert_allclose([x[0] for x in got], G[p + "single"], rtol=1e-12)
    np.testing.assert_allclose(got[-1][2], G[p + "single_cv"], rtol=1e-15, atol=0)

    class Legacy(object):  # the reference draws from numpy's global legacy generator
        def __init__(self, seed):
            self.rs = np.random.RandomState(seed)

        def random(self, n):
            return self.rs.rand(n)

    guess = ou.init_guess(t, G[p + "guess_mean"], 0.7, Legacy(11))
    np.testing.assert_array_equal(guess, G[p + "guess"])


def test_mstep_improves_objective_and_updates_model():
    """optimise_state (SLSQP, phylo_hmrf.py:1327-1403): the fitted parameters score at least as
    well as the initial estimate and respect the box."""
    p = "example_"
    t = ou.OUTree(G[p + "edges"])
    stats = _stats(p)
    rng = np.random.default_rng(0)
    init = G[p + "init_ou_params"]
    for c in (0, 5):
        params, lik, values, V = ou.optimise_state(t, c, stats, 5000, 1.0, 1e-3, init[c], init[c], 0.3, 0.1, 1, 0, rng)
        assert t.check_params(params) == 1
        assert lik <= ou.mstep_objective(t, init[c], c, stats, 5000, 1.0, 1e-3)[0] + 1e-9
        assert np.all(np.linalg.eigvalsh(V) > 0)
