"""BASELINE configs 1-2 on the reference's SHIPPED example data: a 240-bin window of chr22 (three species
as shipped + a stand-in for the missing hg38 track), the example tree's 4 leaves, K = 10 and K = 20.
tests/golden/example_chr22.npz holds what the reference's own `_predict_posteriors` /
`_compute_posteriors_graph` produced on it (tests/golden/make_golden_example.py).  CPU tier: the oracle
against that fixture; GPU tier: the library through `phyloHMRF`."""
import os

import numpy as np
import pytest

from oracle import phmrf_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "example_chr22.npz"))
X = G["X"]
W, STRIDE = int(G["window"]), int(G["stride"])
N = len(X)
LEN_VEC = [[N, 0, N, W, W, 0, 0, 0, 1, 22]]


def _edges_cpu():
    e = orc.triangle_edges(W)
    return e, orc.edge_distances(X, e, W)


def test_fixture_is_the_example_window():
    assert X.shape == (W * (W + 1) // 2, 4) and N == 28920
    assert 0.2 < np.mean(X == 0) < 0.9 and X.max() < 12          # sparse log-scaled contact counts
    e, dist = _edges_cpu()
    assert len(e) == int(G["n_edges"])
    np.testing.assert_allclose(dist.sum(), float(G["edge_dist_sum"]), rtol=1e-12)
    np.testing.assert_array_equal(e[:200], G["edge_head"][:, :2].astype(np.int64))
    np.testing.assert_allclose(dist[:200], G["edge_head"][:, 2], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("K", [10, 20])
def test_oracle_reproduces_the_reference_on_the_example_data(K):
    p = "k%d_" % K
    means, covars, labels = G[p + "means"], G[p + "covars"], G[p + "labels"].astype(np.int64)
    e, dist = _edges_cpu()
    w = np.exp(-float(G["beta1"]) * dist)
    np.testing.assert_allclose(w.sum(), float(G[p + "ref_edge_w_sum"]), rtol=1e-12)
    V = orc.pairwise_potential(K, float(G["beta"]))
    lp = orc.compute_log_likelihood(X, means, covars)
    np.testing.assert_allclose(lp[::STRIDE], G[p + "ref_logprob_rows"], rtol=1e-12)
    assert np.abs(lp).max() == float(G[p + "ref_absmax_unary"])
    # the reference's softmax is the naive one: fine here, the window's log-likelihoods stay above -700
    ref = orc.compute_posteriors_graph(V, labels, lp, w, e, None, int(G["estimate_type"]), faithful=False, stable=True)
    np.testing.assert_allclose(ref[0][::STRIDE], G[p + "ref_post_rows"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(ref[1:], G[p + "ref_costs"], rtol=1e-10)
    st = orc.sufficient_statistics(ref[0], X)
    np.testing.assert_allclose(st["post"], G[p + "ref_stats_post"], rtol=1e-10)
    np.testing.assert_allclose(st["obs"], G[p + "ref_stats_obs"], rtol=1e-10)
    np.testing.assert_allclose(st["obs*obs.T"], G[p + "ref_stats_obsobsT"], rtol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("K", [10, 20])
def test_library_reproduces_the_reference_on_the_example_data(K):
    from phylo_hmrf_b200 import phyloHMRF, utility
    p = "k%d_" % K
    means, covars, labels = G[p + "means"], G[p + "covars"], G[p + "labels"].astype(np.int64)
    iu = np.triu_indices(W)
    el = np.asarray(utility.edge_weightlist_grid3_undirected_unsym(X, iu[0] * W + iu[1], W, '', 8))   # on the GPU
    assert len(el) == int(G["n_edges"])
    np.testing.assert_array_equal(el[:200, :2], G["edge_head"][:, :2])
    np.testing.assert_allclose(el[:200, 2], G["edge_head"][:, 2], rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(el[:, 2].sum(), float(G["edge_dist_sum"]), rtol=1e-11)
    m = phyloHMRF(n_samples=N, n_features=4, observation=X, edge_list_1=[el], len_vec=LEN_VEC, n_components=K,
                  estimate_type=int(G["estimate_type"]), beta=float(G["beta"]), beta1=float(G["beta1"]))
    try:
        m.means_, m._covars_ = means, covars
        lp = m._compute_log_likelihood(X)
        ref_rows = G[p + "ref_logprob_rows"]
        assert np.max(np.abs(lp[::STRIDE] - ref_rows) / np.abs(ref_rows)) <= 1e-9
        # integer cost arrays: bit-identical to the contract applied to the same log-likelihood
        reg = m._regions[0]
        q = reg.quantise(boundary_cap=lp.size)
        u_ref, w_ref, V_ref, dwf = orc.pygco_quantise(-lp, m.edge_weightList_undirected_vec[0], m.edge_potential)
        assert q["dwf"] == dwf and abs(dwf - (float(G[p + "ref_absmax_unary"]) + 1e-10)) <= 1e-9 * dwf
        assert np.array_equal(q["unary_i32"], u_ref) and np.array_equal(q["w_i32"], w_ref)
        assert np.array_equal(q["V_i32"], V_ref)
        out = m._compute_posteriors_graph(X, labels, lp, 0)
        np.testing.assert_allclose(out[0][::STRIDE], G[p + "ref_post_rows"], rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(out[1:], G[p + "ref_costs"], rtol=1e-9)
        st = m._last_stats
        for key, ref in (("post", "ref_stats_post"), ("obs", "ref_stats_obs"), ("obs*obs.T", "ref_stats_obsobsT")):
            np.testing.assert_allclose(st[key], G[p + ref], rtol=1e-9, atol=1e-12 * np.abs(G[p + ref]).max())
    finally:
        m.close()
