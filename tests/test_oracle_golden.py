"""Pin the CPU oracle against fixtures produced by the reference's own code
(tests/golden/make_golden.py) and against scipy for the third-party density."""
import numpy as np
import pytest
from scipy.stats import multivariate_normal

from oracle import phmrf_oracle as orc

RTOL = 1e-12


def _inc(g, p):
    flat, off = g[p + "ref_inc_flat"], g[p + "ref_inc_off"]
    return [list(flat[off[i]:off[i + 1]]) for i in range(len(off) - 1)]


def test_pairwise_potential_matches_reference(golden):
    V = orc.pairwise_potential(int(golden["K"]), float(golden["beta"]))
    assert np.array_equal(V, golden["ref_V"])


def test_edge_weights_and_incidence_match_reference(golden):
    for r in range(int(golden["n_regions"])):
        p = "r%d_" % r
        N = int(golden["len_vec"][r][0])
        w, ids, inc = orc.edge_weight_undirected(golden[p + "edge_list"], N, float(golden["beta1"]))
        assert np.array_equal(w, golden[p + "ref_edge_w"])
        assert np.array_equal(ids, golden[p + "ref_edge_ids"])
        assert inc == _inc(golden, p)


def test_gco_call_arguments(golden):
    """phylo_hmrf.py:490-498: unary = -logprob, V, weights, edges, swap, 5000 cycles, dwf None."""
    for r in range(int(golden["n_regions"])):
        p = "r%d_" % r
        s1, s2 = golden["len_vec"][r][1:3]
        lp = orc.compute_log_likelihood(golden["X"][s1:s2], golden["means"], golden["covars"])
        assert np.array_equal(-lp, golden[p + "ref_gco_unary"])
        assert np.array_equal(golden[p + "ref_gco_V"], golden["ref_V"])
        assert np.array_equal(golden[p + "ref_gco_w"], golden[p + "ref_edge_w"])
        assert np.array_equal(golden[p + "ref_gco_edges"], golden[p + "ref_edge_ids"])
        assert np.array_equal(golden[p + "ref_gco_init"], golden["init_labels"][s1:s2])
        assert int(golden[p + "ref_gco_n_iter"]) == 5000
        assert str(golden[p + "ref_gco_algorithm"]) == "swap"
        assert bool(golden[p + "ref_gco_dwf_is_none"])


@pytest.mark.parametrize("faithful", [True, False])
def test_pairwise_compare_matches_reference(golden, faithful):
    V = golden["ref_V"]
    et = int(golden["estimate_type"])
    for r in range(int(golden["n_regions"])):
        p = "r%d_" % r
        lab, w, ids = golden[p + "labels"], golden[p + "ref_edge_w"], golden[p + "ref_edge_ids"]
        if faithful:
            pp = orc.pairwise_compare_faithful(V, lab, _inc(golden, p), w, ids, et)
            assert np.array_equal(pp, golden[p + "ref_pp"])
        else:
            pp = orc.pairwise_compare_vec(V, lab, w, ids, et)
            np.testing.assert_allclose(pp, golden[p + "ref_pp"], rtol=RTOL, atol=0)


@pytest.mark.parametrize("faithful", [True, False])
def test_posteriors_costs_stats_match_reference(golden, faithful):
    V = golden["ref_V"]
    et = int(golden["estimate_type"])
    tot = orc.initialize_sufficient_statistics(int(golden["K"]), int(golden["d"]))
    for r in range(int(golden["n_regions"])):
        p = "r%d_" % r
        s1, s2 = golden["len_vec"][r][1:3]
        X = golden["X"][s1:s2]
        out = orc.compute_posteriors_graph(V, golden[p + "labels"], golden[p + "logprob"], golden[p + "ref_edge_w"],
                                           golden[p + "ref_edge_ids"], _inc(golden, p), et, faithful=faithful)
        post, costs = out[0], np.asarray(out[1:])
        if faithful:
            assert np.array_equal(post, golden[p + "ref_post"])
            assert np.array_equal(costs, golden[p + "ref_costs"])
        else:
            np.testing.assert_allclose(post, golden[p + "ref_post"], rtol=RTOL, atol=1e-300)
            np.testing.assert_allclose(costs, golden[p + "ref_costs"], rtol=RTOL)
        st = orc.sufficient_statistics(golden[p + "ref_post"], X)
        assert np.array_equal(st["post"], golden[p + "ref_stats_post"])
        assert np.array_equal(st["obs"], golden[p + "ref_stats_obs"])
        assert np.array_equal(st["obs*obs.T"], golden[p + "ref_stats_obsobsT"])
        tot = orc.accumulate_sufficient_statistics_1(tot, st)
    assert np.array_equal(tot["post"], golden["ref_total_post"])
    assert np.array_equal(tot["obs"], golden["ref_total_obs"])
    assert np.array_equal(tot["obs*obs.T"], golden["ref_total_obsobsT"])


def test_stable_softmax_equals_naive_where_finite(golden):
    V = golden["ref_V"]
    et = int(golden["estimate_type"])
    for r in range(int(golden["n_regions"])):
        p = "r%d_" % r
        args = (V, golden[p + "labels"], golden[p + "logprob"], golden[p + "ref_edge_w"], golden[p + "ref_edge_ids"],
                None, et)
        a = orc.compute_posteriors_graph(*args, faithful=False, stable=False)
        b = orc.compute_posteriors_graph(*args, faithful=False, stable=True)
        finite = np.isfinite(a[0]).all(axis=1)
        assert finite.any()
        np.testing.assert_allclose(b[0][finite], a[0][finite], rtol=1e-12, atol=1e-300)


def test_density_matches_scipy(golden):
    """Third-party anchor: sklearn-0.18 'full' density == scipy's multivariate normal logpdf."""
    X, means, covars = golden["X"], golden["means"], golden["covars"]
    lp = orc.log_multivariate_normal_density_full(X, means, covars)
    for k in range(len(means)):
        ref = multivariate_normal(mean=means[k], cov=covars[k]).logpdf(X)
        np.testing.assert_allclose(lp[:, k], ref, rtol=1e-11, atol=1e-11)


def test_density_cholesky_fallback():
    """sklearn 0.18 retries with +1e-7*I when the Cholesky fails."""
    X = np.array([[0.1, 0.2], [0.3, 0.1]])
    cv = np.array([[[1.0, 1.0], [1.0, 1.0]]])  # singular
    lp = orc.log_multivariate_normal_density_full(X, np.zeros((1, 2)), cv)
    ref = multivariate_normal(mean=np.zeros(2), cov=cv[0] + 1e-7 * np.eye(2)).logpdf(X)
    np.testing.assert_allclose(lp[:, 0], ref, rtol=1e-6)
    with pytest.raises(ValueError):
        orc.log_multivariate_normal_density_full(X, np.zeros((1, 2)), np.array([[[1.0, 2.0], [2.0, 1.0]]]))


def test_quantiser_contract():
    """pygco contract (third-party, parity unpinned): divide, multiply, truncate toward zero.
    Scale factors 1e5 (unary), 1e3 (edge weights), 1e2 (V): pygco's "pairwise * smooth = unary"."""
    unary = np.array([[0.0, 3.0, -1.5], [2.999999, 1.0, 0.5]])
    w = np.array([0.9, 0.25])
    V = orc.pairwise_potential(3, 2.0)
    u_i, w_i, V_i, dwf = orc.pygco_quantise(unary, w, V)
    assert dwf == 3.0 + 1e-10
    assert u_i.dtype == np.intc and u_i.flags.c_contiguous
    assert u_i[0, 1] == 99999 and u_i[0, 0] == 0
    assert u_i[0, 2] == -49999  # truncation toward zero, not floor
    assert np.array_equal(V_i, np.array([[0, 200, 200], [200, 0, 200], [200, 200, 0]]))
    assert orc.PYGCO_PAIRWISE_FLOAT_PRECISION * orc.PYGCO_SMOOTH_COST_PRECISION == orc.PYGCO_UNARY_FLOAT_PRECISION
    # the scale factors are parameters (a maintainer can match another pygco build)
    _, w_k, V_k, _ = orc.pygco_quantise(unary, w, V, pairwise_precision=100, smooth_precision=1000)
    assert np.array_equal(V_k, 10 * V_i) and np.array_equal(w_k, np.array([int(0.9 / dwf * 100), int(0.25 / dwf * 100)]))
    assert np.array_equal(w_i, np.array([int(0.9 / dwf * 1000), int(0.25 / dwf * 1000)]))
    # weights dominate the down-weight factor when the unary is small
    _, _, _, dwf2 = orc.pygco_quantise(unary * 1e-3, w, V)
    assert dwf2 == 0.9 * 2.0 + 1e-10


def test_triangle_edges_match_reference_builder(golden):
    """oracle.triangle_edges/edge_distances reproduce utility.py's diagonal-region edge list."""
    lv = golden["len_vec"][0]
    if int(lv[8]) != 1:
        pytest.skip("first region is not diagonal")
    B = int(lv[3])
    el = golden["r0_edge_list"]
    e = orc.triangle_edges(B)
    if len(e) != len(el):  # the isolate cases removed edges of one node
        keep = np.ones(len(e), bool)
        have = set(map(tuple, np.int64(el[:, :2])))
        keep = np.array([tuple(x) in have for x in e])
        e = e[keep]
    assert np.array_equal(e, np.int64(el[:, :2]))
    X = golden["X"][lv[1]:lv[2]]
    np.testing.assert_allclose(orc.edge_distances(X, e, B), el[:, 2], rtol=1e-12, atol=0)
