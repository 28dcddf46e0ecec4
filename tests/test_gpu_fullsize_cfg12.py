"""BASELINE configurations 1 and 2 at their FULL sizes (the 652-bin chr21 and 683-bin chr22 synteny regions
of the shipped example: 212 878 and 233 586 nodes; d=4 as shipped; K=10 for config 1, K=20 for the two regions
jointly in config 2) -- sizes the vectorised oracle finishes in seconds, so every array is compared in full:
log-likelihood, integer unary / edge weights / V (bit-identical away from listed boundaries), labels from the
real GCO swap, posteriors, the four cost scalars and the statistics; through `phyloHMRF` (the product path),
region by region and summed like base.py:384-396."""
import queue

import numpy as np
import pytest

from oracle import phmrf_oracle as orc

pytestmark = pytest.mark.gpu

D, ET, BETA, BETA1 = 4, 3, 1.0, 0.1


def _regions(bins, seed):
    from phylo_hmrf_b200 import synth
    Xs, els, len_vec, s = [], [], [], 0
    for r, B in enumerate(bins):
        g = synth.make_band(seed + r, B, D, beta1=BETA1)
        n = g["n_own"]
        Xs.append(g["X_own"])
        els.append(np.column_stack([g["edge_ids"].astype(np.float64), g["edge_dist"]]))
        len_vec.append([n, s, s + n, B, B, 0, 0, r, 1, 21 + r])
        s += n
    return np.concatenate(Xs), els, len_vec


@pytest.mark.parametrize("bins,K", [([652], 10), ([652, 683], 20)])
def test_configs_1_and_2_in_full(bins, K):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from phylo_hmrf_b200 import phyloHMRF, synth
    X, els, len_vec = _regions(bins, 20261021)
    assert [lv[0] for lv in len_vec] == [b * (b + 1) // 2 for b in bins]
    means, covars = synth.model(20261021, X[:100000], K, D)
    m = phyloHMRF(n_samples=len(X), n_features=D, observation=X, edge_list_1=els, len_vec=len_vec, n_components=K,
                  estimate_type=ET, beta=BETA, beta1=BETA1)
    try:
        m.means_, m._covars_ = means, covars
        m.labels_local = np.zeros(len(X), dtype=np.int64)
        V = m.edge_potential
        total = m._initialize_sufficient_statistics()
        ref_total = orc.initialize_sufficient_statistics(K, D)
        for r, lv in enumerate(len_vec):
            s1, s2 = lv[1], lv[2]
            Xr, ids, w = X[s1:s2], m.edge_idList_undirected_vec[r], m.edge_weightList_undirected_vec[r]
            # log-likelihood, every entry
            lp = m._compute_log_likelihood(Xr)
            lp_ref = orc.compute_log_likelihood(Xr, means, covars)
            # 1e-9 relative; a log-likelihood that happens to cross zero (|logp| < 1e-3 for a handful of the
            # 4.7e6 entries) cannot be held to a relative tolerance by ANY FP64 evaluation: absolute 1e-12 there
            # (the posteriors depend on exp(logp), i.e. on the absolute error)
            np.testing.assert_allclose(lp, lp_ref, rtol=1e-9, atol=1e-12)
            assert np.max(np.abs(lp - lp_ref)[np.abs(lp_ref) > 1e-3] / np.abs(lp_ref)[np.abs(lp_ref) > 1e-3]) <= 1e-9
            # one E-step of the product path (emission, integer costs, real GCO swap, posteriors, statistics)
            q = queue.Queue()
            m._predict_posteriors(X, len_vec, r, q)
            rid, stats, labels, c_raw, c_pair, c_unary, c_total = q.get()
            qd = m.last_quantise
            u_ref, w_ref, V_ref, dwf_ref = orc.pygco_quantise(-lp_ref, w, V)
            assert abs(qd["dwf"] - dwf_ref) <= 1e-12 * dwf_ref
            assert np.array_equal(qd["w_i32"], w_ref) and np.array_equal(qd["V_i32"], V_ref)
            u_same, _, _, _ = orc.pygco_quantise(-lp, w, V, down_weight_factor=qd["dwf"])
            assert np.array_equal(qd["unary_i32"], u_same)            # bit-exact on identical inputs
            diff = np.argwhere(qd["unary_i32"] != u_ref)
            assert len(diff) <= 10 and all(abs(int(qd["unary_i32"][i, k]) - int(u_ref[i, k])) == 1 for i, k in diff)
            assert labels.shape == (lv[0],) and labels.min() >= 0 and labels.max() < K
            ref = orc.compute_posteriors_graph(V, np.asarray(labels, dtype=np.int64), lp_ref, w, ids, None, ET,
                                               faithful=False, stable=True)
            np.testing.assert_allclose([c_raw, c_pair, c_unary, c_total], ref[1:], rtol=1e-9)
            ref_stats = orc.sufficient_statistics(ref[0], Xr)
            for key in ref_stats:
                np.testing.assert_allclose(stats[key], ref_stats[key], rtol=1e-9,
                                           atol=1e-12 * np.abs(ref_stats[key]).max())
            post = m._compute_posteriors_graph(Xr, labels, lp, r)[0]
            np.testing.assert_allclose(post, ref[0], rtol=1e-9, atol=1e-300)
            total = m._accumulate_sufficient_statistics_1(total, stats)
            ref_total = orc.accumulate_sufficient_statistics_1(ref_total, ref_stats)
        np.testing.assert_allclose(total['post'].sum(), len(X), rtol=1e-11)
        for key in ('post', 'obs', 'obs*obs.T'):
            np.testing.assert_allclose(total[key], ref_total[key], rtol=1e-9, atol=1e-12 * np.abs(ref_total[key]).max())
    finally:
        m.close()
