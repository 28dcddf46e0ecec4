"""End-to-end EM loop on the GPU path: re-hosted driver (base.py:301-455) + real kernels + real
GCO swap.  The M-step and the initialisation are plain Gaussian-moment STAND-INS supplied through
the hooks (the reference's OU/SLSQP M-step is outside the hot-path scope)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_em_driver_end_to_end_two_regions():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import phylo_hmrf_b200 as ph
    from phylo_hmrf_b200 import synth, utility
    d, K, seed = 4, 5, 99
    regs = [synth.make_band(seed + r, B, d) for r, B in enumerate((40, 33))]
    X = np.concatenate([g["X_own"] for g in regs])
    len_vec, els, s = [], [], 0
    for r, g in enumerate(regs):
        n = g["n_own"]
        len_vec.append([n, s, s + n, g["B"], g["B"], 0, 0, r, 1, 21 + r])
        # edge lists from the GPU builder (next row f-1), as the reference's loader would hand them over
        serial = g["x"] * g["B"] + g["y"]
        els.append(utility.edge_weightlist_grid3_undirected_unsym(g["X_own"], serial, g["B"], '', 8))
        s += n
    model = ph.phyloHMRF(len(X), d, beta=1.0, beta1=0.1, observation=X, edge_list_1=els, len_vec=len_vec,
                         n_components=K, estimate_type=3)
    seen = []

    def init_fn(m, Xall):
        rng = np.random.default_rng(0)
        m.means_ = Xall[rng.choice(len(Xall), K, replace=False)].copy()
        cv = np.cov(Xall.T) + m.min_covar * np.eye(d)
        m._covars_ = np.stack([cv] * K)
        m.params_vec1 = np.zeros((K, 3))
        d2 = ((Xall[:, None, :] - m.means_[None]) ** 2).sum(-1)
        m.labels = np.argmin(d2, axis=1).astype(np.int64)
        m.labels_local = m.labels.copy()

    def mstep_fn(m, stats):
        post = np.maximum(stats['post'], 1e-8)
        np.testing.assert_allclose(stats['post'].sum(), len(X), rtol=1e-9)
        mu = stats['obs'] / post[:, None]
        cv = stats['obs*obs.T'] / post[:, None, None] - mu[:, :, None] * mu[:, None, :]
        m.means_ = mu
        m._covars_ = 0.5 * (cv + cv.transpose(0, 2, 1)) + m.min_covar * np.eye(d)
        m.params_vec1 = m.params_vec1 + 1.0
        seen.append(stats['post'].copy())

    model.init_fn, model.mstep_fn = init_fn, mstep_fn
    res = model.fit_accumulate_test(X, len_vec, 1e-3, "test", 12)
    params_vec, params_vec1, plist, it1, it2, cost_vec, t_labels = res
    assert cost_vec.shape[1] == 4 and 7 <= len(cost_vec) <= 12
    assert np.isfinite(cost_vec).all()
    assert cost_vec[-1, 3] < cost_vec[0, 3]          # total cost went down
    assert t_labels.shape == (len(X),) and set(np.unique(t_labels)) <= set(range(K))
    assert len(seen) >= 6 and plist.shape[0] == len(cost_vec)
    # labels_local is the labelling of the best iteration and feeds the next graph cut (quirk 9)
    assert model.labels_local.shape == (len(X),)
    model.close()


def test_em_with_ou_mstep_on_example_tree():
    """The whole per-iteration chain of the reference on the shipped example tree (4 leaves):
    K-means + OU initial fits, GPU E-step, host GCO swap, OU/SLSQP M-step (ou.py)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import phylo_hmrf_b200 as ph
    from phylo_hmrf_b200 import synth, utility
    tree_edges = np.array([[0, 1], [1, 2], [1, 3], [3, 4], [4, 5], [4, 6], [3, 7]])  # example_input/edge.1.txt
    d, K = 4, 3
    g = synth.make_band(5, 36, d)
    X = g["X_own"]
    len_vec = [[len(X), 0, len(X), 36, 36, 0, 0, 0, 1, 21]]
    els = [utility.edge_weightlist_grid3_undirected_unsym(X, g["x"] * 36 + g["y"], 36, '', 8)]
    model = ph.phyloHMRF(len(X), d, edge_list=tree_edges, branch_list=np.ones(7), cons_param=1, beta=1.0, beta1=0.1,
                         observation=X, edge_list_1=els, len_vec=len_vec, n_components=K, estimate_type=3,
                         random_state=3)
    assert model.n_params == 23
    res = model.fit_accumulate_test(X, len_vec, 1e-3, "test", 8)
    params_vec, params_vec1, plist, it1, it2, cost_vec, t_labels = res
    assert params_vec.shape == (K, 23) and np.isfinite(cost_vec).all() and len(cost_vec) >= 7
    assert model.tree.check_params(params_vec[0]) == 1
    assert model.means_.shape == (K, d) and model._covars_.shape == (K, d, d)
    for c in range(K):
        assert np.linalg.eigvalsh(model._covars_[c]).min() > 0
    assert t_labels.shape == (len(X),)
    model.close()
