"""Regions built on the device from the grid geometry (phmrf_region_create_grid)
against regions built from host edge lists: same edge list, same neighbour slots, same E-step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ph():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import phylo_hmrf_b200 as ph
    return ph


def _compare(ph, m, X_window, host_edges, host_w, kind, n1, n2, r0, r1, nn, beta1, own_offset, n_own, et, K, seed):
    rg = m.region_grid(X_window, kind, n1, n2, r0, r1, nn, beta1)
    assert (rg.n, rg.n_window, rg.own_offset) == (n_own, len(X_window), own_offset)
    ids, w = rg.edges()
    assert np.array_equal(ids, host_edges)
    np.testing.assert_allclose(w, host_w, rtol=1e-14, atol=0)
    rh = m.region(X_window[own_offset:own_offset + n_own], host_edges, host_w, n_window=len(X_window),
                  own_offset=own_offset)
    lab = np.random.default_rng(seed).integers(0, K, size=len(X_window)).astype(np.int32)
    out = []
    for r in (rg, rh):
        r.emit_loglik()
        q = r.quantise()
        r.set_labels(lab)
        st, su, _ = r.estep_stats(et)
        out.append((q, st, su))
        r.close()
    (qg, sg, ug), (qh, sh, uh) = out
    assert np.array_equal(qg["unary_i32"], qh["unary_i32"])
    assert np.abs(qg["w_i32"].astype(np.int64) - qh["w_i32"]).max() <= 1  # exp on device vs numpy: <= 1 ulp
    for k in sg:
        np.testing.assert_allclose(sg[k], sh[k], rtol=1e-12, atol=1e-12 * np.abs(sh[k]).max())
    np.testing.assert_allclose(ug, uh, rtol=1e-12)


@pytest.mark.parametrize("nn", [8, 4])
def test_triangle_whole_and_bands(ph, nn):
    from phylo_hmrf_b200 import synth
    from oracle import phmrf_oracle as orc
    B, d, K, beta1 = 57, 5, 11, 0.3
    means, covars = synth.model(3, synth.make_band(3, B, d)["X_own"], K, d)
    m = ph.Model(K, d)
    m.set_model(means, covars, synth.potts(K, 1.0))
    for r0, r1 in [(0, B)] + synth.band_rows(B, 3):
        g = synth.make_band(3, B, d, r0, r1, beta1=beta1)
        e, w = g["edge_ids"], g["edge_w"]
        if nn == 4:  # keep right / lower edges only: same row (id2 == id1+1) or same column
            x, y = g["x"], g["y"]
            keep = (x[e[:, 0]] == x[e[:, 1]]) | (y[e[:, 0]] == y[e[:, 1]])
            e, w = e[keep], w[keep]
        _compare(ph, m, g["X_window"], e, w, 1, B, B, r0, r1, nn, beta1, g["own_offset"], g["n_own"], 3, K, 7)
    m.close()


def test_rectangle_matches_reference_style_edge_list(ph):
    from phylo_hmrf_b200 import synth, utility
    n1, n2, d, K, beta1 = 23, 31, 4, 6, 0.5
    rng = np.random.default_rng(0)
    X = np.log1p(rng.gamma(2.0, 0.5, size=(n1 * n2, d)))
    el = utility.edge_weightlist_grid3_undirected(X, np.arange(n1 * n2), (n1, n2), '', 8)
    e, w = np.int64(el[:, :2]), np.exp(-beta1 * el[:, 2])
    means, covars = synth.model(1, X, K, d)
    m = ph.Model(K, d)
    m.set_model(means, covars, synth.potts(K, 0.7))
    _compare(ph, m, X, e, w, 0, n1, n2, 0, n1, 8, beta1, 0, n1 * n2, 3, K, 1)
    # a band of the rectangle: rows [5, 14) with halo rows 4 and 14
    r0, r1 = 5, 14
    w0, w1 = (r0 - 1) * n2, (r1 + 1) * n2
    o0, o1 = r0 * n2, r1 * n2
    inc = ((e[:, 0] >= o0) & (e[:, 0] < o1)) | ((e[:, 1] >= o0) & (e[:, 1] < o1))
    _compare(ph, m, X[w0:w1], e[inc] - w0, w[inc], 0, n1, n2, r0, r1, 8, beta1, o0 - w0, o1 - o0, 0, K, 2)
    with pytest.raises(ValueError):
        m.region_grid(X[:10], 0, n1, n2)
    m.close()


def test_product_path_takes_the_implicit_grid_for_grid_built_edge_lists():
    """`phyloHMRF` recognises an edge list that is exactly the grid's (geometry in len_vec, same ids, same
    weights) and builds the region on the device; integer cost arrays are those of the host's own weights and the
    E-step agrees with the explicit-slot region to rounding.  A list with an edge removed stays explicit."""
    import queue
    from phylo_hmrf_b200 import engine, phyloHMRF, synth, utility
    B, d, K = 40, 5, 6
    g = synth.make_band(91, B, d, beta1=0.1)
    X = g["X_own"]
    el = np.asarray(utility.edge_weightlist_grid3_undirected_unsym(X, g["x"] * B + g["y"], B, '', 8))
    len_vec = [[len(X), 0, len(X), B, B, 0, 0, 0, 1, 21]]
    means, covars = synth.model(91, X, K, d)
    out = {}
    for flag in (True, False):
        m = phyloHMRF(n_samples=len(X), n_features=d, observation=X, edge_list_1=[el], len_vec=len_vec, n_components=K,
                      estimate_type=3, beta=1.0, beta1=0.1, implicit_grid=flag)
        try:
            assert isinstance(m._regions[0], engine.GridRegion) == flag
            m.means_, m._covars_ = means, covars
            m.labels_local = np.zeros(len(X), dtype=np.int64)
            q = queue.Queue()
            m._predict_posteriors(X, len_vec, 0, q)
            tup = q.get()
            out[flag] = (tup, {k: v.copy() for k, v in m.last_quantise.items() if k in ("unary_i32", "w_i32", "V_i32")},
                         m.last_quantise["dwf"])
        finally:
            m.close()
    (ta, qa, da), (tb, qb, db) = out[True], out[False]
    assert da == db and all(np.array_equal(qa[k], qb[k]) for k in qa)       # same integers -> same graph cut
    assert np.array_equal(ta[2], tb[2])
    for k in ta[1]:
        np.testing.assert_allclose(ta[1][k], tb[1][k], rtol=1e-12, atol=1e-13 * np.abs(tb[1][k]).max())
    np.testing.assert_allclose(ta[3:], tb[3:], rtol=1e-12)
    # one edge fewer: no longer the grid's list
    m = phyloHMRF(n_samples=len(X), n_features=d, observation=X, edge_list_1=[el[:-1]], len_vec=len_vec,
                  n_components=K, estimate_type=3, beta=1.0, beta1=0.1)
    try:
        assert not isinstance(m._regions[0], engine.GridRegion)
    finally:
        m.close()
