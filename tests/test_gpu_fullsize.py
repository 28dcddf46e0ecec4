"""Full-size checks (BASELINE.json config 3: synthetic chr1 @50 kb, 12 397 710 nodes, d=5, K=20)
through size-independent properties, plus oracle parity on a random sample of rows.

* sum_k post_k == N and the implied means/second moments are consistent (posterior rows sum
  to one; statistics are linear in the posteriors);
* obs*obs.T is symmetric; statistics of the three row bands add up to the whole region;
* the integer unary reproduces numpy's conversion on sampled rows; arg-min labels agree;
* repeat runs are bit-identical (fixed-order reductions).
"""
import numpy as np
import pytest

from oracle import phmrf_oracle as orc

pytestmark = pytest.mark.gpu

B, D, K, ET, SEED = 4979, 5, 20, 3, 20261020


@pytest.fixture(scope="module")
def setup():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import phylo_hmrf_b200 as ph
    from phylo_hmrf_b200 import synth
    g = synth.make_band(SEED, B, D)
    assert g["n_own"] == 12397710
    means, covars = synth.model(SEED, g["X_own"][:200000], K, D)
    V = synth.potts(K, 1.0)
    m = ph.Model(K, D)
    m.set_model(means, covars, V)
    reg = m.region(g["X_own"], g["edge_ids"], g["edge_w"])
    yield ph, synth, g, means, covars, V, m, reg
    reg.close()
    m.close()


def test_config3_properties_and_sampled_parity(setup):
    ph, synth, g, means, covars, V, m, reg = setup
    n = g["n_own"]
    X = g["X_own"]
    absmax = reg.emit_loglik(want_absmax=True)
    q = reg.quantise(want_unary=True, want_edges=True, boundary_cap=1 << 20)
    assert q["dwf"] == max(absmax, np.abs(g["edge_w"]).max() * V.max()) + 1e-10
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(n, size=100000, replace=False))
    lp_ref = orc.compute_log_likelihood(X[rows], means, covars)
    assert absmax >= np.abs(lp_ref).max() * (1 - 1e-12)
    u_ref = ((-lp_ref / q["dwf"]) * 100000).astype(np.intc)
    u_gpu = q["unary_i32"][rows]
    diff = np.argwhere(u_gpu != u_ref)
    listed = set(q["boundary_idx"].tolist())
    assert q["n_boundary"] <= (1 << 20)
    for r, k in diff:
        assert int(rows[r]) * K + int(k) in listed and abs(int(u_gpu[r, k]) - int(u_ref[r, k])) == 1
    assert len(diff) <= 5  # boundary hits are ~1e-9 * |t| rare
    lab = reg.labels_argmin_unary()
    assert np.array_equal(lab[rows], np.argmin(u_gpu, axis=1))
    stats, sums, _ = reg.estep_stats(ET)
    # posterior rows sum to one
    np.testing.assert_allclose(stats["post"].sum(), n, rtol=1e-11)
    np.testing.assert_allclose(stats["obs"].sum(axis=0), X.sum(axis=0), rtol=1e-10)
    np.testing.assert_allclose(stats["obs*obs.T"].sum(axis=0), X.T @ X, rtol=1e-10)
    for k in range(K):
        assert np.array_equal(stats["obs*obs.T"][k], stats["obs*obs.T"][k].T)
    # cost scalars: unary part equals the mean log-likelihood at the labels on the sample
    c = ph.costs_from_sums(sums, n)
    assert np.isfinite(c).all() and c[0] >= 0 and c[1] >= 0
    lp_lab = lp_ref[np.arange(len(rows)), lab[rows]]
    assert abs(-lp_lab.mean() - c[2]) < 0.05 * abs(c[2]) + 0.05  # sample estimate of the full mean
    # determinism
    stats2, sums2, _ = reg.estep_stats(ET)
    assert all(np.array_equal(stats[k], stats2[k]) for k in stats) and np.array_equal(sums, sums2)
    # row bands add up (shared dwf, halo labels)
    tot, tot_sums = None, np.zeros(3)
    for r0, r1 in synth.band_rows(B, 3):
        gb = synth.make_band(SEED, B, D, r0, r1)
        b = m.region(gb["X_own"], gb["edge_ids"], gb["edge_w"], n_window=gb["n_window"], own_offset=gb["own_offset"])
        b.emit_loglik()
        b.set_labels(lab[gb["win_start"]:gb["win_start"] + gb["n_window"]])
        st, su, _ = b.estep_stats(ET)
        tot = st if tot is None else {k: tot[k] + st[k] for k in st}
        tot_sums += su
        b.close()
    for k in stats:
        np.testing.assert_allclose(tot[k], stats[k], rtol=1e-11, atol=1e-9)
    np.testing.assert_allclose(tot_sums, sums, rtol=1e-11)
