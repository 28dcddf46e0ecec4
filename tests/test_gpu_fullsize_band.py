"""The bench workload itself -- one row band of BASELINE.json's headline configuration (synthetic chr1
@10 kb, 24 895 bins, d=9, K=30; rows [0,1608) = 38 739 132 nodes, built on the device like bench.py does)
-- checked through size-independent properties:

* posterior rows sum to one: sum_k post_k == n; the statistics are linear in the posteriors, so
  sum_k obs_k == sum_i x_i and sum_k obs*obs.T_k == X^T X;
* obs*obs.T is symmetric per state; repeat runs are bit-identical;
* sampled rows of the integer unary reproduce NumPy's conversion of the oracle's log-likelihood
  (the arg-min labels follow);
* the two halves of the band (as separate device-built bands sharing the down-weight factor and the
  halo labels) add up to the whole band's statistics.
"""
import numpy as np
import pytest

from oracle import phmrf_oracle as orc

pytestmark = pytest.mark.gpu

B, D, K, ET, SEED, BETA1 = 24895, 9, 30, 3, 20261022, 0.1
R0, R1 = 0, 1608


def test_headline_band_properties():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~45 GB of device memory")
    import phylo_hmrf_b200 as ph
    from phylo_hmrf_b200 import synth

    g0 = synth.make_band(SEED, B, D, 0, 24, beta1=BETA1)
    means, covars = synth.model(SEED, g0["X_own"], K, D)
    V = synth.potts(K, 1.0)
    m = ph.Model(K, D)
    m.set_model(means, covars, V)

    def build(r0, r1):
        g = synth.window_xy(B, r0, r1)
        Xw = synth.features(SEED, g["x"], g["y"], D)
        reg = m.region_grid(Xw, 1, B, B, r0, r1, 8, BETA1)
        return reg, Xw, g["n_own"], g["n_window"], g["own_offset"]

    reg, Xw, n, n_window, off = build(R0, R1)
    assert n == 38739132
    X = Xw[off:off + n]
    absmax = reg.emit_loglik(want_absmax=True)
    q = reg.quantise(want_unary=True, want_edges=False, boundary_cap=1 << 20)
    # window labels (owned + halo row) the way bench.py obtains them
    win = m.region(Xw, np.zeros((0, 2), np.int64), np.zeros(0))
    win.emit_loglik()
    win.quantise(want_unary=False, want_edges=False)
    labels_window = win.labels_argmin_unary()
    win.close()
    # sampled integer parity
    rng = np.random.default_rng(9)
    rows = np.sort(rng.choice(n, size=50000, replace=False))
    lp_ref = orc.compute_log_likelihood(X[rows], means, covars)
    assert absmax >= np.abs(lp_ref).max() * (1 - 1e-12)
    u_ref = ((-lp_ref / q["dwf"]) * 100000).astype(np.intc)
    u_gpu = q["unary_i32"][rows]
    listed = set(q["boundary_idx"].tolist())
    for r, k in np.argwhere(u_gpu != u_ref):
        assert int(rows[r]) * K + int(k) in listed and abs(int(u_gpu[r, k]) - int(u_ref[r, k])) == 1
    assert np.array_equal(labels_window[off:off + n][rows], np.argmin(u_gpu, axis=1))
    del q, u_gpu
    reg.set_labels(labels_window)
    stats, sums, _ = reg.estep_stats(ET)
    np.testing.assert_allclose(stats["post"].sum(), n, rtol=1e-11)
    np.testing.assert_allclose(stats["obs"].sum(axis=0), X.sum(axis=0), rtol=1e-10)
    np.testing.assert_allclose(stats["obs*obs.T"].sum(axis=0), X.T @ X, rtol=1e-10)
    for k in range(K):
        assert np.array_equal(stats["obs*obs.T"][k], stats["obs*obs.T"][k].T)
    stats2, sums2, _ = reg.estep_stats(ET)
    assert all(np.array_equal(stats[k], stats2[k]) for k in stats) and np.array_equal(sums, sums2)
    wmax = reg.weight_max()
    reg.close()
    del Xw, X

    # two half bands: same region-wide quantities, halo labels from the whole band's window
    tot, tot_sums = None, np.zeros(3)
    mid = (R0 + R1) // 2
    for r0, r1 in ((R0, mid), (mid, R1)):
        b, Xb, nb, nwb, offb = build(r0, r1)
        b.set_weight_max(wmax)
        b.emit_loglik()
        first = synth.tri_row_start(B, max(r0 - 1, 0))          # global index of the half's first window node
        base = synth.tri_row_start(B, max(R0 - 1, 0))
        b.set_labels(labels_window[first - base:first - base + nwb])
        s, c, _ = b.estep_stats(ET)
        tot = s if tot is None else {k: tot[k] + s[k] for k in s}
        tot_sums += c
        b.close()
        del Xb
    for k in stats:
        np.testing.assert_allclose(tot[k], stats[k], rtol=1e-10, atol=1e-6)
    np.testing.assert_allclose(tot_sums, sums, rtol=1e-10)
    m.close()
