"""Host-side input preparation (synthetic contact maps, band sharding) against the oracle's
edge builder, which tests/test_oracle_golden.py pins to the reference's utility.py."""
import numpy as np
import pytest

from oracle import phmrf_oracle as orc
from phylo_hmrf_b200 import synth


@pytest.mark.parametrize("B", [1, 2, 3, 7, 33])
def test_whole_triangle_matches_oracle_edges(B):
    g = synth.triangle_band(B)
    assert g["n_own"] == g["n_window"] == B * (B + 1) // 2 and g["own_offset"] == 0
    xs, ys = np.triu_indices(B)
    assert np.array_equal(g["x"], xs) and np.array_equal(g["y"], ys)
    assert np.array_equal(g["edge_ids"].reshape(-1, 2), orc.triangle_edges(B).reshape(-1, 2))


def test_edge_distances_match_oracle():
    B, d = 21, 4
    g = synth.make_band(3, B, d)
    ref = orc.edge_distances(g["X_window"], g["edge_ids"], B)
    np.testing.assert_allclose(g["edge_dist"], ref, rtol=1e-13)
    assert (g["X_window"] == 0).mean() > 0.15


@pytest.mark.parametrize("n_bands", [2, 3, 5])
def test_bands_partition_nodes_and_cover_incident_edges(n_bands):
    B, d = 40, 3
    whole = synth.make_band(11, B, d)
    E_whole = {tuple(e) for e in whole["edge_ids"]}
    rows = synth.band_rows(B, n_bands)
    assert rows[0][0] == 0 and rows[-1][1] == B
    total = 0
    seen_incident = 0
    for r0, r1 in rows:
        g = synth.make_band(11, B, d, r0, r1)
        total += g["n_own"]
        o0 = g["win_start"] + g["own_offset"]
        # features are a pure function of the node
        np.testing.assert_array_equal(g["X_own"], whole["X_window"][o0:o0 + g["n_own"]])
        glob = g["edge_ids"] + g["win_start"]
        assert {tuple(e) for e in glob} <= E_whole
        assert np.all(np.diff(glob[:, 0]) >= 0)
        own = (glob >= o0) & (glob < o0 + g["n_own"])
        assert own.any(axis=1).all()
        # every whole-region edge touching an owned node is present
        inc = [(a, b) for (a, b) in E_whole if o0 <= a < o0 + g["n_own"] or o0 <= b < o0 + g["n_own"]]
        assert len(inc) == len(glob)
        seen_incident += own.sum()
        w_ref = {tuple(e): w for e, w in zip(whole["edge_ids"], whole["edge_w"])}
        np.testing.assert_array_equal(g["edge_w"], [w_ref[tuple(e)] for e in glob])
    assert total == whole["n_own"]
    assert seen_incident == 2 * len(whole["edge_ids"])


def test_model_generator_is_spd():
    g = synth.make_band(5, 30, 9)
    means, covars = synth.model(5, g["X_own"], 30, 9)
    assert means.shape == (30, 9) and covars.shape == (30, 9, 9)
    for c in covars:
        assert np.allclose(c, c.T) and np.linalg.eigvalsh(c).min() > 1e-4
