"""GPU edge-list builder against fixtures produced by the reference's own
utility.py builders (tests/golden/edges_cases.npz) and against the oracle at a larger size."""
import os

import numpy as np
import pytest

from oracle import phmrf_oracle as orc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "edges_cases.npz")


@pytest.fixture(scope="module")
def util():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from phylo_hmrf_b200 import utility
    return utility


def test_edge_builders_match_reference_fixtures(util):
    g = np.load(GOLD)
    for c in range(int(g["n_cases"])):
        kind, n1, n2, nn, d = (int(v) for v in g["c%d_meta" % c])
        X, serial, ref = g["c%d_X" % c], g["c%d_serial" % c], g["c%d_edge_list" % c]
        if kind == 1:
            el = util.edge_weightlist_grid3_undirected_unsym(X, serial, n2, '', nn)
        else:
            el = util.edge_weightlist_grid3_undirected(X, serial, (n1, n2), '', nn)
        assert el.shape == ref.shape, (c, el.shape, ref.shape)
        assert np.array_equal(el[:, :2], ref[:, :2]), "edge ids / order differ in case %d" % c
        np.testing.assert_allclose(el[:, 2], ref[:, 2], rtol=1e-12, atol=1e-300)


def test_edge_builder_large_triangle_matches_oracle(util):
    from phylo_hmrf_b200 import synth
    B, d = 700, 9
    g = synth.make_band(4, B, d)
    el = util.edge_weightlist_grid3_undirected_unsym(g["X_own"], synth.tri_row_start(B, g["x"]) * 0 + g["x"] * B + g["y"],
                                                     B, '', 8)
    e = orc.triangle_edges(B)
    assert np.array_equal(np.int64(el[:, :2]), e)
    np.testing.assert_allclose(el[:, 2], orc.edge_distances(g["X_own"], e, B), rtol=1e-12, atol=1e-300)


def test_sparse_regions_are_refused(util):
    X = np.zeros((5, 2))
    with pytest.raises(NotImplementedError):
        util.edge_weightlist_grid3_undirected(X, np.asarray([0, 1, 2, 4, 5]), (2, 3), '', 8)
    with pytest.raises(ValueError):
        util.edge_weightlist_grid3_undirected(np.zeros((6, 2)), np.arange(6), (2, 3), '', 10)
