"""Row a5: the float -> int conversion pygco applies before the graph cut.  pygco is third-party, not
vendored and unpinned (README.md:84), so the conversion is a stated contract with three named scale
factors (phmrf_set_quantiser).  Default: unary 1e5, edge weights 1e3, label compatibility 1e2 -- pygco's
own rule "pairwise * smooth = unary" (1e3 * 1e2 = 1e5), which puts w_ij * V[a,b] on the unary's scale.
The alternative smooth factor 1e3 (an edge term ten times stronger) is covered too, and a fixture dumped
from a real pygco install ("Pinning the quantiser") is checked when present."""
import os

import numpy as np
import pytest

from oracle import phmrf_oracle as orc
from phylo_hmrf_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "pygco_fixture.npz")


def test_default_scale_rule_pairwise_times_smooth_equals_unary():
    assert orc.PYGCO_PAIRWISE_FLOAT_PRECISION * orc.PYGCO_SMOOTH_COST_PRECISION == orc.PYGCO_UNARY_FLOAT_PRECISION
    # with the rule in force an edge term w*V and a unary term u of equal float size get equal integers
    u, w, V = np.array([[0.37, 0.0]]), np.array([0.37]), np.array([[0.0, 1.0], [1.0, 0.0]])
    u_i, w_i, V_i, dwf = orc.pygco_quantise(u, w, V, down_weight_factor=1.0)
    assert u_i[0, 0] == w_i[0] * V_i[0, 1] == 37000


@pytest.mark.parametrize("smooth", [100, 1000])
def test_oracle_contract_for_both_smooth_factors(smooth):
    rng = np.random.default_rng(5)
    u, w = rng.random((50, 4)) * 30, rng.random(80)
    V = synth.potts(4, 1.7)
    u_i, w_i, V_i, dwf = orc.pygco_quantise(u, w, V, smooth_precision=smooth)
    assert dwf == max(np.abs(u).max(), np.abs(w).max() * V.max()) + 1e-10
    assert np.array_equal(u_i, np.trunc((u / dwf) * 100000).astype(np.int32))
    assert np.array_equal(w_i, np.trunc((w / dwf) * 1000).astype(np.int32))
    assert np.array_equal(V_i, np.trunc(V * smooth).astype(np.int32)) and V_i[0, 1] == int(1.7 * smooth)


@pytest.mark.skipif(not os.path.exists(FIXTURE), reason="no pygco fixture (see INTEGRATION.md)")
def test_oracle_matches_a_fixture_dumped_from_a_real_pygco():
    f = np.load(FIXTURE)
    u_i, w_i, V_i, _ = orc.pygco_quantise(f["unary"], f["edge_weights"], f["pairwise"],
                                          smooth_precision=float(f["smooth_precision"]))
    assert np.array_equal(u_i, f["unary_i"]) and np.array_equal(w_i, f["edge_weights_i"])
    assert np.array_equal(V_i, f["pairwise_i"])


@pytest.mark.gpu
@pytest.mark.parametrize("smooth", [100, 1000])
def test_device_quantiser_follows_the_context_scale_factors(smooth):
    import phylo_hmrf_b200 as ph
    B, d, K = 30, 4, 5
    g = synth.make_band(11, B, d)
    means, covars = synth.model(11, g["X_own"], K, d)
    V = synth.potts(K, 1.3)
    m = ph.Model(K, d, device=0)
    try:
        assert m.quantiser() == (100000.0, 1000.0, 100.0)          # the defaults
        m.set_quantiser(smooth_precision=smooth)
        m.set_model(means, covars, V)
        reg = m.region(g["X_own"], g["edge_ids"], g["edge_w"])
        reg.emit_loglik()
        lp = reg.logprob()
        q = reg.quantise(boundary_cap=lp.size)
        u_ref, w_ref, V_ref, dwf = orc.pygco_quantise(-lp, g["edge_w"], V, smooth_precision=smooth)
        assert q["dwf"] == dwf
        assert np.array_equal(q["unary_i32"], u_ref) and np.array_equal(q["w_i32"], w_ref)
        assert np.array_equal(q["V_i32"], V_ref) and q["V_i32"][0, 1] == int(1.3 * smooth)
        # the graph cut sees the difference: same unary, edge terms ten times apart
        lab = ph.gco_cut_int(q["unary_i32"], g["edge_ids"], q["w_i32"], q["V_i32"], n_iter=50, algorithm='swap',
                             init_labels=np.argmin(q["unary_i32"], axis=1).astype(np.int32), return_energy=True)
        assert lab[1] <= lab[2]                                    # the swap never raises the energy
        with pytest.raises(ph._lib.PhmrfError):
            m.set_quantiser(smooth_precision=0.5)
        reg.close()
    finally:
        m.close()
