"""Host staging and concurrent regions: page-locked buffers from the C ABI, staged quantise / label buffers,
and two regions driven from two threads (the product's thread pool; the library lets one bulk transfer per
direction run at a time) giving exactly the results of driving them one after the other."""
import threading

import numpy as np
import pytest

from phylo_hmrf_b200 import synth

pytestmark = pytest.mark.gpu


def _setup(ph, seed, B, d, K):
    g = synth.make_band(seed, B, d)
    means, covars = synth.model(seed, g["X_own"], K, d)
    return g, means, covars


def test_pinned_arrays_and_staged_buffers():
    import phylo_hmrf_b200 as ph
    a = ph.engine.pinned_empty((1000, 7), np.int32)
    a[:] = 5
    assert a.shape == (1000, 7) and a.dtype == np.int32 and int(a.sum()) == 35000
    g, means, covars = _setup(ph, 3, 40, 4, 6)
    m = ph.Model(6, 4, device=0)
    try:
        m.set_model(means, covars, synth.potts(6, 1.0))
        reg = m.region(g["X_own"], g["edge_ids"], g["edge_w"])
        reg.emit_loglik()
        q_plain = reg.quantise()
        q_staged = reg.quantise(staged=True)
        assert np.array_equal(q_plain["unary_i32"], q_staged["unary_i32"])
        assert np.array_equal(q_plain["w_i32"], q_staged["w_i32"]) and q_plain["dwf"] == q_staged["dwf"]
        again = reg.quantise(staged=True)
        assert again["unary_i32"] is q_staged["unary_i32"]          # the region's own buffer, reused
        lab = reg.label_staging()
        lab[:] = np.argmin(q_plain["unary_i32"], axis=1)
        out = ph.gco_cut_int(q_plain["unary_i32"], g["edge_ids"], q_plain["w_i32"], q_plain["V_i32"], n_iter=20,
                             algorithm='swap', init_labels=lab.copy(), out=lab)
        assert out is lab
        reg.set_labels(lab)
        s1, c1, _ = reg.estep_stats(3)
        reg.set_labels(lab.copy())                                   # pageable copy: same result
        s2, c2, _ = reg.estep_stats(3)
        assert all(np.array_equal(s1[k], s2[k]) for k in s1) and np.array_equal(c1, c2)
        with pytest.raises(ValueError):
            ph.gco_cut_int(q_plain["unary_i32"], g["edge_ids"], q_plain["w_i32"], q_plain["V_i32"],
                           out=np.zeros(3, np.int32))
        reg.close()
    finally:
        m.close()


def test_two_regions_from_two_threads_equal_the_serial_run():
    import phylo_hmrf_b200 as ph
    K, d = 8, 5
    ga, means, covars = _setup(ph, 5, 300, d, K)
    gb, _, _ = _setup(ph, 6, 260, d, K)
    m = ph.Model(K, d, device=0)
    try:
        m.set_model(means, covars, synth.potts(K, 1.0))
        regs = [m.region(g["X_own"], g["edge_ids"], g["edge_w"]) for g in (ga, gb)]

        def one_pass(reg, g, out, reps):
            for _ in range(reps):
                reg.update_X(g["X_own"])
                reg.emit_loglik()
                q = reg.quantise(staged=True)
                lab = np.argmin(q["unary_i32"], axis=1).astype(np.int32)
                reg.set_labels(lab)
                st, sums, _ = reg.estep_stats(3)
                out.append((q["unary_i32"].copy(), q["w_i32"].copy(), q["dwf"], st, sums))

        serial = [[], []]
        for r, g, o in zip(regs, (ga, gb), serial):
            one_pass(r, g, o, 1)
        threaded = [[], []]
        ths = [threading.Thread(target=one_pass, args=(r, g, o, 4)) for r, g, o in zip(regs, (ga, gb), threaded)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        for ref, got in zip(serial, threaded):
            assert len(got) == 4
            for res in got:                                       # bit-identical, every repetition
                assert np.array_equal(res[0], ref[0][0]) and np.array_equal(res[1], ref[0][1]) and res[2] == ref[0][2]
                assert all(np.array_equal(res[3][k], ref[0][3][k]) for k in res[3])
                assert np.array_equal(res[4], ref[0][4])
        for r in regs:
            r.close()
    finally:
        m.close()
