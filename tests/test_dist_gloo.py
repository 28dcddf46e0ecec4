"""World-size-2 (gloo, CPU) test of the multi-GPU host logic: row-band sharding with halo,
the shared down-weight factor, and the all-reduce of statistics / cost sums.  The per-band
arithmetic is the CPU oracle here (no GPU in this tier); the same glue drives the CUDA path
in bench.py and tests/test_gpu_parity.py::test_row_bands_add_up_to_the_whole_region."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import phmrf_oracle as orc
from phylo_hmrf_b200 import synth
from phylo_hmrf_b200 import dist as pdist

B, D, K, ET, SEED = 40, 4, 6, 3, 31


def _band_oracle(g, means, covars, V, labels_window, dwf=None, et=ET):
    """What one rank computes for its band (oracle stand-in for the kernels): the window-sized label
    array lets the vectorised oracle see the halo labels; only the owned rows enter the sums."""
    o0, n = g["own_offset"], g["n_own"]
    X = g["X_own"]
    lp = orc.compute_log_likelihood(X, means, covars)
    pp = orc.pairwise_compare_vec(V, labels_window, g["edge_w"], g["edge_ids"], et)[o0:o0 + n]
    lab = labels_window[o0:o0 + n]
    post = orc._stable_softmax(lp - pp)
    pwn = orc._stable_softmax(-pp)
    st = orc.sufficient_statistics(post, X)
    e = g["edge_ids"]
    la, lb = labels_window[e[:, 0]], labels_window[e[:, 1]]
    per_node = np.zeros(g["n_window"])
    np.add.at(per_node, e[:, 1], V[la, lb] * g["edge_w"])
    np.add.at(per_node, e[:, 0], V[lb, la] * g["edge_w"])
    sums = np.array([per_node[o0:o0 + n].sum(), np.log(pwn[np.arange(n), lab] + 1e-16).sum(),
                     lp[np.arange(n), lab].sum()])
    flat = np.concatenate([st["post"], st["obs"].ravel(), st["obs*obs.T"].ravel()])
    return lp, flat, sums


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        whole = synth.make_band(SEED, B, D)
        means, covars = synth.model(SEED, whole["X_own"], K, D)
        V = synth.potts(K, 1.0)
        r0, r1 = synth.band_rows(B, world)[rank]
        g = synth.make_band(SEED, B, D, r0, r1)
        lp = orc.compute_log_likelihood(g["X_own"], means, covars)
        dwf = pdist.global_dwf(np.abs(lp).max(), np.abs(g["edge_w"]).max(), V.max(), dist)
        u_band = ((-lp / dwf) * 100000).astype(np.intc)
        # labels of the whole region: arg-min unary gathered from all bands (stand-in for GCO)
        parts = [None] * world
        dist.all_gather_object(parts, (g["win_start"] + g["own_offset"], np.argmin(u_band, axis=1)))
        labels = np.zeros(whole["n_own"], dtype=np.int64)
        for off, lab in parts:
            labels[off:off + len(lab)] = lab
        lab_win = labels[g["win_start"]:g["win_start"] + g["n_window"]]
        _, flat, sums = _band_oracle(g, means, covars, V, lab_win)
        stats, costs, n_total = pdist.combine_band_results(flat, sums, g["n_own"], dist)
        if rank == 0:
            np.savez(out, dwf=dwf, stats=stats, costs=np.asarray(costs), n_total=n_total, labels=labels)
    finally:
        dist.destroy_process_group()


def test_two_rank_band_sharding_matches_single_region(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    whole = synth.make_band(SEED, B, D)
    means, covars = synth.model(SEED, whole["X_own"], K, D)
    V = synth.potts(K, 1.0)
    lp = orc.compute_log_likelihood(whole["X_own"], means, covars)
    u, _, _, dwf = orc.pygco_quantise(-lp, whole["edge_w"], V)
    assert got["dwf"] == dwf
    labels = np.argmin(u, axis=1)
    assert np.array_equal(got["labels"], labels)
    ref = orc.compute_posteriors_graph(V, labels, lp, whole["edge_w"], whole["edge_ids"], None, ET, faithful=False,
                                       stable=True)
    st = orc.sufficient_statistics(ref[0], whole["X_own"])
    flat = np.concatenate([st["post"], st["obs"].ravel(), st["obs*obs.T"].ravel()])
    assert int(got["n_total"]) == whole["n_own"]
    np.testing.assert_allclose(got["stats"], flat, rtol=1e-12)
    np.testing.assert_allclose(got["costs"], ref[1:], rtol=1e-12)


def test_plan_shards_covers_every_row_once_and_balances():
    from phylo_hmrf_b200 import dist as pdist
    # BASELINE config 4: chr1..22 at 50 kb, one diagonal region per chromosome (bins from hg38.chrom.sizes)
    sizes = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
             133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
             58617616, 64444167, 46709983, 50818468]
    regions = [(1, s // 50000, s // 50000) for s in sizes]
    regions.append((0, 120, 340))          # and one off-diagonal block
    for world in (1, 2, 3, 8):
        plan = pdist.plan_shards(regions, world)
        assert len(plan) == world and plan == pdist.plan_shards(regions, world)
        seen = {}
        for rank, pieces in enumerate(plan):
            for rid, r0, r1, n in pieces:
                rows = pdist.region_row_sizes(*regions[rid])
                assert 0 <= r0 < r1 <= len(rows) and n == int(rows[r0:r1].sum())
                seen.setdefault(rid, []).append((r0, r1))
        for rid, reg in enumerate(regions):
            spans = sorted(seen[rid])
            assert spans[0][0] == 0 and spans[-1][1] == reg[1]
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        loads = [sum(p[3] for p in pieces) for pieces in plan]
        total = sum(int(pdist.region_row_sizes(*r).sum()) for r in regions)
        assert sum(loads) == total == 89321427 + 120 * 340
        assert max(loads) <= 1.06 * total / world
    # one big region on 8 GPUs: eight bands of (almost) equal node count, like bench.py's workload
    plan = pdist.plan_shards([(1, 24895, 24895)], 8)
    loads = [sum(p[3] for p in pieces) for pieces in plan]
    assert sorted(p[0][1:3] for p in plan)[0][0] == 0 and max(loads) - min(loads) < 2 * 24895
    assert pdist.split_rows([5, 1, 1], 3) == [(0, 1), (1, 2), (2, 3)]


def _em_worker(rank, world, port, out_dir):
    """Two ranks run the re-hosted EM driver on the scripted model: each executes only the regions it owns,
    the result tuples are all-gathered, and every rank must end with the reference driver's outcome."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import fit_script as fs
    from phylo_hmrf_b200.hmrf import phyloHMRF
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        owner = pdist.assign_regions([lv[0] for lv in fs.LEN_VEC], world)
        n_mine = sum(1 for o in owner if o == rank)

        class Driver(fs.ScriptedModel, phyloHMRF):
            def _predict_posteriors(self, X, len_vec, region_id, m_queue):
                assert owner[region_id] == rank          # a rank only ever works on its own regions
                it = self.iteration
                fs.ScriptedModel._predict_posteriors(self, X, len_vec, region_id, m_queue)
                # the script advances its clock after len(len_vec) calls; here a rank makes n_mine per iteration
                self.iteration, self.calls_this_iter = it, (self.calls_this_iter % len(len_vec))
                self._mine = getattr(self, "_mine", 0) + 1
                if self._mine == n_mine:
                    self._mine, self.calls_this_iter, self.iteration = 0, 0, it + 1
                return True

        name = "converges"
        m_iter, thr, _ = fs.SCENARIOS[name]
        m = object.__new__(Driver)
        m.script(name)
        res = m.fit_accumulate_test(np.zeros((fs.N, fs.D)), fs.LEN_VEC, thr, "test", m_iter, n_threads=1)
        np.savez(os.path.join(out_dir, "em_rank%d.npz" % rank), cost_vec=res[5], params_vec=res[0], t_labels=res[6],
                 it=np.array([res[3], res[4]]), n_iter=m.iteration, owner=np.array(owner))
    finally:
        dist.destroy_process_group()


def test_two_rank_em_driver_matches_the_reference_driver(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_em_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fit_driver.npz"))
    for rank in (0, 1):
        got = np.load(str(tmp_path / ("em_rank%d.npz" % rank)))
        assert sorted(got["owner"].tolist()) == [0, 1]
        assert int(got["n_iter"]) == int(gold["converges_n_iter"])
        np.testing.assert_array_equal(got["cost_vec"], gold["converges_cost_vec"])
        np.testing.assert_array_equal(got["params_vec"], gold["converges_params_vec"])
        np.testing.assert_array_equal(got["t_labels"], gold["converges_t_labels"])
        assert got["it"].tolist() == gold["converges_it"].tolist()


def test_assign_regions_is_balanced_and_deterministic():
    sizes = [100, 90, 50, 40, 30, 20, 10, 5]
    owner = pdist.assign_regions(sizes, 3)
    assert owner == pdist.assign_regions(sizes, 3) and set(owner) == {0, 1, 2}
    loads = [sum(s for s, o in zip(sizes, owner) if o == r) for r in range(3)]
    assert max(loads) - min(loads) <= 20
    assert pdist.assign_regions([7, 7], 1) == [0, 0]
