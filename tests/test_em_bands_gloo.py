"""The EM driver with ONE region cut into row bands over two ranks (gloo, CPU): emission per band,
shared down-weight factor, integer unary gathered on the owner, real GCO swap there, label windows
sent back, E-step per band, one all-reduce -- against the same problem run whole in one process.
The per-band arithmetic is the CPU oracle (no GPU in this tier); tests/test_gpu_em_bands.py runs the
same comparison with the library on two GPUs over NCCL."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import em_band_case as case  # noqa: E402
from test_dist_gloo import _band_oracle  # noqa: E402

from oracle import phmrf_oracle as orc  # noqa: E402
from phylo_hmrf_b200 import em, engine, synth  # noqa: E402
from phylo_hmrf_b200.hmrf import phyloHMRF  # noqa: E402

B, D, K, M_ITER = 20, 3, 4, 9
B_SMALL = 7          # a second, small region for the mixed (banded + whole) plan


class OracleModel(phyloHMRF):
    """phyloHMRF with the device work replaced by the oracle (constructor bypassed: no GPU here)."""

    def setup(self, X, len_vec, edge_list_vec):
        self.n_components, self.n_features, self.estimate_type = K, D, case.ET
        self.beta, self.beta1 = case.BETA, case.BETA1
        self.edge_potential = synth.potts(K, case.BETA)
        self.X = X
        self.edge_idList_undirected_vec = [np.int64(e[:, :2]) for e in edge_list_vec]
        self.edge_weightList_undirected_vec = [np.exp(-case.BETA1 * e[:, 2]) for e in edge_list_vec]
        self.init_fn, self.mstep_fn, self.finalize_fn = case.init_fn, case.mstep_fn, None

    def _sync_model(self):
        pass

    def _predict_posteriors(self, X, len_vec, rid, q):          # a whole region on one process
        ids, w, V = self.edge_idList_undirected_vec[rid], self.edge_weightList_undirected_vec[rid], self.edge_potential
        s1, s2 = int(len_vec[rid][1]), int(len_vec[rid][2])
        Xr = X[s1:s2]
        lp = orc.compute_log_likelihood(Xr, self.means_, self._covars_)
        u, wi, Vi, _ = orc.pygco_quantise(-lp, w, V)
        labels = engine.gco_cut_int(u, ids, wi, Vi, n_iter=5000, algorithm='swap',
                                    init_labels=self.labels_local[s1:s2].astype(np.int32))
        ref = orc.compute_posteriors_graph(V, labels, lp, w, ids, None, case.ET, faithful=False, stable=True)
        q.put((rid, orc.sufficient_statistics(ref[0], Xr), labels) + tuple(ref[1:]))
        return True

    def _prepare_bands(self, X, len_vec, banded, comm):
        self._g = {}
        for rid, plist in banded.items():
            for bi, (rank, r0, r1) in enumerate(plist):
                if rank == comm.rank:
                    self._g[(rid, bi)] = synth.make_band(self._region_seeds[rid], self._region_bins[rid], D, r0, r1,
                                                         beta1=case.BETA1)

    def _band_emit(self, rid, bi):
        g = self._g[(rid, bi)]
        g["lp"] = orc.compute_log_likelihood(g["X_own"], self.means_, self._covars_)
        return float(np.abs(g["lp"]).max())

    def _band_quantise(self, rid, bi, dwf):
        return ((-self._g[(rid, bi)]["lp"] / dwf) * 100000).astype(np.intc)

    def _region_edge_costs(self, rid, dwf):
        _, wi, Vi, _ = orc.pygco_quantise(np.zeros((1, K)), self.edge_weightList_undirected_vec[rid],
                                          self.edge_potential, down_weight_factor=dwf)
        return wi, Vi

    def _band_estep(self, rid, bi, labels_window):
        _, flat, sums = _band_oracle(self._g[(rid, bi)], self.means_, self._covars_, self.edge_potential,
                                     np.asarray(labels_window, dtype=np.int64))
        return flat, sums


def _problem(two_regions):
    X, len_vec, edge_list_vec = case.problem(B, D)
    if two_regions:       # region 1: a small triangle appended after the big one
        g = synth.make_band(case.SEED + 1, B_SMALL, D, beta1=case.BETA1)
        n0, n1 = len(X), g["n_own"]
        X = np.concatenate([X, g["X_own"]])
        len_vec = len_vec + [[n1, n0, n0 + n1, B_SMALL, B_SMALL, 0, 0, 1, 1, 2]]
        edge_list_vec = edge_list_vec + [np.column_stack([g["edge_ids"].astype(np.float64), g["edge_dist"]])]
    return X, len_vec, edge_list_vec


def _run(two_regions=False):
    X, len_vec, edge_list_vec = _problem(two_regions)
    m = object.__new__(OracleModel)
    m.setup(X, len_vec, edge_list_vec)
    m._region_seeds = [case.SEED, case.SEED + 1]
    m._region_bins = [B, B_SMALL]
    res = m.fit_accumulate_test(X, len_vec, 1e-12, "test", M_ITER, n_threads=1)
    return m, res


def _worker(rank, world, port, out_dir, two_regions=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, res = _run(two_regions)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), cost_vec=res[5], t_labels=res[6], params=res[0],
                 means=m.means_, labels_local=m.labels_local)
    finally:
        dist.destroy_process_group()


def test_plan_cuts_the_region_into_two_bands():
    _, len_vec, _ = case.problem(B, D)
    whole, banded = em.make_plan(len_vec, 2)
    assert whole == [[], []] and sorted(banded) == [0]
    (ra, a0, a1), (rb, b0, b1) = banded[0]
    assert {ra, rb} == {0, 1} and a0 == 0 and a1 == b0 and b1 == B
    assert em.make_plan(len_vec, 1) == ([[0]], {})
    # a len_vec without grid geometry keeps regions whole
    assert em.make_plan([[7, 0, 7, 0, 0, 0, 0, 0, 1, 21], [5, 7, 12, 0, 0, 0, 0, 1, 1, 22]], 2)[1] == {}


def test_two_rank_banded_em_matches_the_single_process_run(tmp_path):
    m1, res1 = _run()
    assert len(res1[5]) >= 7                      # several iterations incl. M-steps and label updates
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for rank in (0, 1):
        got = np.load(str(tmp_path / ("rank%d.npz" % rank)))
        np.testing.assert_allclose(got["cost_vec"], res1[5], rtol=1e-10)
        np.testing.assert_array_equal(got["t_labels"], res1[6])
        np.testing.assert_array_equal(got["labels_local"], m1.labels_local)
        np.testing.assert_allclose(got["params"], res1[0], rtol=1e-10)
        np.testing.assert_allclose(got["means"], m1.means_, rtol=1e-10)


def test_three_ranks_with_a_banded_and_a_whole_region(tmp_path):
    """Mixed plan on three ranks: the big region is cut into bands, the small one stays whole on one rank; the
    point-to-point exchanges of the banded region and the single all-reduce must interleave without a deadlock
    and reproduce the single-process run."""
    _, len_vec, _ = _problem(True)
    whole, banded = em.make_plan(len_vec, 3)
    assert sorted(banded) == [0] and len(banded[0]) >= 2 and sum(len(w) for w in whole) == 1
    m1, res1 = _run(True)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(3, port, str(tmp_path), True), nprocs=3, join=True)
    for rank in range(3):
        got = np.load(str(tmp_path / ("rank%d.npz" % rank)))
        np.testing.assert_allclose(got["cost_vec"], res1[5], rtol=1e-10)
        np.testing.assert_array_equal(got["t_labels"], res1[6])
        np.testing.assert_array_equal(got["labels_local"], m1.labels_local)
        np.testing.assert_allclose(got["means"], m1.means_, rtol=1e-10)


def test_plan_of_the_genome_wide_configuration():
    """BASELINE config 4 on 8 ranks through the EM driver's planner: chr1 and chr2 hold more than 1/8 of the
    nodes each and are cut into two bands, every other autosome stays whole, the load is balanced to 6 %."""
    bins = synth.autosome_bins(50000)
    len_vec, s = [], 0
    for r, b in enumerate(bins):
        n = b * (b + 1) // 2
        len_vec.append([n, s, s + n, b, b, 0, 0, r, 1, r + 1])
        s += n
    whole, banded = em.make_plan(len_vec, 8)
    assert sorted(banded) == [0, 1] and all(len(banded[r]) == 2 for r in banded)
    assert sorted(r for w in whole for r in w) == list(range(2, 22))
    load = [sum(len_vec[r][0] for r in w) for w in whole]
    for rid in banded:
        assert banded[rid][0][1] == 0 and banded[rid][0][2] == banded[rid][1][1] and banded[rid][1][2] == bins[rid]
        for rank, r0, r1 in banded[rid]:
            own0, own1, _, _ = em.band_window(1, bins[rid], bins[rid], r0, r1)
            load[rank] += own1 - own0
    assert sum(load) == s == 89321427 and max(load) <= 1.06 * s / 8
