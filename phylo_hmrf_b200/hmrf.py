"""``phyloHMRF`` -- the reference's model class re-hosted on the B200 library, hot path only.

Every method below keeps the name, argument order and return convention of the method it
replaces in the reference (file:line in each docstring), so code written against
``phylo_hmrf.py`` / ``base.py`` can call into this class unchanged for the per-region
E-step: emission -> integer cost arrays -> graph cut (host GCO) -> posteriors, cost
scalars and sufficient statistics.

The callers either side of the path (SURVEY 8 "next" rows) are re-hosted too: the fork-free EM
driver ``fit_accumulate_test`` (f-2) and, when the constructor is given the tree
(``edge_list``), the OU initialisation and SLSQP M-step of ``ou.py`` (f-3), which write
``means_`` / ``_covars_`` exactly where ``_do_mstep`` does (phylo_hmrf.py:1522-1524).  Without a
tree, ``means_`` / ``_covars_`` are plain attributes and ``init_fn`` / ``mstep_fn`` hooks can be
supplied by the caller.

There is no CPU fallback: all arithmetic of the hot path runs in ``libphmrf.so``.
"""
from __future__ import annotations

import numpy as np

from . import engine

small_eps = 1e-16  # phylo_hmrf.py:49


class _IncidentEdges:
    """List-of-lists view of ``_connected_edge`` (phylo_hmrf.py:674-689) backed by CSR
    arrays: ``inc[i]`` is the ascending list of edge indices incident to node ``i``."""

    def __init__(self, edge_ids, n_samples):
        e = np.asarray(edge_ids, dtype=np.int64).reshape(-1, 2)
        E = len(e)
        node = np.concatenate([e[:, 1], e[:, 0]]) if E else np.zeros(0, np.int64)
        eid = np.concatenate([np.arange(E), np.arange(E)]) if E else np.zeros(0, np.int64)
        order = np.lexsort((eid, node))
        self._eid = eid[order]
        self._off = np.zeros(n_samples + 1, dtype=np.int64)
        np.add.at(self._off, node + 1, 1)
        self._off = np.cumsum(self._off)
        self._n = n_samples

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        return self._eid[self._off[i]:self._off[i + 1]].tolist()


class phyloHMRF(object):
    """Hot-path subset of the reference's ``phyloHMRF(_BaseGraph)`` (phylo_hmrf.py:51-150)."""

    def __init__(self, n_samples, n_features, edge_list=None, branch_list=None, cons_param=1, beta=1.0, beta1=0.5,
                 initial_mode=0, initial_weight=0.3, initial_weight1=0.1, initial_magnitude=1, observation=None,
                 edge_list_1=None, len_vec=None, type_id=0, max_iter=10, n_components=1, run_id=0, estimate_type=0,
                 covariance_type='full', min_covar=1e-3, startprob_prior=1.0, transmat_prior=1.0, means_prior=0,
                 means_weight=0, covars_prior=1e-2, covars_weight=1, algorithm="viterbi", random_state=None,
                 n_iter=10, tol=1e-2, verbose=False, params="stmc", init_params="stmc", learning_rate=0.001,
                 device=0):
        if covariance_type != 'full':
            raise ValueError("the device path implements covariance_type='full' (phylo_hmrf.py:57)")
        self.n_components = int(n_components)
        self.n_features = int(n_features)
        self.n_samples = int(n_samples)
        self.run_id, self.estimate_type = run_id, int(estimate_type)
        self.covariance_type, self.min_covar = covariance_type, min_covar
        self.beta, self.beta1 = beta, beta1
        self.type_id, self.max_iter, self.n_iter, self.tol = type_id, max_iter, n_iter, tol
        self.lambda_0 = cons_param
        self.tree_edge_list, self.branch_params = edge_list, branch_list
        self.observation = observation
        self.edge_list_vec = edge_list_1
        self.len_vec = len_vec
        self.device = device

        self._model = engine.Model(self.n_components, self.n_features, device)
        self._regions = []
        self._model_key = None
        self._last_logprob = {}

        self.initial_mode, self.initial_w1, self.initial_w1a, self.initial_w2 = (
            initial_mode, initial_weight, initial_weight1, initial_magnitude)
        self.init_fn = self.mstep_fn = self.finalize_fn = None
        if edge_list is not None:
            # OU tree algebra, initialisation and M-step (host side, phylo_hmrf.py:715-1528)
            from . import ou
            ou.attach(self, edge_list, initial_weight, initial_weight1, initial_magnitude, initial_mode, random_state)

        self.edge_potential = self._pairwise_potential()
        if observation is not None and len_vec is not None and edge_list_1 is not None:
            (self.edge_weightList_undirected_vec, self.edge_idList_undirected_vec,
             self.neighbor_edgeIdx_vec) = self._edge_weight_undirected_vec(observation, len_vec, edge_list_1)
            n_total = int(sum(int(v[0]) for v in len_vec))
            self.labels = np.zeros(n_total, dtype=np.int64)
            self.labels_local = np.zeros(n_total, dtype=np.int64)

    # ------------------------------------------------------------------ model upkeep
    def _sync_model(self):
        """Push means_/_covars_/edge_potential to the device when they changed."""
        key = (self.means_.tobytes(), self._covars_.tobytes(), self.edge_potential.tobytes())
        if key != self._model_key:
            self._model.set_model(self.means_, self._covars_, self.edge_potential)
            self._model_key = key

    def _region_for(self, X, edge_ids, edge_w):
        """The resident region whose inputs are these very arrays, else a temporary one."""
        for r, reg in enumerate(self._regions):
            if edge_ids is self.edge_idList_undirected_vec[r] and len(X) == reg.n:
                return reg, False
        return self._model.region(X, edge_ids, edge_w), True

    # ------------------------------------------------------------------ constants
    def _pairwise_potential(self):
        """phylo_hmrf.py:524-536: Potts ``beta*(1-I)``."""
        K = self.n_components
        V = np.full((K, K), float(self.beta))
        np.fill_diagonal(V, 0.0)
        self.edge_potential = V
        return V

    def _connected_edge(self, edge_list_1, n_samples):
        """phylo_hmrf.py:674-689."""
        return _IncidentEdges(edge_list_1, n_samples)

    def _edge_weight_undirected_vec(self, X, len_vec, edge_list_vec):
        """phylo_hmrf.py:567-598.  Also makes each region resident on the device."""
        num_region = len(len_vec)
        w_vec, id_vec, inc_vec = [None] * num_region, [None] * num_region, [None] * num_region
        for reg in self._regions:
            reg.close()
        self._regions = []
        for r in range(num_region):
            n_samples, s1, s2 = int(len_vec[r][0]), int(len_vec[r][1]), int(len_vec[r][2])
            edge_list = np.asarray(edge_list_vec[r], dtype=np.float64)
            w = np.exp(-self.beta1 * edge_list[:, 2])
            ids = np.int64(edge_list[:, 0:2])
            w_vec[r], id_vec[r] = w, ids
            inc_vec[r] = self._connected_edge(ids, n_samples)
            self._regions.append(self._model.region(np.asarray(X)[s1:s2], ids, w))
        return w_vec, id_vec, inc_vec

    # ------------------------------------------------------------------ phase A
    def _compute_log_likelihood(self, X):
        """phylo_hmrf.py:266-268 (sklearn 0.18 'full' density) on the device."""
        self._sync_model()
        reg = self._model.region(X, np.zeros((0, 2), dtype=np.int64), np.zeros(0))
        try:
            reg.emit_loglik()
            return reg.logprob()
        finally:
            reg.close()

    def _estimate_state_graphcuts_gco(self, X, init_labels1, edge_idList_undirected, edge_weightList_undirected,
                                      want_logprob=True):
        """phylo_hmrf.py:486-507: emission, integer cost arrays (GPU), alpha-beta swap with
        5000 cycles from ``init_labels1`` (host GCO).  Returns (labels, logprob)."""
        self._sync_model()
        reg, temp = self._region_for(X, edge_idList_undirected, edge_weightList_undirected)
        try:
            reg.emit_loglik()
            q = reg.quantise()
            max_cycles1 = 5000
            labels = engine.gco_cut_int(q["unary_i32"], edge_idList_undirected, q["w_i32"], q["V_i32"],
                                        n_iter=max_cycles1, algorithm='swap', init_labels=init_labels1)
            self.last_quantise = q
            logprob = reg.logprob() if want_logprob else None
            if not temp:
                self._last_logprob[id(reg)] = logprob
            return labels, logprob
        finally:
            if temp:
                reg.close()

    def predict(self, X, region_id):
        """phylo_hmrf.py:470-484."""
        lv = self.len_vec[region_id]
        id1, id2 = int(lv[1]), int(lv[2])
        init_labels = self.labels_local[id1:id2].copy()
        state, logprob = self._estimate_state_graphcuts_gco(
            X, init_labels, self.edge_idList_undirected_vec[region_id], self.edge_weightList_undirected_vec[region_id])
        self.labels[id1:id2] = state
        return state, logprob

    # ------------------------------------------------------------------ phase B
    def _prepare_phase_b(self, reg, label, logprob):
        if logprob is not None and logprob is not self._last_logprob.get(id(reg)):
            reg.set_logprob(logprob)
            self._last_logprob[id(reg)] = logprob
        reg.set_labels(np.asarray(label))

    def _pairwise_compare(self, label, neighbor_edgeIdx, edge_weightList, edge_idList):
        """phylo_hmrf.py:398-410 -> pp [N,K]."""
        self._sync_model()
        n = len(label)
        reg, temp = self._region_for(np.zeros((n, self.n_features)), edge_idList, edge_weightList)
        try:
            if id(reg) not in self._last_logprob:
                reg.set_logprob(np.zeros((n, self.n_components)))
                if not temp:
                    self._last_logprob[id(reg)] = None
            reg.set_labels(np.asarray(label))
            return reg.pairwise_potential(self.estimate_type)
        finally:
            if temp:
                reg.close()

    def _compute_posteriors_graph(self, X, label, logprob, region_id):
        """phylo_hmrf.py:334-355 -> (posteriors, pairwise_cost, pairwise_cost_normalize,
        unary_cost, cost1)."""
        self._sync_model()
        reg = self._regions[region_id]
        self._prepare_phase_b(reg, label, logprob)
        stats, sums, post = reg.estep_stats(self.estimate_type, want_post=True)
        self._last_stats = stats
        return (post,) + engine.costs_from_sums(sums, reg.n)

    def _compute_cost_v1(self, X, label, logprob1, pairwise_prob_normalize, neighbor_edgeIdx, edge_weightList,
                         edge_idList):
        """phylo_hmrf.py:374-396 -> (pairwise_cost, pairwise_cost_normalize, unary_cost, cost1).
        ``pairwise_prob_normalize`` is recomputed on the device from (label, edges): it is a
        function of exactly those inputs at the reference's only call site (:352)."""
        self._sync_model()
        reg, temp = self._region_for(X, edge_idList, edge_weightList)
        try:
            reg.set_logprob(logprob1)
            if not temp:
                self._last_logprob[id(reg)] = logprob1
            reg.set_labels(np.asarray(label))
            _, sums, _ = reg.estep_stats(self.estimate_type)
            return engine.costs_from_sums(sums, reg.n)
        finally:
            if temp:
                reg.close()

    def _pairwise_compare_ensemble(self, label, neighbor_edgeIdx, edge_weightList, edge_idList):
        """phylo_hmrf.py:438-447 -> mean over nodes of the label-pair potential."""
        n = len(label)
        return self._compute_cost_v1(np.zeros((n, self.n_features)), label, np.zeros((n, self.n_components)), None,
                                     neighbor_edgeIdx, edge_weightList, edge_idList)[0]

    def _predict_posteriors(self, X, len_vec, region_id, m_queue):
        """phylo_hmrf.py:297-322: one region of one EM iteration; puts
        ``(region_id, stats, labels, pairwise_cost, pairwise_cost_normalize, unary_cost,
        cost1)`` on ``m_queue``.  The log-likelihood and the posteriors stay on the device."""
        self._sync_model()
        lv = len_vec[region_id]
        s1, s2 = int(lv[1]), int(lv[2])
        reg = self._regions[region_id]
        init_labels = self.labels_local[s1:s2].copy()
        labels, _ = self._estimate_state_graphcuts_gco(
            X[s1:s2], init_labels, self.edge_idList_undirected_vec[region_id],
            self.edge_weightList_undirected_vec[region_id], want_logprob=False)
        reg.set_labels(labels)
        stats, sums, _ = reg.estep_stats(self.estimate_type)
        c = engine.costs_from_sums(sums, reg.n)
        m_queue.put((region_id, stats, labels, c[0], c[1], c[2], c[3]))
        return True

    # ------------------------------------------------------------------ statistics helpers
    def _initialize_sufficient_statistics(self):
        """base.py:562-569 + phylo_hmrf.py:691-698."""
        K, d = self.n_components, self.n_features
        return {'nobs': 0, 'start': np.zeros(K), 'trans': np.zeros((K, K)), 'post': np.zeros(K),
                'obs': np.zeros((K, d)), 'obs**2': np.zeros((K, d)), 'obs*obs.T': np.zeros((K, d, d))}

    def _accumulate_sufficient_statistics_1(self, stats, stats1):
        """base.py:571-580."""
        stats['post'] += stats1['post']
        stats['obs'] += stats1['obs']
        stats['obs*obs.T'] += stats1['obs*obs.T']
        return stats

    # ------------------------------------------------------------------ EM driver
    def _init(self, X, lengths=None):
        """phylo_hmrf.py:205-264 (MiniBatchKMeans labels, per-cluster OU fit).  Out of the
        hot-path scope: supply `init_fn(model, X)` (it must set means_, _covars_, params_vec1,
        labels, labels_local) or subclass."""
        if getattr(self, "init_fn", None) is None:
            raise NotImplementedError("initialisation is outside the re-hosted hot path: set model.init_fn")
        self.init_fn(self, X)

    def _check(self):
        """base.py:515-538 / phylo_hmrf.py:152-182: parameter validation before fitting."""
        K, d = self.n_components, self.n_features
        self.means_ = np.asarray(self.means_)
        if self.means_.shape != (K, d):
            raise ValueError("means_ must have shape (n_components, n_features)")
        if np.asarray(self._covars_).shape != (K, d, d):
            raise ValueError("'full' covars must have shape (n_components, n_dim, n_dim)")

    def _do_mstep(self, stats):
        """phylo_hmrf.py:1500-1528 (K constrained SLSQP fits of the OU tree parameters).  Out of
        the hot-path scope: supply `mstep_fn(model, stats)` (it must update means_, _covars_
        and params_vec1) or subclass."""
        if getattr(self, "mstep_fn", None) is None:
            raise NotImplementedError("the OU M-step is outside the re-hosted hot path: set model.mstep_fn")
        self.stats = stats.copy()
        self.mstep_fn(self, stats)

    def _ou_param_varied_constraint(self, params_vec):
        """phylo_hmrf.py:985-1036: writes means_/_covars_ for the final parameters; delegated
        to `finalize_fn(model, params_vec)` when given."""
        if getattr(self, "finalize_fn", None) is not None:
            self.finalize_fn(self, params_vec)

    def fit_accumulate_test(self, X, len_vec, threshold, annotation, m_iter, lengths=None, n_threads=None):
        """base.py:301-455, without fork.  Same iteration, cost aggregation (N_r/N weights,
        :332-337, :384-396), convergence tests (:402-435) and best-iteration bookkeeping;
        regions run through `_predict_posteriors` on a thread pool (the GPU kernels and the
        host graph cut release the GIL), results are gathered from a queue exactly as the
        parent process does.  Returns (params_vec, params_vec1, params_vecList, iter_id1,
        iter_id2, cost_vec, t_labels)."""
        try:
            import queue as _queue
        except ImportError:  # py2
            import Queue as _queue
        from concurrent.futures import ThreadPoolExecutor

        self._init(X, lengths=lengths)
        self._check()
        max_iter = m_iter
        max_iter1 = 50  # iterations after the previous minimum
        pairwise_cost_pre, unary_cost_pre, cost1_pre = 0.001, 0.001, 0.001
        threshold1, threshold2 = threshold, threshold
        cost_vec = []
        min_cost = [0, 1000]
        min_cost1 = [0, 1000]
        params_vec = self.params_vec1.copy()
        params_vec1 = self.params_vec1.copy()
        num_region = len(len_vec)
        ratio_vec = np.zeros(num_region)
        for i in range(0, num_region):
            ratio_vec[i] = len_vec[i][0]
        n_samples = int(sum(ratio_vec))
        ratio_vec = ratio_vec * 1.0 / n_samples
        params_vecList = []
        t_labels = np.zeros(n_samples)
        workers = n_threads or min(num_region, 8)
        # one process per GPU (torchrun): whole regions are dealt to the ranks by node count, every rank
        # gathers all result tuples and then runs the same (deterministic) aggregation and M-step
        from . import dist as _pdist
        world, rank = _pdist.world_info()
        if world > 1:
            owner = _pdist.assign_regions([lv[0] for lv in len_vec], world)
            my_regions = [r for r in range(num_region) if owner[r] == rank]
        else:
            my_regions = list(range(num_region))

        for iter in range(max_iter):
            stats = self._initialize_sufficient_statistics()
            self._sync_model()
            self.queue = _queue.Queue()
            if workers > 1 and len(my_regions) > 1:
                with ThreadPoolExecutor(max_workers=workers) as pool:
                    list(pool.map(lambda r: self._predict_posteriors(X, len_vec, r, self.queue), my_regions))
            else:
                for region_id in my_regions:
                    self._predict_posteriors(X, len_vec, region_id, self.queue)
            results = [self.queue.get() for _ in range(len(my_regions))]
            if world > 1:
                results = _pdist.all_gather_results(results)

            pairwise_cost1, pairwise_cost, unary_cost, cost1 = 0, 0, 0, 0
            id1 = 3
            labels = np.zeros(n_samples)
            # fixed (region id) order so that the floating-point sums do not depend on thread timing
            for vec1 in sorted(results, key=lambda v: v[0]):
                region_id = vec1[0]
                pairwise_cost1 += vec1[id1] * ratio_vec[region_id]
                pairwise_cost += vec1[id1 + 1] * ratio_vec[region_id]
                unary_cost += vec1[id1 + 2] * ratio_vec[region_id]
                cost1 += vec1[id1 + 3] * ratio_vec[region_id]
                s1, s2 = len_vec[region_id][1], len_vec[region_id][2]
                stats = self._accumulate_sufficient_statistics_1(stats, vec1[1])
                labels[s1:s2] = vec1[2]

            t_difference1 = abs((pairwise_cost - pairwise_cost_pre) * 1.0 / pairwise_cost_pre)
            t_difference2 = abs((unary_cost - unary_cost_pre) * 1.0 / unary_cost_pre)
            t_difference3 = abs((cost1 - cost1_pre) * 1.0 / cost1_pre)
            pairwise_cost_pre, unary_cost_pre, cost1_pre = pairwise_cost, unary_cost, cost1
            cost_vec.append([iter, pairwise_cost, unary_cost, cost1])
            params_vecList.append(self.params_vec1.copy())
            self.labels = labels.copy()

            if cost1 < min_cost[1]:
                min_cost = [iter, cost1]
                params_vec = self.params_vec1.copy()
                self.labels_local = self.labels.copy()  # current local optimal state estimate
            if cost1 < min_cost1[1] and iter >= 3:
                min_cost1 = [iter, cost1]
                params_vec1 = self.params_vec1.copy()
                t_labels = self.labels.copy()  # keep the estimated labels
            if ((t_difference1 < threshold1 and t_difference2 < threshold2) or (t_difference3 < threshold1)) and (iter > 5):
                break
            if iter > max_iter:
                break
            if iter - min_cost1[0] > max_iter1:
                break
            self._do_mstep(stats)

        self.params_vec1 = params_vec1.copy()
        self._ou_param_varied_constraint(params_vec)
        cost_vec = np.asarray(cost_vec)
        params_vecList = np.asarray(params_vecList)
        return params_vec, params_vec1, params_vecList, min_cost[0], min_cost1[0], cost_vec, t_labels

    def close(self):
        for reg in self._regions:
            reg.close()
        self._regions = []
        self._model.close()
