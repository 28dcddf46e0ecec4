"""the step immediately after the hot path -- the Ornstein-Uhlenbeck tree
algebra and the M-step objective that turn the K(1+d+d^2) sufficient statistics into new
`means_` / `_covars_`.  Host-side NumPy (K tiny d x d problems; the north star keeps them on
the host), written batched over the K states so that one objective evaluation serves every
state at once.

Reference: phylo_hmrf.py:715-919 (tree structure: `_initilize_tree_mtx`, `_search_leaf`,
`_search_ancestor`, `_compute_base_struct`, `_matrix1`), :985-1036
(`_ou_param_varied_constraint`), :1038-1138 (`_ou_lik_varied_constraint`), :1246-1325
(`_ou_lik_varied_single`), :1327-1425 (`_ou_optimize2*`, `_check_params`), :1427-1498
(initial fits), :184-264 (`_init_ou_param`, `_init`), :1500-1528 (`_do_mstep`).

Parameter vector of one state (length 3*node_num - 1, phylo_hmrf.py:105-107 and SURVEY
appendix B): [root variance | beta per branch | lambda per branch | theta per node].
"""
from __future__ import annotations

import sys

import numpy as np

small_eps = 1e-16  # phylo_hmrf.py:49


class OUTree(object):
    """Tree bookkeeping built once from the `edge.1.txt` rows (parent, child)."""

    def __init__(self, edge_list):
        edges = np.asarray(edge_list, dtype=np.int64).reshape(-1, 2)
        self.node_num = int(edges.max()) + 1
        # the reference stores tree_mtx[min, max] = 1 (phylo_hmrf.py:718-720): the smaller id is the parent
        self.tree_mtx = np.zeros((self.node_num, self.node_num))
        self.tree_mtx[edges.min(axis=1), edges.max(axis=1)] = 1
        parent = np.full(self.node_num, -1, dtype=np.int64)
        for child in range(self.node_num):
            above = np.flatnonzero(self.tree_mtx[:, child] > 0)
            if len(above):
                parent[child] = above[0]
        self.parent = parent
        self.branch_dim = self.node_num - 1
        self.n_params = self.node_num + 2 * self.branch_dim + 1
        self.leaf_vec = np.flatnonzero(self.tree_mtx.sum(axis=1) == 0)  # :855-865
        self.n_leaves = len(self.leaf_vec)
        self.leaf_list = {int(leaf): rank for rank, leaf in enumerate(self.leaf_vec)}  # :748-767
        # root-to-leaf paths, leaf included (:840-853)
        self.path_vec = []
        for leaf in self.leaf_vec:
            chain = [int(leaf)]
            while parent[chain[0]] >= 0:
                chain.insert(0, int(parent[chain[0]]))
            self.path_vec.append(np.asarray(chain))
        # leaf pairs: nearest common ancestor and the branches strictly below it (:866-919)
        n1 = self.n_leaves
        self.A1 = np.zeros((n1, self.node_num))
        self.A1[np.arange(n1), parent[self.leaf_vec]] = 1
        pairs, rows = [], []
        for i in range(n1):
            for j in range(i + 1, n1):
                common = np.intersect1d(self.path_vec[i], self.path_vec[j])
                row = np.zeros(self.node_num)
                row[np.setdiff1d(self.path_vec[i], common)] = 1
                row[np.setdiff1d(self.path_vec[j], common)] = 1
                rows.append(row)
                pairs.append([int(self.leaf_vec[i]), int(self.leaf_vec[j]), int(common.max())])
        self.pair_list = pairs
        self.A2 = np.asarray(rows).reshape(len(pairs), self.node_num)
        pl = np.asarray(pairs, dtype=np.int64).reshape(-1, 3)
        self._pair_i = np.asarray([self.leaf_list[a] for a in pl[:, 0]], dtype=np.int64)
        self._pair_j = np.asarray([self.leaf_list[b] for b in pl[:, 1]], dtype=np.int64)
        self._pair_anc = pl[:, 2]

    # ------------------------------------------------------------------ parameters -> moments
    def split(self, params):
        """[..., P] -> (root variance, beta [..., B], lambda [..., B], theta [..., B+1])."""
        B = self.branch_dim
        p = np.asarray(params, dtype=np.float64)
        return p[..., 0], p[..., 1:1 + B], p[..., 1 + B:1 + 2 * B], p[..., 1 + 2 * B:2 + 3 * B]

    def moments(self, params, guard_small_beta=True):
        """Per-node mean/variance recursion and the leaf covariance (without min_covar), for
        one parameter vector [P] or a batch [K,P] (:996-1031, :1056-1088).  Returns
        (values [..., node, 2], leaf means [..., d], leaf covariance [..., d, d])."""
        root_var, beta, lam, theta = self.split(params)
        if guard_small_beta:  # :1001-1003; `_ou_lik_varied_single` divides unguarded (:1257)
            ratio = np.where(beta > 1e-7, lam / np.where(beta > 1e-7, 2 * beta, 1.0), 0.0)
        else:
            with np.errstate(divide='ignore', invalid='ignore'):
                ratio = lam / (2 * beta)
        decay = np.exp(-beta)
        lead = np.zeros(beta.shape[:-1] + (1,))
        beta0 = np.concatenate([lead, beta], axis=-1)    # a zero "branch" into node 0
        decay0 = np.concatenate([lead, decay], axis=-1)
        ratio0 = np.concatenate([lead, ratio], axis=-1)
        values = np.zeros(beta.shape[:-1] + (self.node_num, 2))
        values[..., 0, 0] = theta[..., 0]
        values[..., 0, 1] = root_var
        for i in range(1, self.node_num):
            p = self.parent[i]
            values[..., i, 0] = values[..., p, 0] * decay0[..., i] + theta[..., i] * (1 - decay0[..., i])
            values[..., i, 1] = ratio0[..., i] * (1 - decay0[..., i] ** 2) + values[..., p, 1] * (decay0[..., i] ** 2)
        s1 = np.matmul(beta0, self.A2.T) if len(self.pair_list) else np.zeros(beta.shape[:-1] + (0,))
        s2 = values[..., self._pair_anc, 1] * np.exp(-s1)
        d = self.n_leaves
        cov = np.zeros(beta.shape[:-1] + (d, d))
        cov[..., self._pair_i, self._pair_j] = s2
        cov[..., self._pair_j, self._pair_i] = s2
        cov[..., np.arange(d), np.arange(d)] = values[..., self.leaf_vec, 1]
        return values, values[..., self.leaf_vec, 0], cov

    def check_params(self, params):
        """:1405-1425: 1 ok, -1 out of the [0,100] / [-100,100] box, -2 out of the box with NaN."""
        _, beta, lam, theta = self.split(params)
        ok1 = (beta >= 0) & (beta <= 1e2) & (lam >= 0) & (lam <= 1e2)
        ok2 = (theta >= -1e2) & (theta <= 1e2)
        if ok1.sum() < self.branch_dim or ok2.sum() < self.branch_dim + 1:
            return -2 if np.isnan(np.asarray(params)[1:]).any() else -1
        return 1


def _regularised_logdet_trace(V, S, weight, min_covar, logdet_eps):
    """`weight*log(det V + eps) + sum(inv(V) * S)` with the reference's conditioning ladder:
    add min_covar*I up to 10 times while cond(V) >= 1/eps_machine, then fall back to the
    pseudo-inverse (:1115-1131).  Returns (value, V actually used)."""
    d = V.shape[-1]
    if not (np.all(np.isfinite(V)) and np.all(np.isfinite(S))):
        # the optimiser wandered to non-finite parameters: LAPACK's SVD would spin (or raise) on
        # them; report an unusable point instead (the reference crashes here)
        return np.inf, V
    for _ in range(11):
        if np.linalg.cond(V) < 1 / sys.float_info.epsilon:
            return weight * np.log(np.linalg.det(V) + logdet_eps) + np.sum(np.linalg.inv(V) * S), V
        V = V + min_covar * np.eye(d)
    V = V - min_covar * np.eye(d)  # the 11th addition is not made by the reference (cnt<10 adds ten)
    return weight * np.log(np.linalg.det(V) + logdet_eps) + np.sum(np.linalg.pinv(V) * S), V


def mstep_objective(tree, params, state_id, stats, n_samples, lambda_0, min_covar, fallback_params=None):
    """`_ou_lik_varied_constraint` (:1038-1138) for one state.  Returns (lik, values, V) where V
    is the regularised leaf covariance the reference stores as `cv_mtx` (:1136)."""
    flag = tree.check_params(params)
    if flag <= -2 and fallback_params is not None:
        return mstep_objective(tree, fallback_params, state_id, stats, n_samples, lambda_0, min_covar)
    params = np.asarray(params, dtype=np.float64)
    values, mu, cov = tree.moments(params)
    V = cov + min_covar * np.eye(tree.n_leaves)
    c = state_id
    obsmean = np.outer(stats['obs'][c], mu)
    Sn_w = stats['obs*obs.T'][c] - obsmean - obsmean.T + np.outer(mu, mu) * stats['post'][c]
    core, V_used = _regularised_logdet_trace(V, Sn_w, stats['post'][c], min_covar, small_eps)
    lik = core / n_samples + lambda_0 * (1.0 / np.sqrt(n_samples)) * np.dot(params.T, params)
    return lik, values, V_used


def mstep_objective_batch(tree, params, stats, n_samples, lambda_0, min_covar):
    """All K states in one evaluation (the "batched M-step objective" of SURVEY 8 f-3): stacked
    moments, determinants and inverses; states whose covariance fails the conditioning test
    take the scalar ladder.  params [K,P] -> lik [K]."""
    params = np.asarray(params, dtype=np.float64)
    K = params.shape[0]
    _, mu, cov = tree.moments(params)
    d = tree.n_leaves
    V = cov + min_covar * np.eye(d)
    obsmean = stats['obs'][:, :, None] * mu[:, None, :]
    Sn_w = (stats['obs*obs.T'] - obsmean - obsmean.transpose(0, 2, 1)
            + mu[:, :, None] * mu[:, None, :] * stats['post'][:, None, None])
    finite = np.all(np.isfinite(V), axis=(1, 2)) & np.all(np.isfinite(Sn_w), axis=(1, 2))
    good = np.zeros(K, dtype=bool)
    if finite.any():
        good[finite] = np.linalg.cond(V[finite]) < 1 / sys.float_info.epsilon
    core = np.empty(K)
    if good.any():
        Vg = V[good]
        core[good] = (stats['post'][good] * np.log(np.linalg.det(Vg) + small_eps)
                      + np.sum(np.linalg.inv(Vg) * Sn_w[good], axis=(1, 2)))
    for c in np.flatnonzero(~good):
        core[c], _ = _regularised_logdet_trace(V[c], Sn_w[c], stats['post'][c], min_covar, small_eps)
    ridge = lambda_0 * (1.0 / np.sqrt(n_samples)) * np.sum(params * params, axis=1)
    return core / n_samples + ridge


def single_objective(tree, params, obs, min_covar):
    """`_ou_lik_varied_single` (:1246-1325): fit of one cluster's raw observations, used by the
    initialisation.  Returns (lik, values, covariance + min_covar*I)."""
    params = np.asarray(params, dtype=np.float64)
    values, mu, cov = tree.moments(params, guard_small_beta=False)
    d = tree.n_leaves
    V = cov + min_covar * np.eye(d)
    n = obs.shape[0]
    obsmean = np.outer(np.mean(obs, axis=0), mu)
    Sn_w = np.dot(obs.T, obs) / n - obsmean - obsmean.T + np.outer(mu, mu)
    Vt = V
    lik = np.nan
    if not (np.all(np.isfinite(V)) and np.all(np.isfinite(Sn_w))):
        return np.inf, values, V
    for _ in range(11):
        if np.linalg.cond(Vt) < 1 / sys.float_info.epsilon:
            lik = np.log(np.linalg.det(Vt)) + np.sum(np.linalg.inv(Vt) * Sn_w)
            break
        Vt = Vt + min_covar * np.eye(d)
    else:
        Vt = Vt - min_covar * np.eye(d)
        try:
            lik = np.log(np.linalg.det(Vt)) + np.sum(np.linalg.inv(Vt) * Sn_w)
        except np.linalg.LinAlgError:
            pass
    return lik, values, V


def init_guess(tree, mean_values, magnitude, rng):
    """`_ou_init_guess` (:1453-1480): random branch parameters, node optima filled bottom-up
    from the cluster's leaf means."""
    guess = magnitude * rng.random(tree.n_params)
    n1 = tree.node_num
    node_mean = np.zeros(n1)
    seen = np.zeros(n1)
    node_mean[tree.leaf_vec] = mean_values
    seen[tree.leaf_vec] = 2
    for j in range(n1 - 1, 0, -1):
        p = tree.parent[j]
        if seen[p] == 0:
            node_mean[p] = node_mean[j]
            seen[p] += 1
        elif seen[p] == 1:
            node_mean[p] = 0.5 * node_mean[p] + 0.5 * node_mean[j]
            seen[p] += 1
    guess[tree.n_params - n1:] = node_mean
    return guess


_BOX = ({'type': 'ineq', 'fun': lambda x: x - small_eps}, {'type': 'ineq', 'fun': lambda x: -x + 100})


def fit_cluster(tree, obs, mean_values, magnitude, min_covar, rng):
    """`_ou_optimize_init` (:1427-1451): SLSQP fit of one K-means cluster, retried up to 11
    times, random guess as the last resort."""
    from scipy.optimize import minimize
    params = None
    for _ in range(11):
        guess = init_guess(tree, mean_values, magnitude, rng)
        try:
            with np.errstate(over='ignore', invalid='ignore', divide='ignore'):   # probes outside the box
                res = minimize(lambda p: single_objective(tree, p, obs, min_covar)[0], guess, constraints=_BOX,
                               tol=1e-6, options={'disp': False})
        except Exception:
            continue
        params = res.x
        if tree.check_params(params) > 0:
            return params, single_objective(tree, params, obs, min_covar)[0]
    params = init_guess(tree, mean_values, magnitude, rng)
    return params, single_objective(tree, params, obs, min_covar)[0]


def optimise_state(tree, state_id, stats, n_samples, lambda_0, min_covar, init_params, current_params, w_init, w_cur,
                   magnitude, initial_mode, rng):
    """`_ou_optimize2` + `_ou_optimize2_unit` (:1327-1403): SLSQP from a mix of the initial
    estimate, the current estimate and a random vector; up to 11 attempts; the initial estimate
    is the fallback.  Returns (params, lik, values, V)."""
    from scipy.optimize import minimize

    def objective(p):
        # SLSQP probes points outside the [0,100] box the constraints describe (the reference's optimiser
        # does the same, phylo_hmrf.py:1365-1383): exp(-beta) overflows there and the objective comes back
        # inf/NaN, which the line search rejects -- evaluate quietly instead of warning on every probe
        with np.errstate(over='ignore', invalid='ignore', divide='ignore'):
            return mstep_objective(tree, p, state_id, stats, n_samples, lambda_0, min_covar, init_params)[0]

    for _ in range(11):
        if initial_mode == 1:
            rnd = 2 * rng.random(tree.n_params) - 1
            rnd[:-tree.node_num] = rng.random(tree.n_params - tree.node_num)
            rnd = magnitude * rnd
        else:
            rnd = magnitude * rng.random(tree.n_params)
        guess = w_init * init_params + w_cur * current_params + (1 - w_init - w_cur) * rnd
        try:
            with np.errstate(over='ignore', invalid='ignore', divide='ignore'):   # finite differences through inf probes
                res = minimize(objective, guess, method='SLSQP', constraints=_BOX, tol=1e-6, options={'disp': False})
        except Exception:
            continue
        if tree.check_params(res.x) > 0:
            lik, values, V = mstep_objective(tree, res.x, state_id, stats, n_samples, lambda_0, min_covar, init_params)
            return res.x, lik, values, V
    lik, values, V = mstep_objective(tree, init_params, state_id, stats, n_samples, lambda_0, min_covar)
    return init_params.copy(), lik, values, V


def attach(model, edge_list, initial_weight=0.3, initial_weight1=0.1, initial_magnitude=1, initial_mode=0, seed=None):
    """Install the OU initialisation / M-step / finalisation hooks on a `phyloHMRF` instance
    (the fork-free driver `fit_accumulate_test` calls them)."""
    tree = OUTree(edge_list)
    if tree.n_leaves != model.n_features:
        raise ValueError("the tree has %d leaves but the model has %d features" % (tree.n_leaves, model.n_features))
    rng = np.random.default_rng(seed)
    model.tree = tree
    model.n_params = tree.n_params

    def init_fn(m, X):
        from sklearn import cluster
        X = np.asarray(X)
        km = cluster.MiniBatchKMeans(n_clusters=m.n_components, random_state=seed, batch_size=2000, max_iter=1000,
                                     n_init=10)
        km.fit(X)
        m.means_ = km.cluster_centers_
        labels = km.labels_
        m.init_ou_params = initial_magnitude * rng.random((m.n_components, tree.n_params))
        for k in range(m.n_components):
            rows = np.flatnonzero(labels == k)
            if len(rows):
                m.init_ou_params[k], _ = fit_cluster(tree, X[rows], m.means_[k], initial_magnitude, m.min_covar, rng)
        m.params_vec1 = m.init_ou_params.copy()
        m.labels = np.int64(labels).copy()
        m.labels_local = m.labels.copy()
        cv = np.cov(X.T) + m.min_covar * np.eye(m.n_features)
        m._covars_ = np.tile(np.atleast_2d(cv), (m.n_components, 1, 1))

    def mstep_fn(m, stats):
        m.means_ = np.array(m.means_, dtype=np.float64)
        m._covars_ = np.array(m._covars_, dtype=np.float64)
        for c in range(m.n_components):
            params, lik, values, V = optimise_state(tree, c, stats, m.n_samples, m.lambda_0, m.min_covar,
                                                    m.init_ou_params[c], m.params_vec1[c], initial_weight,
                                                    initial_weight1, initial_magnitude, initial_mode, rng)
            m.lik = lik
            m.params_vec1[c] = params
            m.means_[c] = values[tree.leaf_vec, 0]
            m._covars_[c] = V + m.min_covar * np.eye(m.n_features)  # :1524 (V already carries one min_covar)

    def finalize_fn(m, params_vec):
        _, mu, cov = tree.moments(np.asarray(params_vec))
        m.means_ = mu.copy()
        m._covars_ = cov + m.min_covar * np.eye(m.n_features)

    model.init_fn, model.mstep_fn, model.finalize_fn = init_fn, mstep_fn, finalize_fn
    return tree
