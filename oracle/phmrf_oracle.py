"""CPU oracle for the Phylo-HMRF E-step hot path -- TEST INFRASTRUCTURE ONLY.

This module is a Python-3 / NumPy restatement of the reference's per-region
E-step arithmetic.  It is the *checker* the CUDA path is compared against.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it; nothing under ``phylo_hmrf_b200/``
does (tests/test_abi.py::test_product_never_imports_oracle enforces that).

Parity anchoring
----------------
* The reference (``/root/reference``) is Python 2 and cannot be imported here;
  ``tests/golden/make_golden.py`` instead *executes the reference's own method
  bodies* (extracted in memory from ``phylo_hmrf.py`` / ``base.py`` with the py2
  ``print`` statements neutralised) and stores their inputs/outputs as fixtures
  under ``tests/golden/``.  ``tests/test_oracle_golden.py`` pins every function
  below that has a counterpart inside ``/root/reference`` against those
  fixtures.
* Two pieces of arithmetic live in third-party packages that are NOT in
  ``/root/reference`` and are not installed here: scikit-learn 0.18's
  ``log_multivariate_normal_density`` (README.md:80, call site
  phylo_hmrf.py:266-268) and yujiali/pygco's float->int cost conversion
  (README.md:84, unpinned; call site phylo_hmrf.py:496-498).  Their published
  algorithms are restated here (``log_multivariate_normal_density_full``,
  ``pygco_quantise``).  The density is cross-checked against
  ``scipy.stats.multivariate_normal.logpdf``; the quantiser is **parity
  unpinned** (no copy of pygco exists in this container) and is therefore
  defined as an explicit contract with named constants.

Every function cites the reference file:line it follows.  "faithful" variants
keep the reference's per-node Python loops (they are what the CPU baseline
times); "vec" variants are vectorised NumPy used for larger test sizes.  The
test-suite asserts that both agree.
"""
from __future__ import annotations

import numpy as np
from scipy import linalg

SMALL_EPS = 1e-16  # phylo_hmrf.py:49

# --- pygco (yujiali/pygco, pygco.py) constants: third-party, not vendored ---------
PYGCO_UNARY_FLOAT_PRECISION = 100000  # _UNARY_FLOAT_PRECISION
PYGCO_PAIRWISE_FLOAT_PRECISION = 1000  # _PAIRWISE_FLOAT_PRECISION (edge weights)
PYGCO_SMOOTH_COST_PRECISION = 100  # _SMOOTH_COST_PRECISION (label compatibility V): "pairwise * smooth = unary"
PYGCO_SMALL_CONSTANT = 1e-10  # _SMALL_CONSTANT


# ---------------------------------------------------------------------------------
# a1  emission log-likelihood
# ---------------------------------------------------------------------------------
def log_multivariate_normal_density_full(X, means, covars, min_covar=1.0e-7):
    """sklearn 0.18 ``_log_multivariate_normal_density_full`` (gmm.py), the function
    behind phylo_hmrf.py:266-268 with ``covariance_type='full'`` (phylo_hmrf.py:57).

    Per state: lower Cholesky (retry with ``+min_covar*I``), log-det from the diagonal,
    triangular solve of ``(X - mu).T``, ``-0.5*(sum(sol**2) + d*log(2*pi) + logdet)``.
    """
    X = np.asarray(X, dtype=np.float64)
    n_samples, n_dim = X.shape
    nmix = len(means)
    log_prob = np.empty((n_samples, nmix))
    for c, (mu, cv) in enumerate(zip(means, covars)):
        try:
            cv_chol = linalg.cholesky(cv, lower=True)
        except linalg.LinAlgError:
            try:
                cv_chol = linalg.cholesky(cv + min_covar * np.eye(n_dim), lower=True)
            except linalg.LinAlgError:
                raise ValueError("'covars' must be symmetric, positive-definite")
        cv_log_det = 2 * np.sum(np.log(np.diagonal(cv_chol)))
        cv_sol = linalg.solve_triangular(cv_chol, (X - mu).T, lower=True).T
        log_prob[:, c] = -0.5 * (np.sum(cv_sol**2, axis=1) + n_dim * np.log(2 * np.pi) + cv_log_det)
    return log_prob


def compute_log_likelihood(X, means, covars):
    """phyloHMRF._compute_log_likelihood, phylo_hmrf.py:266-268."""
    return log_multivariate_normal_density_full(X, means, covars)


# ---------------------------------------------------------------------------------
# a3 / a4  model-side constants
# ---------------------------------------------------------------------------------
def pairwise_potential(n_components, beta):
    """phyloHMRF._pairwise_potential, phylo_hmrf.py:524-536: Potts ``beta*(1-I)``."""
    V = np.zeros((n_components, n_components))
    for i in range(n_components):
        for j in range(i + 1, n_components):
            V[i, j] = beta
            V[j, i] = V[i, j]
    return V


def connected_edge(edge_ids, n_samples):
    """phyloHMRF._connected_edge, phylo_hmrf.py:674-689: per node, the incident edge
    indices in ascending edge order (list of lists)."""
    inc = [[] for _ in range(n_samples)]
    for e in range(len(edge_ids)):
        j, i = edge_ids[e][0], edge_ids[e][1]
        inc[i].append(e)
        inc[j].append(e)
    return inc


def edge_weight_undirected(edge_list, n_samples, beta1):
    """One region of phyloHMRF._edge_weight_undirected_vec, phylo_hmrf.py:567-598:
    ``w = exp(-beta1 * d_ij)`` (:585), ``ids = int64(edge_list[:, 0:2])`` (:589),
    incident-edge index (:591)."""
    edge_list = np.asarray(edge_list, dtype=np.float64)
    w = np.exp(-beta1 * edge_list[:, 2])
    ids = np.int64(edge_list[:, 0:2])
    return w, ids, connected_edge(ids, n_samples)


# ---------------------------------------------------------------------------------
# a5  pygco float -> int conversion (contract; third-party, parity unpinned)
# ---------------------------------------------------------------------------------
def pygco_down_weight_factor(unary, edge_weights, V):
    """``max(|unary|.max(), |w|.max() * V.max()) + 1e-10`` (pygco.cut_general_graph with
    ``down_weight_factor=None``, as called at phylo_hmrf.py:496-498)."""
    return max(np.abs(unary).max(), np.abs(edge_weights).max() * V.max()) + PYGCO_SMALL_CONSTANT


def pygco_quantise(unary, edge_weights, V, down_weight_factor=None, unary_precision=PYGCO_UNARY_FLOAT_PRECISION,
                   pairwise_precision=PYGCO_PAIRWISE_FLOAT_PRECISION, smooth_precision=PYGCO_SMOOTH_COST_PRECISION):
    """The three integer arrays pygco hands to GCO: divide by dwf, multiply by the
    precision constant, ``astype(np.intc)`` (C cast: truncation toward zero).
    Returns (unary_i32[N,K], w_i32[E], V_i32[K,K], dwf)."""
    unary = np.asarray(unary, dtype=np.float64)
    edge_weights = np.asarray(edge_weights, dtype=np.float64)
    V = np.asarray(V, dtype=np.float64)
    dwf = pygco_down_weight_factor(unary, edge_weights, V) if down_weight_factor is None else down_weight_factor
    u_i = ((unary / dwf) * unary_precision).astype(np.intc)
    w_i = ((edge_weights / dwf) * pairwise_precision).astype(np.intc)
    V_i = (V * smooth_precision).astype(np.intc)
    return u_i, w_i, V_i, dwf


def unary_boundary_mask(unary, dwf, tol=1e-9):
    """Entries of the scaled unary ``t=(u/dwf)*1e5`` that lie within relative ``tol`` of an
    integer (a truncation boundary): the only entries where two FP64 evaluations of the
    log-likelihood that agree to ``tol`` may produce different integers."""
    t = (np.asarray(unary, dtype=np.float64) / dwf) * PYGCO_UNARY_FLOAT_PRECISION
    dist = np.abs(t - np.rint(t))
    return dist <= tol * np.maximum(1.0, np.abs(t))


# ---------------------------------------------------------------------------------
# a7  neighbour-weighted pairwise potential
# ---------------------------------------------------------------------------------
def pairwise_compare_local_faithful(V, label, i, neighbor_edgeIdx, edge_w, edge_ids, estimate_type):
    """phyloHMRF._pairwise_compareLocal, phylo_hmrf.py:412-436."""
    idx = neighbor_edgeIdx[i]
    if len(idx) == 0:
        return V[label[i]]  # isolated node: unweighted row of V (:421-423)
    acc = np.zeros(V.shape[0])
    for k in idx:
        pair = edge_ids[k]
        other = pair[pair != i][0]
        s = label[other]
        if estimate_type == 3:
            acc = acc + V[s] * edge_w[k]
        else:
            acc = acc + V[s]
    return acc


def pairwise_compare_faithful(V, label, neighbor_edgeIdx, edge_w, edge_ids, estimate_type):
    """phyloHMRF._pairwise_compare, phylo_hmrf.py:398-410 (N Python calls)."""
    edge_ids = np.asarray(edge_ids)
    rows = []
    for i in range(len(label)):
        rows.append(pairwise_compare_local_faithful(V, label, i, neighbor_edgeIdx, edge_w, edge_ids, estimate_type))
    return np.asarray(rows)


def pairwise_compare_vec(V, label, edge_w, edge_ids, estimate_type):
    """Vectorised equivalent of :func:`pairwise_compare_faithful`.  Adds the id2-side
    contributions first and the id1-side second, which reproduces the faithful
    summation order for a (id1<id2, sorted) edge list (utility.py:1960)."""
    label = np.asarray(label, dtype=np.int64)
    N, K = len(label), V.shape[0]
    e = np.asarray(edge_ids, dtype=np.int64)
    pp = np.zeros((N, K))
    if len(e):
        w = np.asarray(edge_w, dtype=np.float64) if estimate_type == 3 else np.ones(len(e))
        a, b = e[:, 0], e[:, 1]
        np.add.at(pp, b, V[label[a]] * w[:, None])
        np.add.at(pp, a, V[label[b]] * w[:, None])
    deg = np.zeros(N, dtype=np.int64)
    if len(e):
        np.add.at(deg, e[:, 0], 1)
        np.add.at(deg, e[:, 1], 1)
    iso = deg == 0
    pp[iso] = V[label[iso]]
    return pp


# ---------------------------------------------------------------------------------
# a9  cost scalars
# ---------------------------------------------------------------------------------
def pairwise_compare_single_faithful(V, label, i, neighbor_edgeIdx, edge_w, edge_ids, estimate_type):
    """phyloHMRF._pairwise_compare_single, phylo_hmrf.py:449-468 (np.setdiff1d pairing)."""
    t_label = label[i]
    t_idx = neighbor_edgeIdx[i]
    ids = np.setdiff1d(edge_ids[t_idx].ravel(), i)
    states = np.asarray(label[ids])
    pot = V[states, t_label]
    if estimate_type == 3:
        pot = pot * edge_w[t_idx]
    return sum(pot)


def pairwise_compare_ensemble_faithful(V, label, neighbor_edgeIdx, edge_w, edge_ids, estimate_type):
    """phyloHMRF._pairwise_compare_ensemble, phylo_hmrf.py:438-447."""
    n = len(label)
    cost = np.zeros(n)
    for i in range(n):
        cost[i] = pairwise_compare_single_faithful(V, label, i, neighbor_edgeIdx, edge_w, edge_ids, estimate_type)
    return np.sum(cost) * 1.0 / n


def pairwise_compare_ensemble_vec(V, label, edge_w, edge_ids, estimate_type):
    label = np.asarray(label, dtype=np.int64)
    e = np.asarray(edge_ids, dtype=np.int64)
    n = len(label)
    if len(e) == 0:
        return 0.0
    w = np.asarray(edge_w, dtype=np.float64) if estimate_type == 3 else np.ones(len(e))
    la, lb = label[e[:, 0]], label[e[:, 1]]
    per_node = np.zeros(n)
    np.add.at(per_node, e[:, 1], V[la, lb] * w)
    np.add.at(per_node, e[:, 0], V[lb, la] * w)
    return np.sum(per_node) * 1.0 / n


def compute_cost_v1(label, logprob, pairwise_prob_normalize, pairwise_cost):
    """phyloHMRF._compute_cost_v1, phylo_hmrf.py:374-396 given the already computed raw
    pairwise cost (:377).  Uses the mask-multiply form of the reference (:380-386)."""
    n, K = logprob.shape
    mask = np.zeros((n, K))
    mask[np.arange(n), np.asarray(label, dtype=np.int64)] = 1
    lp = logprob.copy() * mask
    pwn = np.log(pairwise_prob_normalize + SMALL_EPS) * mask
    unary_cost = np.sum(lp, axis=1)
    unary_cost = -np.sum(unary_cost) * 1.0 / n
    pairwise_cost_normalize = -np.sum(pwn) * 1.0 / n
    cost1 = unary_cost + pairwise_cost_normalize
    return pairwise_cost, pairwise_cost_normalize, unary_cost, cost1


# ---------------------------------------------------------------------------------
# a8  posteriors
# ---------------------------------------------------------------------------------
def _naive_softmax(a):
    """``exp(a) / rowsum(exp(a))`` exactly as phylo_hmrf.py:342-345 / :347-350 (no
    max-subtraction)."""
    wp = np.exp(a)
    s = np.sum(wp, axis=1).reshape((-1, 1))
    return wp / np.dot(s, 1.0 * np.ones((1, a.shape[1])))


def _stable_softmax(a):
    m = np.max(a, axis=1, keepdims=True)
    wp = np.exp(a - m)
    return wp / np.sum(wp, axis=1, keepdims=True)


def compute_posteriors_graph(V, label, logprob, edge_w, edge_ids, neighbor_edgeIdx, estimate_type,
                             faithful=True, stable=False):
    """phyloHMRF._compute_posteriors_graph, phylo_hmrf.py:334-355.

    Returns (posteriors[N,K], pairwise_cost, pairwise_cost_normalize, unary_cost, cost1).
    ``stable=True`` swaps the naive softmax for a max-subtracted one (the specified
    behaviour of the CUDA path; equal wherever the naive form is finite).
    """
    label = np.asarray(label, dtype=np.int64)
    edge_ids = np.asarray(edge_ids, dtype=np.int64)
    if faithful:
        pp = pairwise_compare_faithful(V, label, neighbor_edgeIdx, edge_w, edge_ids, estimate_type)
        pc = pairwise_compare_ensemble_faithful(V, label, neighbor_edgeIdx, edge_w, edge_ids, estimate_type)
    else:
        pp = pairwise_compare_vec(V, label, edge_w, edge_ids, estimate_type)
        pc = pairwise_compare_ensemble_vec(V, label, edge_w, edge_ids, estimate_type)
    sm = _stable_softmax if stable else _naive_softmax
    posteriors = sm(logprob - pp)
    pwn = sm(-pp)
    costs = compute_cost_v1(label, logprob, pwn, pc)
    return (posteriors,) + costs


# ---------------------------------------------------------------------------------
# a10 / a11  sufficient statistics
# ---------------------------------------------------------------------------------
def sufficient_statistics(posteriors, X):
    """phylo_hmrf.py:311-314."""
    stats = dict()
    stats['post'] = posteriors.sum(axis=0)
    stats['obs'] = np.dot(posteriors.T, X)
    stats['obs*obs.T'] = np.einsum('ij,ik,il->jkl', posteriors, X, X)
    return stats


def initialize_sufficient_statistics(n_components, n_features):
    """base.py:562-569 + phylo_hmrf.py:691-698."""
    return {
        'nobs': 0,
        'start': np.zeros(n_components),
        'trans': np.zeros((n_components, n_components)),
        'post': np.zeros(n_components),
        'obs': np.zeros((n_components, n_features)),
        'obs**2': np.zeros((n_components, n_features)),
        'obs*obs.T': np.zeros((n_components, n_features, n_features)),
    }


def accumulate_sufficient_statistics_1(stats, stats1):
    """base.py:571-580."""
    stats['post'] += stats1['post']
    stats['obs'] += stats1['obs']
    stats['obs*obs.T'] += stats1['obs*obs.T']
    return stats


# ---------------------------------------------------------------------------------
# a12  one region of one EM iteration, minus the graph cut
# ---------------------------------------------------------------------------------
def estep_region(X, means, covars, V, edge_list, beta1, estimate_type, labels=None, faithful=False,
                 stable=False):
    """phylo_hmrf.py:297-322 for one region with the GCO call (phylo_hmrf.py:496-498)
    replaced by ``labels`` (default: arg-min of the integer unary, the bench stand-in of
    SURVEY 8(d)).  Returns a dict with every intermediate the CUDA path is checked on."""
    X = np.asarray(X, dtype=np.float64)
    N = len(X)
    w, ids, inc = edge_weight_undirected(edge_list, N, beta1) if faithful else (
        np.exp(-beta1 * np.asarray(edge_list)[:, 2]), np.int64(np.asarray(edge_list)[:, 0:2]), None)
    logprob = compute_log_likelihood(X, means, covars)
    unary = -logprob.copy()  # phylo_hmrf.py:490
    u_i, w_i, V_i, dwf = pygco_quantise(unary, w, V)
    if labels is None:
        labels = np.argmin(u_i, axis=1)
    post, c_pair, c_pair_norm, c_unary, c_total = compute_posteriors_graph(
        V, labels, logprob, w, ids, inc, estimate_type, faithful=faithful, stable=stable)
    stats = sufficient_statistics(post, X)
    return dict(logprob=logprob, unary_i32=u_i, w_i32=w_i, V_i32=V_i, dwf=dwf, labels=np.asarray(labels),
                posteriors=post, costs=(c_pair, c_pair_norm, c_unary, c_total), stats=stats,
                edge_w=w, edge_ids=ids)


# ---------------------------------------------------------------------------------
# synthetic geometry / features; shared by tests and bench
# ---------------------------------------------------------------------------------
def triangle_edges(B, row0=0, row1=None):
    """Undirected 8-neighbourhood edge ids of a B-bin diagonal region (row-major upper
    triangle incl. diagonal, utility.py:2310-2317; directions right, lower-right, lower,
    lower-left kept when inside the triangle, utility.py:1898-1931), sorted by (id1,id2)
    (utility.py:1960).  Rows restricted to [row0,row1) sources when given."""
    row1 = B if row1 is None else row1
    xs, ys = np.triu_indices(B)
    sel = (xs >= row0) & (xs < row1)
    xs, ys = xs[sel], ys[sel]

    def serial(x, y):  # index of (x,y), y>=x, in the row-major upper triangle
        return x * B - (x * (x - 1)) // 2 + (y - x)

    src = serial(xs, ys)
    out = []
    for dx, dy in ((0, 1), (1, 1), (1, 0), (1, -1)):
        nx, ny = xs + dx, ys + dy
        ok = (nx <= ny) & (ny < B) & (nx < B)
        out.append(np.stack([src[ok], serial(nx[ok], ny[ok])], axis=1))
    e = np.concatenate(out, axis=0)
    order = np.lexsort((e[:, 1], e[:, 0]))
    return e[order]


def edge_distances(X, e, B=None):
    """``d_ij = |x_i-x_j|^2 / (|x_i||x_j| + 1e-16)`` with diagonal-diagonal edges halved
    (utility.py:1919-1953)."""
    X = np.asarray(X)
    nrm = np.sqrt(np.sum(X * X, axis=1))
    a, b = e[:, 0], e[:, 1]
    d = np.sum((X[a] - X[b]) ** 2, axis=1) / (nrm[a] * nrm[b] + 1e-16)
    if B is not None:
        xs, ys = np.triu_indices(B)
        diag = xs == ys
        both = diag[a] & diag[b]
        d = np.where(both, 0.5 * d, d)
    return d
