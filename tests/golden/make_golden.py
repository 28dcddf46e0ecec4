"""Generate tests/golden/*.npz by running the reference's own code (see ref_loader.py).

Run in the build container only:  ``python tests/golden/make_golden.py``

What comes from where
---------------------
* ``ref_*`` arrays are produced by the reference's method bodies executed verbatim:
  ``_pairwise_potential``, ``_edge_weight_undirected_vec``, ``_connected_edge``,
  ``_pairwise_compare``, ``_compute_posteriors_graph`` (+ ``_compute_cost_v1``,
  ``_pairwise_compare_ensemble``/``_single``), ``_predict_posteriors`` (statistics
  triple and the queue tuple), ``predict`` / ``_estimate_state_graphcuts_gco`` (the
  arguments handed to ``pygco.cut_general_graph``), base.py's statistics helpers and
  utility.py's edge-list builders.
* ``logprob`` comes from the restated sklearn-0.18 density (the real one is not
  installable here); it is an *input* to the reference code above, and is pinned
  separately against scipy in tests/test_oracle_golden.py.
* labels returned by the pygco stub are chosen by this script (arg-min of the unary with
  a few seeded flips) -- GCO itself is exercised in tests/test_gco.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from oracle import phmrf_oracle as orc  # noqa: E402


class _Recorder:
    """Stand-in for the ``pygco`` module: records the call, returns preset labels."""

    def __init__(self):
        self.calls = []
        self.next_labels = None

    def cut_general_graph(self, edges, edge_weights, unary_cost, pairwise_cost, n_iter=-1, algorithm='expansion',
                          init_labels=None, down_weight_factor=None):
        self.calls.append(dict(edges=np.array(edges), edge_weights=np.array(edge_weights),
                               unary_cost=np.array(unary_cost), pairwise_cost=np.array(pairwise_cost),
                               n_iter=n_iter, algorithm=algorithm, init_labels=np.array(init_labels),
                               down_weight_factor=down_weight_factor))
        return self.next_labels(unary_cost)


class _Queue:
    def __init__(self):
        self.items = []

    def put(self, x):
        self.items.append(x)


def synth_features(rng, N, d, zero_frac=0.3):
    base = rng.gamma(2.0, 0.6, size=(N, 1))
    z = base + 0.5 * rng.standard_normal((N, d))
    X = np.log1p(np.maximum(z, 0.0))
    X[rng.random((N, d)) < zero_frac * 0.5] = 0.0  # per-entry zeros
    X[rng.random(N) < zero_frac * 0.3] = 0.0  # whole-node zeros (empty bins in every species)
    return X


def synth_model(rng, K, d):
    means = rng.uniform(0.0, 2.0, size=(K, d))
    covars = np.empty((K, d, d))
    for k in range(K):
        A = rng.standard_normal((d, d)) * 0.4
        covars[k] = A @ A.T + (0.05 + 0.5 * rng.random()) * np.eye(d) + 1e-3 * np.eye(d)
    return means, covars


def region_geometry(util, rng, kind, n1, n2, d):
    """Build (X, edge_list) for one region through the reference's own edge builders."""
    if kind == "diag":
        serial = np.asarray([i * n2 + j for i in range(n1) for j in range(i, n2)])
        X = synth_features(rng, len(serial), d)
        el = util["edge_weightlist_grid3_undirected_unsym"](X, serial, n2, '', 8)
    else:
        serial = np.asarray([i * n2 + j for i in range(n1) for j in range(n2)])
        X = synth_features(rng, len(serial), d)
        el = util["edge_weightlist_grid3_undirected"](X, serial, (n1, n2), '', 8)
    return X, np.asarray(el, dtype=np.float64)


def flatten_ragged(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    for i, l in enumerate(lists):
        off[i + 1] = off[i] + len(l)
    flat = np.asarray([x for l in lists for x in l], dtype=np.int64)
    return flat, off


def make_case(name, seed, regions, d, K, beta, beta1, estimate_type, isolate=()):
    rng = np.random.default_rng(seed)
    util = ref_loader.load_utility(["mapping_Idx", "_sort_array", "edge_weightlist_grid3_undirected_unsym",
                                    "edge_weightlist_grid3_undirected"])
    rec = _Recorder()
    cls = ref_loader.build_reference_class({
        "pygco": rec,
        "log_multivariate_normal_density":
            lambda X, m, c, t: orc.log_multivariate_normal_density_full(X, m, c),
    })

    Xs, els, len_vec = [], [], []
    s = 0
    for r, (kind, n1, n2) in enumerate(regions):
        X, el = region_geometry(util, rng, kind, n1, n2, d)
        if r in dict(isolate):
            node = dict(isolate)[r]
            el = el[(el[:, 0] != node) & (el[:, 1] != node)]
        Xs.append(X)
        els.append(el)
        len_vec.append([len(X), s, s + len(X), n1, n2, 0, 0, r, 1 if kind == "diag" else 0, 21])
        s += len(X)
    X_all = np.concatenate(Xs, axis=0)
    means, covars = synth_model(rng, K, d)

    m = object.__new__(cls)
    m.n_components, m.n_features = K, d
    m.beta, m.beta1, m.estimate_type = beta, beta1, estimate_type
    m.covariance_type = 'full'
    m.means_, m._covars_ = means, covars
    m.len_vec = len_vec
    m.edge_potential = m._pairwise_potential()
    (m.edge_weightList_undirected_vec, m.edge_idList_undirected_vec,
     m.neighbor_edgeIdx_vec) = m._edge_weight_undirected_vec(X_all, len_vec, els)
    m.labels_local = rng.integers(0, K, size=len(X_all)).astype(np.int64)
    m.labels = m.labels_local.copy()

    flip_rng = np.random.default_rng(seed + 1)

    def choose(unary):
        lab = np.argmin(unary, axis=1)
        flips = flip_rng.random(len(lab)) < 0.15
        lab[flips] = flip_rng.integers(0, K, size=int(flips.sum()))
        return lab.astype(np.int64)

    rec.next_labels = choose

    out = dict(d=d, K=K, beta=beta, beta1=beta1, estimate_type=estimate_type, n_regions=len(regions),
               X=X_all, means=means, covars=covars, len_vec=np.asarray(len_vec, dtype=np.int64),
               ref_V=m.edge_potential, init_labels=m.labels_local.copy())
    q = _Queue()
    stats_total = m._initialize_sufficient_statistics()
    for r in range(len(regions)):
        m._predict_posteriors(X_all, len_vec, r, q)
        rid, stats, labels, c_pair, c_pair_norm, c_unary, c_total = q.items[-1]
        call = rec.calls[-1]
        s1, s2 = len_vec[r][1], len_vec[r][2]
        logprob = -call["unary_cost"]
        pp = m._pairwise_compare(labels, m.neighbor_edgeIdx_vec[r], m.edge_weightList_undirected_vec[r],
                                 m.edge_idList_undirected_vec[r])
        post = m._compute_posteriors_graph(X_all[s1:s2], labels, logprob, r)[0]
        flat, off = flatten_ragged(m.neighbor_edgeIdx_vec[r])
        p = "r%d_" % r
        out.update({
            p + "edge_list": els[r],
            p + "ref_edge_w": m.edge_weightList_undirected_vec[r],
            p + "ref_edge_ids": m.edge_idList_undirected_vec[r],
            p + "ref_inc_flat": flat, p + "ref_inc_off": off,
            p + "logprob": logprob,
            p + "ref_gco_unary": call["unary_cost"], p + "ref_gco_V": call["pairwise_cost"],
            p + "ref_gco_w": call["edge_weights"], p + "ref_gco_edges": call["edges"],
            p + "ref_gco_init": call["init_labels"], p + "ref_gco_n_iter": call["n_iter"],
            p + "ref_gco_algorithm": call["algorithm"],
            p + "ref_gco_dwf_is_none": call["down_weight_factor"] is None,
            p + "labels": np.asarray(labels, dtype=np.int64),
            p + "ref_pp": pp, p + "ref_post": post,
            p + "ref_costs": np.asarray([c_pair, c_pair_norm, c_unary, c_total]),
            p + "ref_stats_post": stats["post"], p + "ref_stats_obs": stats["obs"],
            p + "ref_stats_obsobsT": stats["obs*obs.T"],
        })
        stats_total = m._accumulate_sufficient_statistics_1(stats_total, stats)
    out.update(ref_total_post=stats_total["post"], ref_total_obs=stats_total["obs"],
               ref_total_obsobsT=stats_total["obs*obs.T"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "N =", len(X_all), "E =", [len(e) for e in els])


def make_edge_cases():
    """Pure geometry fixtures: the reference's two edge-list builders on dense regions."""
    util = ref_loader.load_utility(["mapping_Idx", "_sort_array", "edge_weightlist_grid3_undirected_unsym",
                                    "edge_weightlist_grid3_undirected"])
    rng = np.random.default_rng(77)
    out = {}
    # (the reference crashes when a direction has no candidate at all, e.g. a 2-bin triangle:
    # mapping_Idx indexes an empty array, utility.py:846 -- so the smallest shapes here are 3 / 2x3)
    cases = [("tri", 13, 13, 8, 4), ("tri", 9, 9, 4, 3), ("tri", 3, 3, 8, 2), ("rect", 6, 8, 8, 5),
             ("rect", 5, 7, 4, 2), ("rect", 2, 3, 8, 3)]
    for c, (kind, n1, n2, nn, d) in enumerate(cases):
        if kind == "tri":
            serial = np.asarray([i * n2 + j for i in range(n1) for j in range(i, n2)])
            X = synth_features(rng, len(serial), d)
            el = util["edge_weightlist_grid3_undirected_unsym"](X, serial, n2, '', nn)
        else:
            serial = np.asarray([i * n2 + j for i in range(n1) for j in range(n2)])
            X = synth_features(rng, len(serial), d)
            el = util["edge_weightlist_grid3_undirected"](X, serial, (n1, n2), '', nn)
        out["c%d_meta" % c] = np.asarray([1 if kind == "tri" else 0, n1, n2, nn, d])
        out["c%d_X" % c] = X
        out["c%d_serial" % c] = serial
        out["c%d_edge_list" % c] = np.asarray(el, dtype=np.float64).reshape(-1, 3)
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(HERE, "edges_cases.npz"), **out)
    print("wrote edges_cases", [len(out["c%d_edge_list" % c]) for c in range(len(cases))])


def make_fit_driver_cases():
    """Run the REFERENCE's fit_accumulate_test (base.py:301-455) on scripted stand-ins
    (fit_script.py); multiprocessing is replaced by an in-line Process/Queue."""
    import time
    import fit_script as fs

    class _Q:
        def __init__(self):
            self.items = []

        def put(self, x):
            self.items.append(x)

        def get(self):
            return self.items.pop(0)

    class _P:
        def __init__(self, target=None, args=()):
            self.target, self.args = target, args

        def start(self):
            self.target(*self.args)

        def join(self):
            pass

    class _MP:
        Queue = _Q
        Process = _P

    g = {"np": np, "time": time, "mp": _MP, "ConvergenceMonitor": lambda *a, **k: None}
    ref_loader.load_functions("base.py", ["fit_accumulate_test"], 1, g)
    Ref = type("RefDriver", (fs.ScriptedModel,), {"fit_accumulate_test": g["fit_accumulate_test"]})
    out = {}
    for name, (m_iter, thr, _) in fs.SCENARIOS.items():
        m = Ref()
        m.script(name)
        res = m.fit_accumulate_test(np.zeros((fs.N, fs.D)), fs.LEN_VEC, thr, "test", m_iter)
        params_vec, params_vec1, plist, it1, it2, cost_vec, t_labels = res
        out.update({name + "_params_vec": params_vec, name + "_params_vec1": params_vec1, name + "_plist": plist,
                    name + "_it": np.asarray([it1, it2]), name + "_cost_vec": cost_vec, name + "_t_labels": t_labels,
                    name + "_labels_local": m.labels_local, name + "_final_params_vec1": m.params_vec1,
                    name + "_finalized_with": m.finalized_with, name + "_n_iter": m.iteration})
        print("fit driver", name, "iterations", m.iteration, "best", it1, it2)
    np.savez_compressed(os.path.join(HERE, "fit_driver.npz"), **out)


def make_ou_cases():
    """Run the REFERENCE's OU tree algebra and M-step objective (phylo_hmrf.py:715-1325) on the
    shipped example tree and on a 5-leaf caterpillar."""
    import sys as _sys
    import tempfile
    from numpy.linalg import det, inv
    names = ["_initilize_tree_mtx", "_sub_tree_leaf", "_compute_base_struct", "_compute_covariance_index",
             "_search_leaf", "_search_ancestor", "_matrix1", "_ou_param_varied_constraint",
             "_ou_lik_varied_constraint", "_check_params", "_ou_lik_varied_single", "_ou_init_guess"]
    g = {"np": np, "sys": _sys, "det": det, "inv": inv, "small_eps": 1e-16}
    ref_loader.load_functions("phylo_hmrf.py", names, 1, g)
    Ref = type("RefOU", (object,), {n: g[n] for n in names})
    trees = {"example": np.loadtxt(os.path.join(ref_loader.REF, "example_input", "edge.1.txt"), dtype=int),
             "caterpillar5": np.asarray([[0, 1], [1, 2], [1, 3], [3, 4], [3, 5], [5, 6], [5, 7], [7, 8], [7, 9]])}
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)  # the reference writes base_mtx_*/ou_A*.txt into the working directory
        try:
            for tname, edges in trees.items():
                rng = np.random.default_rng(len(tname))
                m = Ref()
                m.tree_mtx, m.node_num = m._initilize_tree_mtx(edges)
                m.branch_dim = m.node_num - 1
                m.n_params = m.node_num + m.branch_dim * 2 + 1
                m.branch_vec = [None] * m.node_num
                m.base_struct = [None] * m.node_num
                m.leaf_list = m._compute_base_struct()
                m.leaf_vec = m._search_leaf()
                m.path_vec = m._search_ancestor()
                m.A1, m.A2, m.pair_list, m.parent_list = m._matrix1()
                d, K = len(m.leaf_vec), 6
                m.n_features, m.n_components, m.min_covar = d, K, 1e-3
                m.n_samples, m.lambda_0 = 5000, 1.0
                params = rng.random((K, m.n_params))
                params[1, 1:3] = 1e-9                      # beta below the 1e-7 guard
                params[2, 1:1 + m.branch_dim] = 50.0       # strong selection: near-diagonal covariance
                params[3, 1 + m.branch_dim:1 + 2 * m.branch_dim] = 1e-12   # vanishing variance: conditioning ladder
                params[3, 0] = 1e-13
                m.means_, m._covars_ = np.zeros((K, d)), np.zeros((K, d, d))
                m._ou_param_varied_constraint(params)
                A = rng.random((K, d, d))
                post = 200 + 800 * rng.random(K)
                mu_k = rng.random((K, d))
                m.stats = {'post': post, 'obs': mu_k * post[:, None],
                           'obs*obs.T': (np.einsum('kij,klj->kil', A, A) + mu_k[:, :, None] * mu_k[:, None, :])
                           * post[:, None, None]}
                m.init_ou_params = rng.random((K, m.n_params))
                liks, vals, cvs = [], [], []
                for c in range(K):
                    liks.append(m._ou_lik_varied_constraint(params[c].copy(), c))
                    vals.append(m.values.copy())
                    cvs.append(m.cv_mtx.copy())
                bad = params[0].copy(); bad[2] = 150.0
                nanp = params[0].copy(); nanp[3] = np.nan; nanp[2] = 150.0
                lik_nan = m._ou_lik_varied_constraint(nanp.copy(), 0)   # falls back to init_ou_params[0]
                obs = rng.random((300, d))
                single = [m._ou_lik_varied_single(params[c].copy(), obs) for c in (0, 2, 4)]
                single_cv = m.cv_mtx.copy()
                np.random.seed(11)
                m.initial_w2 = 0.7
                guess = m._ou_init_guess(mu_k[0])
                p = tname + "_"
                out.update({p + "edges": edges, p + "leaf_vec": m.leaf_vec, p + "A1": m.A1, p + "A2": m.A2,
                            p + "pair_list": np.asarray(m.pair_list), p + "parent": np.asarray(
                                [-1 if (isinstance(x, list) and not x) else int(x) for x in m.parent_list]),
                            p + "path_flat": np.concatenate(m.path_vec), p + "path_len": np.asarray(
                                [len(x) for x in m.path_vec]),
                            p + "leaf_rank": np.asarray([m.leaf_list[int(l)] for l in m.leaf_vec]),
                            p + "params": params, p + "means": m.means_.copy(), p + "covars": m._covars_.copy(),
                            p + "post": post, p + "obs": m.stats['obs'], p + "obsobsT": m.stats['obs*obs.T'],
                            p + "init_ou_params": m.init_ou_params, p + "liks": np.asarray(liks),
                            p + "values": np.asarray(vals), p + "cv_mtx": np.asarray(cvs),
                            p + "check": np.asarray([m._check_params(params[0]), m._check_params(bad),
                                                     m._check_params(nanp)]),
                            p + "bad": bad, p + "nanp": nanp, p + "lik_nan": lik_nan, p + "single_obs": obs,
                            p + "single": np.asarray(single), p + "single_cv": single_cv, p + "guess": guess,
                            p + "guess_mean": mu_k[0]})
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "ou_cases.npz"), **out)
    print("wrote ou_cases")


def make_cli_defaults():
    """The reference's own parse_args() (phylo_hmrf.py:1531-1568) on an empty command line."""
    import json
    from optparse import OptionParser
    g = {"OptionParser": OptionParser}
    ref_loader.load_functions("phylo_hmrf.py", ["parse_args"], 0, g)
    old = sys.argv
    sys.argv = ["phylo_hmrf.py"]
    try:
        opts = g["parse_args"]()
    finally:
        sys.argv = old
    with open(os.path.join(HERE, "cli_defaults.json"), "w") as f:
        json.dump(vars(opts), f, indent=1, sort_keys=True)
    print("wrote cli_defaults", len(vars(opts)))


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not present; fixtures can only be regenerated in the build container")
    make_case("case_diag_w", 101, [("diag", 14, 14)], d=4, K=6, beta=1.0, beta1=0.1, estimate_type=3)
    make_case("case_two_regions", 202, [("diag", 11, 11), ("rect", 7, 9)], d=5, K=8, beta=0.7, beta1=0.5,
              estimate_type=3, isolate=((1, 17),))
    make_case("case_unweighted_iso", 303, [("diag", 10, 10)], d=3, K=4, beta=2.0, beta1=0.1, estimate_type=0,
              isolate=((0, 23),))
    make_case("case_d9_k30", 404, [("diag", 16, 16)], d=9, K=30, beta=1.0, beta1=0.1, estimate_type=3)
    make_edge_cases()
    make_fit_driver_cases()
    make_ou_cases()
    make_cli_defaults()


if __name__ == "__main__":
    main()
