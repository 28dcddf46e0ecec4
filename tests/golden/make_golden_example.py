"""Fixture from the reference's SHIPPED example data (BASELINE configs 1-2): chr22 of the three species
present under example_input/test_data (gorGor4, panTro5, panPan2; hg38 is one of the missing blobs) plus a
deterministic stand-in for the fourth leaf, aligned by `loader.multi_contact_matrix3A`, a 240-bin window at
the start of the chr22 synteny block (example_input/chr22.synteny.txt), the example tree
(example_input/edge.1.txt: 4 leaves) for the state covariances, K = 10 and K = 20.

The `ref_*` arrays come from the reference's own method bodies (ref_loader.py) exactly as in make_golden.py;
only every 13th row of the [N,K] arrays is kept so that the fixture stays small.

Run in the build container only:  ``python tests/golden/make_golden_example.py``
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
import make_golden as mg  # noqa: E402
from oracle import phmrf_oracle as orc  # noqa: E402
from phylo_hmrf_b200 import loader, ou  # noqa: E402

WINDOW, STRIDE, RES = 240, 13, 50000


def example_features():
    root = os.path.join(ref_loader.REF, "example_input")
    species = ["gorGor4", "panTro5", "panPan2"]
    dirs = [os.path.join(root, "test_data", "hic_" + s) for s in species]
    df = loader.multi_contact_matrix3A("22", RES, os.path.join(root, "hg38.chrom.sizes"), dirs, species, "", 0)
    with open(os.path.join(root, "chr22.synteny.txt")) as f:
        start_bp = int(f.readline().split()[0])
    b0 = start_bp // RES
    p1, p2 = np.asarray(df[0]), np.asarray(df[1])
    keep = (p1 >= b0) & (p2 < b0 + WINDOW)
    img = np.zeros((3, WINDOW, WINDOW))
    for s, name in enumerate(species):
        v = np.maximum(np.asarray(df[name])[keep], 0.0)        # -1 marks "absent in this species"
        img[s, p1[keep] - b0, p2[keep] - b0] = v
    iu = np.triu_indices(WINDOW)                                # row-major upper triangle (utility.py:2310-2317)
    raw = np.stack([img[s][iu] for s in range(3)], axis=1)
    # stand-in for the missing hg38 track: the species median, modulated deterministically
    fourth = np.median(raw, axis=1) * (1.0 + 0.15 * np.sin(0.37 * iu[0] + 0.11 * iu[1]))
    X = np.log1p(np.column_stack([raw, fourth]) / 50.0)
    serial = iu[0] * WINDOW + iu[1]
    return X, serial


def model_from_tree(rng, X, K):
    tree_edges = [[int(v) for v in line.split()] for line in
                  open(os.path.join(ref_loader.REF, "example_input", "edge.1.txt")) if line.strip()]
    tree = ou.OUTree(tree_edges)
    assert tree.n_leaves == X.shape[1] == 4
    picks = X[rng.choice(len(X), size=K, replace=False)]
    covars = np.empty((K, 4, 4))
    for k in range(K):
        _, _, cov = tree.moments(rng.random(tree.n_params))
        covars[k] = 0.2 * cov + 1e-3 * np.eye(4)
    return picks + 0.01 * rng.standard_normal(picks.shape), covars


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not present")
    X, serial = example_features()
    util = ref_loader.load_utility(["mapping_Idx", "_sort_array", "edge_weightlist_grid3_undirected_unsym"])
    el = np.asarray(util["edge_weightlist_grid3_undirected_unsym"](X, serial, WINDOW, '', 8), dtype=np.float64)
    N = len(X)
    out = dict(X=X, window=WINDOW, stride=STRIDE, n_edges=len(el), edge_dist_sum=el[:, 2].sum(),
               edge_head=el[:200], beta=1.0, beta1=0.1, estimate_type=3)
    for K in (10, 20):
        rng = np.random.default_rng(2200 + K)
        means, covars = model_from_tree(rng, X, K)
        rec = mg._Recorder()
        cls = ref_loader.build_reference_class({
            "pygco": rec,
            "log_multivariate_normal_density": lambda X_, m_, c_, t_: orc.log_multivariate_normal_density_full(X_, m_, c_),
        })
        m = object.__new__(cls)
        m.n_components, m.n_features = K, 4
        m.beta, m.beta1, m.estimate_type, m.covariance_type = 1.0, 0.1, 3, 'full'
        m.means_, m._covars_ = means, covars
        len_vec = [[N, 0, N, WINDOW, WINDOW, 0, 0, 0, 1, 22]]
        m.len_vec = len_vec
        m.edge_potential = m._pairwise_potential()
        (m.edge_weightList_undirected_vec, m.edge_idList_undirected_vec,
         m.neighbor_edgeIdx_vec) = m._edge_weight_undirected_vec(X, len_vec, [el])
        m.labels_local = rng.integers(0, K, size=N).astype(np.int64)
        m.labels = m.labels_local.copy()
        flip = np.random.default_rng(2300 + K)

        def choose(unary):
            lab = np.argmin(unary, axis=1)
            f = flip.random(len(lab)) < 0.1
            lab[f] = flip.integers(0, K, size=int(f.sum()))
            return lab.astype(np.int64)

        rec.next_labels = choose
        q = mg._Queue()
        m._predict_posteriors(X, len_vec, 0, q)
        _, stats, labels, c_pair, c_pair_norm, c_unary, c_total = q.items[-1]
        call = rec.calls[-1]
        logprob = -call["unary_cost"]
        post = m._compute_posteriors_graph(X, labels, logprob, 0)[0]
        p = "k%d_" % K
        out.update({p + "means": means, p + "covars": covars, p + "labels": np.asarray(labels, dtype=np.int16),
                    p + "ref_logprob_rows": logprob[::STRIDE], p + "ref_post_rows": post[::STRIDE],
                    p + "ref_costs": np.asarray([c_pair, c_pair_norm, c_unary, c_total]),
                    p + "ref_stats_post": stats["post"], p + "ref_stats_obs": stats["obs"],
                    p + "ref_stats_obsobsT": stats["obs*obs.T"],
                    p + "ref_absmax_unary": np.abs(call["unary_cost"]).max(),
                    p + "ref_edge_w_sum": m.edge_weightList_undirected_vec[0].sum()})
        print("K =", K, "N =", N, "E =", len(el), "costs", c_pair, c_pair_norm, c_unary, c_total)
    np.savez_compressed(os.path.join(HERE, "example_chr22.npz"), **out)
    print("wrote example_chr22.npz", os.path.getsize(os.path.join(HERE, "example_chr22.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
