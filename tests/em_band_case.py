"""A small one-region EM problem shared by the row-band tests (CPU/gloo with the oracle as the
per-band arithmetic, GPU/NCCL with the library): deterministic initialisation and a closed-form
M-step, so that a run on one process and a run on two can be compared number by number."""
import numpy as np

from phylo_hmrf_b200 import synth

SEED, BETA, BETA1, ET = 77, 1.0, 0.1, 3


def problem(B, D):
    g = synth.make_band(SEED, B, D, beta1=BETA1)
    N = g["n_own"]
    len_vec = [[N, 0, N, B, B, 0, 0, 0, 1, 1]]       # outputfile_description.txt:8-33: N, s1, s2, n1, n2, ..., type=1
    edge_list = np.column_stack([g["edge_ids"].astype(np.float64), g["edge_dist"]])
    return g["X_own"].copy(), len_vec, [edge_list]


def init_fn(m, X):
    means, covars = synth.model(SEED, X, m.n_components, m.n_features)
    m.means_, m._covars_ = means, covars
    m.params_vec1 = means.copy()
    m.labels = np.zeros(len(X), dtype=np.int64)
    m.labels_local = np.zeros(len(X), dtype=np.int64)


def mstep_fn(m, stats):
    post = np.maximum(stats['post'], 1e-300)
    m.means_ = stats['obs'] / post[:, None]
    m.params_vec1 = m.means_.copy()
