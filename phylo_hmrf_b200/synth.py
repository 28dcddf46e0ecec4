"""Synthetic contact maps of the shapes BASELINE.json names: host-side input
preparation for tests and bench.py.  NumPy only; nothing here is on the timed path.

Geometry follows the reference's data preparation: a diagonal region of B bins holds the
row-major upper triangle incl. the diagonal (utility.py:2310-2317); undirected
8-neighbourhood edges right / lower-left / lower / lower-right kept inside the triangle,
id1<id2, sorted by (id1,id2) (utility.py:1898-1931, 1960); edge distance
``|xi-xj|^2/(|xi||xj|+1e-16)``, halved between two diagonal nodes (utility.py:1919-1953).
A *band* is a contiguous range of rows plus a one-row halo on either side, which is all
phase B needs because every neighbour lies in rows x-1, x, x+1.
"""
from __future__ import annotations

import numpy as np


# hg38 autosome lengths (chr1..chr22; the public assembly constants the reference reads from
# example_input/hg38.chrom.sizes): BASELINE config 4 is one diagonal region per autosome at 50 kb
HG38_AUTOSOME_BP = (248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
                    133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
                    58617616, 64444167, 46709983, 50818468)


def autosome_bins(resolution=50000):
    """Bins per autosome with the reference's floor division (utility.py:2516)."""
    return [bp // resolution for bp in HG38_AUTOSOME_BP]


def tri_row_start(B, x):
    """Index of node (x, x) in the row-major upper triangle of a B-bin region."""
    x = np.asarray(x, dtype=np.int64)
    return x * B - (x * (x - 1)) // 2


def tri_nodes(B):
    return B * (B + 1) // 2


def band_rows(B, n_bands):
    """Row ranges balancing the node count (not the row count) across bands."""
    total = tri_nodes(B)
    starts = tri_row_start(B, np.arange(B + 1))
    cuts = [0]
    for b in range(1, n_bands):
        cuts.append(int(np.searchsorted(starts, total * b / n_bands)))
    cuts.append(B)
    return [(cuts[i], cuts[i + 1]) for i in range(n_bands)]


def window_xy(B, r0=0, r1=None):
    """Coordinates of the window nodes (rows [r0-1, r1+1) clipped) of a band, without any edge
    array: what `Model.region_grid` needs besides the features."""
    r1 = B if r1 is None else r1
    h0, h1 = max(r0 - 1, 0), min(r1 + 1, B)
    rows = np.arange(h0, h1, dtype=np.int64)
    lens = B - rows
    x = np.repeat(rows, lens)
    win_start = int(tri_row_start(B, h0))
    n_window = int(tri_row_start(B, h1)) - win_start
    y = np.arange(n_window, dtype=np.int64) - np.repeat(tri_row_start(B, rows) - win_start, lens) + x
    own_start, own_end = int(tri_row_start(B, r0)), int(tri_row_start(B, r1))
    return dict(B=B, r0=r0, r1=r1, x=x, y=y, n_window=n_window, n_own=own_end - own_start,
                own_offset=own_start - win_start, win_start=win_start)


def triangle_band(B, r0=0, r1=None):
    """Nodes and edges of rows [r0,r1) of a B-bin diagonal region with a one-row halo.

    Returns a dict: n_own, n_window, own_offset, win_start (global id of window node 0),
    x, y (int64 coordinates of the window nodes), edge_ids [E,2] window-local, sorted.
    """
    r1 = B if r1 is None else r1
    h0, h1 = max(r0 - 1, 0), min(r1 + 1, B)
    win_start = int(tri_row_start(B, h0))
    win_end = int(tri_row_start(B, h1))
    own_start, own_end = int(tri_row_start(B, r0)), int(tri_row_start(B, r1))
    rows = np.arange(h0, h1, dtype=np.int64)
    lens = B - rows
    x = np.repeat(rows, lens)
    y = np.arange(win_end - win_start, dtype=np.int64) - np.repeat(tri_row_start(B, rows) - win_start, lens) + x
    gid = np.arange(win_start, win_end, dtype=np.int64)
    # forward neighbours in ascending id order: right, lower-left, lower, lower-right
    nxt = tri_row_start(B, x + 1)
    cand = np.stack([gid + 1, nxt + (y - 1) - (x + 1), nxt + y - (x + 1), nxt + (y + 1) - (x + 1)], axis=1)
    ok = np.stack([y + 1 < B, (x + 1 <= y - 1), (x + 1 <= y) & (x + 1 < B), (y + 1 < B) & (x + 1 < B)], axis=1)
    ok[:, 1] &= x + 1 < B
    src_owned = (x >= r0) & (x < r1)
    dst_owned_next = (x + 1 >= r0) & (x + 1 < r1)
    keep = ok & np.stack([src_owned, src_owned | dst_owned_next, src_owned | dst_owned_next,
                          src_owned | dst_owned_next], axis=1)
    keep &= cand < win_end
    src = np.broadcast_to(gid[:, None], cand.shape)[keep]
    dst = cand[keep]
    e = np.stack([src - win_start, dst - win_start], axis=1)
    return dict(B=B, r0=r0, r1=r1, n_own=own_end - own_start, n_window=win_end - win_start,
                own_offset=own_start - win_start, win_start=win_start, x=x, y=y, edge_ids=e)


def features(seed, x, y, d, zero_frac=0.3, n_latent=6):
    """X [n,d] = log1p(max(0,z)): distance-decaying signal with blocky latent states, shared
    inter-species correlation ~0.6 and ~30 % zeros.  A pure function of (seed, x, y), so the
    halo of one band equals the owned rows of its neighbour."""
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    n = len(x)
    X = np.empty((n, d))
    if n == 0:
        return X
    latent_gain = 0.6 + 0.25 * np.arange(n_latent)
    species = 0.05 * np.arange(d)
    bounds = np.flatnonzero(np.diff(x)) + 1
    starts = np.concatenate([[0], bounds])
    ends = np.concatenate([bounds, [n]])
    for s, e in zip(starts, ends):
        rng = np.random.default_rng([seed, int(x[s])])
        m = e - s
        # draw for the full row suffix so a node's value does not depend on the band cut
        yy = y[s:e]
        dist = (yy - x[s]).astype(np.float64)
        lat = ((x[s] // 16) * 7 + (yy // 16) * 13) % n_latent
        amp = (2.5 * (1.0 + dist) ** -0.3 + 0.2) * latent_gain[lat]
        shared = rng.standard_normal(m)
        own = rng.standard_normal((m, d))
        z = amp[:, None] * (1.0 + 0.35 * (np.sqrt(0.6) * shared[:, None] + np.sqrt(0.4) * own)) + species[None, :]
        row = np.log1p(np.maximum(z, 0.0))
        row[rng.random((m, d)) < 0.5 * zero_frac] = 0.0
        row[rng.random(m) < 0.4 * zero_frac] = 0.0
        X[s:e] = row
    return X


def edge_distances(X, edge_ids, x, y):
    a, b = edge_ids[:, 0], edge_ids[:, 1]
    nrm = np.sqrt(np.sum(X * X, axis=1))
    diff = X[a] - X[b]
    dist = np.sum(diff * diff, axis=1) / (nrm[a] * nrm[b] + 1e-16)
    diag = x == y
    both = diag[a] & diag[b]
    return np.where(both, 0.5 * dist, dist)


def ou_covariance(rng, d, min_covar=1e-3):
    """Leaf covariance of an OU process on a caterpillar tree with d leaves and random
    branch parameters in (0,1), by the reference's recursion (phylo_hmrf.py:1056-1090):
    ``var_i = lambda_i/(2 beta_i) (1-e_i^2) + var_parent e_i^2`` with ``e_i = exp(-beta_i)``
    and ``cov(a,b) = var_mrca * exp(-sum of beta on the a<->b path below the MRCA)``; plus
    ``min_covar*I``.  Positive semi-definite by construction."""
    n_int = max(d - 1, 1)
    # internal chain I_0 <- I_1 <- ... ; branch i leads into I_i (I_0 hangs off the remote root)
    ib = rng.random(n_int) + 1e-3
    il = rng.random(n_int)
    var_int = np.empty(n_int)
    v = rng.random()  # variance at the remote root
    for i in range(n_int):
        e = np.exp(-ib[i])
        v = il[i] / (2 * ib[i]) * (1 - e * e) + v * e * e
        var_int[i] = v
    lb = rng.random(d) + 1e-3
    ll = rng.random(d)
    anc = np.minimum(np.arange(d), n_int - 1)  # leaf a hangs off I_anc[a]
    cov = np.empty((d, d))
    for a in range(d):
        ea = np.exp(-lb[a])
        cov[a, a] = ll[a] / (2 * lb[a]) * (1 - ea * ea) + var_int[anc[a]] * ea * ea
        for b in range(a + 1, d):
            m = min(anc[a], anc[b])
            path = lb[a] + lb[b] + ib[m + 1:anc[a] + 1].sum() + ib[m + 1:anc[b] + 1].sum()
            cov[a, b] = cov[b, a] = var_int[m] * np.exp(-path)
    return cov + min_covar * np.eye(d)


def model(seed, X_sample, K, d, scale=0.15):
    """K state means picked from the data (k-means++ style seeding on a subsample) and OU
    leaf covariances scaled to the data's spread."""
    rng = np.random.default_rng([seed, 7919])
    S = X_sample[rng.choice(len(X_sample), size=min(len(X_sample), 20000), replace=False)]
    means = np.empty((K, d))
    means[0] = S[rng.integers(len(S))]
    d2 = np.sum((S - means[0]) ** 2, axis=1)
    for k in range(1, K):
        p = d2 / d2.sum() if d2.sum() > 0 else None
        means[k] = S[rng.choice(len(S), p=p)]
        d2 = np.minimum(d2, np.sum((S - means[k]) ** 2, axis=1))
    covars = np.stack([scale * ou_covariance(rng, d) + 1e-3 * np.eye(d) for _ in range(K)])
    return means, covars


def potts(K, beta):
    V = np.full((K, K), float(beta))
    np.fill_diagonal(V, 0.0)
    return V


def make_band(seed, B, d, r0=0, r1=None, beta1=0.1):
    """Everything one band needs on the host: X_own, window geometry, edges and weights."""
    g = triangle_band(B, r0, r1)
    Xw = features(seed, g["x"], g["y"], d)
    dist = edge_distances(Xw, g["edge_ids"], g["x"], g["y"])
    g["edge_dist"] = dist
    g["edge_w"] = np.exp(-beta1 * dist)
    g["X_window"] = Xw
    g["X_own"] = Xw[g["own_offset"]:g["own_offset"] + g["n_own"]]
    return g
