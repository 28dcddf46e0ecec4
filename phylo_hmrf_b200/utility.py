"""the reference's region edge-list builders (utility.py) re-hosted on the GPU
library, with their original names and signatures.  Only DENSE regions are supported -- which
is what the reference's own pipeline produces (write_matrix_array_v1 enumerates every cell of
the upper triangle / rectangle, utility.py:2300-2317, 2375-2382)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f64, check, dptr


def _dense(serial, expect):
    serial = np.asarray(serial)
    if len(serial) != len(expect) or not np.array_equal(np.int64(serial), expect):
        raise NotImplementedError("only dense regions (every cell present, row-major order) are supported")


def _edges(data1, kind, n1, n2, num_neighbor, device):
    X = as_f64(data1)
    E = int(_lib.lib().phmrf_grid_edge_count(kind, n1, n2, int(num_neighbor)))
    if E < 0:
        raise ValueError("num_neighbor must be 8 or 4 (the reference's only runnable values, utility.py:1898-1916)")
    out = np.empty((E, 3), dtype=np.float64)
    check(_lib.lib().phmrf_grid_edges(int(device), dptr(X), X.shape[1], kind, n1, n2, int(num_neighbor), dptr(out), E))
    return out


def edge_weightlist_grid3_undirected_unsym(data1, serial, window_size, output_filename, num_neighbor, device=0):
    """utility.py:1871-1973 (diagonal region).  Returns edge_list [E,3] = (id1, id2, d_ij)."""
    N = int(window_size)
    _dense(serial, np.asarray([i * N + j for i in range(N) for j in range(i, N)], dtype=np.int64)
           if N < 2048 else _tri_serial(N))
    edge_list = _edges(data1, 1, N, N, num_neighbor, device)
    _maybe_write(edge_list, output_filename)
    return edge_list


def edge_weightlist_grid3_undirected(data1, serial, window_size, output_filename, num_neighbor, device=0):
    """utility.py:1975-2053 (rectangular, off-diagonal region)."""
    N1, N2 = int(window_size[0]), int(window_size[1])
    _dense(serial, np.arange(N1 * N2, dtype=np.int64))
    edge_list = _edges(data1, 0, N1, N2, num_neighbor, device)
    _maybe_write(edge_list, output_filename)
    return edge_list


def _tri_serial(N):
    xs, ys = np.triu_indices(N)
    return xs.astype(np.int64) * N + ys


def _maybe_write(edge_list, output_filename):
    if output_filename != '':  # same tab-separated file as utility.py:1962-1971
        import pandas as pd
        data2 = pd.DataFrame({'id1': np.int64(edge_list[:, 0]), 'id2': np.int64(edge_list[:, 1]),
                              'weight': edge_list[:, 2]})
        data2.to_csv(output_filename, index=False, header=False, sep='\t')
