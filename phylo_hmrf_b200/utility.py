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


# ---- per-region preprocessing (image pipeline) ---------------------------------

def normalize_feature(x1, x_min, x_max, device=0):
    """utility.py:867-897.  Returns (x1, vec1, x_min, x_max); like the reference, ``x1`` is
    transformed in place when it is a C-contiguous float64 array."""
    return _normalise(x1, x_min, x_max, 0, device)


def normalize_log_feature(x1, x_min, x_max, device=0):
    """normalize_feature followed by ``x = np.log(1 + x1)`` (utility.py:358-362) in one pass."""
    return _normalise(x1, x_min, x_max, 1, device)


def _normalise(x1, x_min, x_max, log1p, device):
    x = x1 if (isinstance(x1, np.ndarray) and x1.dtype == np.float64 and x1.flags.c_contiguous) else as_f64(x1)
    if x.ndim != 2 or x.size == 0:
        raise ValueError("x1 must be a non-empty [n_samples, n_species] array")
    lo, hi = C.c_double(float(x_min)), C.c_double(float(x_max))
    vec1 = np.empty((x.shape[1], 2), dtype=np.float64)
    check(_lib.lib().phmrf_prep_normalise(int(device), dptr(x), x.shape[0], x.shape[1], C.byref(lo), C.byref(hi),
                                          dptr(vec1), int(log1p)))
    return x, vec1, lo.value, hi.value


def _region_image(value, pos, kind, filter_mode, filter_param1, filter_param2, want_image, device, sigma=0.0):
    value = as_f64(value)
    pos = np.asarray(pos)
    if value.ndim != 2 or pos.ndim != 2 or pos.shape[0] != value.shape[0] or pos.shape[1] < 2 or value.shape[0] == 0:
        raise ValueError("value must be [n, d] and pos [n, >=2] (bin pair in the first two columns)")
    pos = np.ascontiguousarray(pos[:, :2], dtype=np.int64)  # utility.py:2204-2222 reads columns 0 and 1 only
    d = value.shape[1]
    s1, s2 = int(pos[:, 0].min()), int(pos[:, 1].min())
    e1, e2 = int(pos[:, 0].max()), int(pos[:, 1].max())
    if kind == 1:  # utility.py:2204-2211: one square window over both coordinates
        s1 = s2 = min(s1, s2)
        n1 = n2 = max(e1, e2) - s1 + 1
        n_nodes = n1 * (n1 + 1) // 2
    else:          # utility.py:2341-2346
        n1, n2 = e1 - s1 + 1, e2 - s2 + 1
        n_nodes = n1 * n2
    niter, kappa, mode = 0, 50.0, -1
    if filter_mode == 0:    # anisotropic diffusion (utility.py:1566-1573)
        niter, kappa = (10, 50.0) if filter_param1 < 0 else (int(filter_param1), float(filter_param2))
        mode = 0
    elif filter_mode == 1:  # skimage denoise_bilateral (utility.py:1575-1582)
        raise NotImplementedError("filter_mode 1 (bilateral filter) is not built")
    elif sigma > 0:         # any other mode: Gaussian blur when sigma > 0 (utility.py:1584-1589)
        mode = 2
    data1 = np.empty((n_nodes, d), dtype=np.float64)
    mtx1 = np.empty((n1, n2, d), dtype=np.float64) if want_image else None
    check(_lib.lib().phmrf_prep_region_image(int(device), dptr(value), pos.ctypes.data_as(C.POINTER(C.c_int64)),
                                             value.shape[0], d, kind, s1, s2, n1, n2, mode,
                                             niter, kappa, 0.1, float(sigma), dptr(data1),
                                             dptr(mtx1) if want_image else None))
    return data1, mtx1, (s1, s2), (n1, n2)


def write_matrix_image_Ctrl_unsym1(value, pos, output_filename1, output_filename2, num_neighbor, sigma, type_id,
                                   filter_mode, filter_param1, filter_param2, device=0, want_image=True):
    """utility.py:1519-1598 (diagonal region): image, hole fill, anisotropic diffusion, upper-triangle
    node list and its edge list.  Returns (data1, mtx1, pos_idx, edge_list)."""
    data1, mtx1, (s1, _), (n1, _) = _region_image(value, pos, 1, filter_mode, filter_param1, filter_param2,
                                                  want_image, device, sigma)
    ii, jj = np.triu_indices(n1)
    pos_idx = np.stack([ii, jj], axis=1) + s1
    edge_list = _edges(data1, 1, n1, n1, num_neighbor, device)
    _maybe_write(edge_list, output_filename2)
    return data1, mtx1, pos_idx, edge_list


def write_matrix_image_Ctrl_sym1(value, pos, output_filename1, output_filename2, num_neighbor, sigma, type_id,
                                 filter_mode, filter_param1, filter_param2, device=0, want_image=True):
    """utility.py:1704-1783 (off-diagonal block).  Returns (data1, mtx1, pos_idx, edge_list)."""
    data1, mtx1, (s1, s2), (n1, n2) = _region_image(value, pos, 0, filter_mode, filter_param1, filter_param2,
                                                    want_image, device, sigma)
    ii, jj = np.meshgrid(np.arange(n1), np.arange(n2), indexing="ij")
    pos_idx = np.stack([ii.ravel() + s1, jj.ravel() + s2], axis=1)
    edge_list = _edges(data1, 0, n1, n2, num_neighbor, device)
    _maybe_write(edge_list, output_filename2)
    return data1, mtx1, pos_idx, edge_list


def select_valuesPosition1_2(position, x, output_filename, position1, position2, position1a, position2a, resolution,
                             border_type=0):
    """utility.py:1331-1364: rows of the aligned matrix whose bin pair lies in the region
    [position1, position2] x [position1a, position2a] (genomic coordinates; a bin pair is placed at the
    start of its first bin and, for border types 0 and 1, the END of its second bin).  Host NumPy.
    Returns (x[rows], rows)."""
    position = np.asarray(position)
    first = position[:, 0] * resolution
    if border_type == 0:
        second = (position[:, 1] + 1) * resolution
        inside = (first >= position1) & (first <= position2) & (second >= position1a) & (second <= position2a)
    elif border_type == 1:
        inside = (first >= position1) & ((position[:, 1] + 1) * resolution <= position2)
    else:
        second = position[:, 1] * resolution
        inside = (first >= position1) & (first < position2) & (second >= position1a) & (second < position2a)
    rows = np.flatnonzero(inside)
    if output_filename != "":   # the reference's tab-separated dump: three position columns, then the species
        import pandas as pd
        table = pd.DataFrame(np.hstack([position[rows, :3], np.asarray(x)[rows]]))
        table[[0, 1, 2]] = position[rows, :3]
        table.to_csv(output_filename, index=False, sep='\t')
    return x[rows, :], rows


def load_data_chromosome_sub3(region_id, chrom_id, region_list, x, position, param_vec, m_queue, device=0):
    """utility.py:470-534: one region of a chromosome -> (region_id, samples, len_vec entry, edge list) on
    ``m_queue`` (anything with ``put``).  The image pipeline and the edge list run on the GPU."""
    t_position1 = region_list[region_id]
    position1, position2, position1a, position2a = t_position1[0], t_position1[1], t_position1[2], t_position1[3]
    region_id1 = t_position1[6]
    resolution, num_neighbor, filter_mode, filter_param1, filter_param2, sigma = param_vec[:6]
    type_id1 = 1 if (position1 == position1a and position2 == position2a) else 0
    x1, idx = select_valuesPosition1_2(position, x, "", position1, position2, position1a, position2a, resolution, 0)
    t_position = np.asarray(position)[idx, :]
    if type_id1 == 1:
        x1, _, _, edge_list_1 = write_matrix_image_Ctrl_unsym1(x1, t_position, "", "", num_neighbor, sigma, 1,
                                                               filter_mode, filter_param1, filter_param2,
                                                               device=device, want_image=False)
        start_region1 = start_region2 = np.min(t_position)          # utility.py:516-517
        lo = int(min(t_position[:, 0].min(), t_position[:, 1].min()))
        n1 = n2 = int(max(t_position[:, 0].max(), t_position[:, 1].max())) - lo + 1
    else:
        x1, _, _, edge_list_1 = write_matrix_image_Ctrl_sym1(x1, t_position, "", "", num_neighbor, sigma, 0,
                                                             filter_mode, filter_param1, filter_param2,
                                                             device=device, want_image=False)
        temp1 = np.min(t_position, 0)
        start_region1, start_region2 = temp1[0], temp1[1]
        n1 = int(t_position[:, 0].max() - t_position[:, 0].min()) + 1
        n2 = int(t_position[:, 1].max() - t_position[:, 1].min()) + 1
    t_lenvec = [x1.shape[0], n1, n2, start_region1, start_region2, region_id1, type_id1, chrom_id]
    m_queue.put((region_id, x1, t_lenvec, edge_list_1))
    return True
