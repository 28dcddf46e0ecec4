"""SURVEY 8(f-4), host side: the reference's data loading re-hosted with its function names --
multi-species alignment of the Hi-C text files, synteny-block regions, and the per-chromosome /
per-region assembly of ``samples``, ``len_vec`` and ``edge_list_vec`` (utility.py:267-468,
2111-2189, 2507-2662).  File parsing and index bookkeeping are host NumPy/pandas like the reference
(one-shot I/O); the numeric stages run on the GPU through ``phylo_hmrf_b200.utility``.  The
reference forks one process per chromosome and per region (CUDA contexts do not survive a fork):
here both fan-outs are plain loops, the queue tuples are the same.

Python-2 integer semantics of the reference are spelled out (``//``)."""
from __future__ import annotations

import math
import os

import numpy as np

from . import utility


class _ListQueue:
    def __init__(self):
        self.items = []

    def put(self, item):
        self.items.append(item)


def multi_contact_matrix3A(chrom, resolution, ref_chromsize, filename_list, species, output_filename, type_id):
    """utility.py:2507-2570 + output_multi_contactMtx (:2631-2662): union-align the species' contact
    files ``<dir>/chr<chrom>.<res/1000>K.txt`` (bin start, bin start, value) on the bin-pair serial.
    Returns a pandas DataFrame with columns [0, 1, 2] (bin x, bin y, serial) + one per species."""
    import pandas as pd
    data1 = pd.read_csv(ref_chromsize, header=None, sep='\t')
    b1 = np.where(np.asarray(data1[0]) == "chr%s" % chrom)[0]
    if len(b1) == 0:
        return -1                                   # "chrom size error!" (utility.py:2521-2523)
    chrom_size = int(np.asarray(data1[1])[b1[0]])
    N = math.ceil(chrom_size // resolution)         # py2: integer division, then ceil (a no-op)
    serial1 = np.zeros(0)
    per_species = []
    for input_path in filename_list[:len(species)]:
        filename1 = "%s/chr%s.%dK.txt" % (input_path, chrom, int(resolution // 1000))
        if not os.path.exists(filename1):
            return False                            # utility.py:2534-2536
        data2 = pd.read_csv(filename1, header=None, sep='\t')
        x1 = np.asarray(data2[0]) // resolution
        x2 = np.asarray(data2[1]) // resolution
        value = np.array(data2[2], dtype=np.float64)
        value[np.isnan(value)] = -1
        serial = np.int64(N * x1 + x2)
        per_species.append((serial, x1, x2, value))
        serial1 = np.union1d(serial1, serial)
    n1 = len(serial1)
    ref = np.int64(serial1)
    mtx1 = np.zeros((n1, len(species)))
    mtx_1 = np.zeros((n1, 3))
    colnames = [0, 1, 2] + list(species)
    data_1 = pd.DataFrame(columns=colnames)
    for i, (serial, x1, x2, value) in enumerate(per_species):
        idx = np.searchsorted(ref, serial)          # mapping_Idx (utility.py:824-863): serial is a subset of ref
        if len(idx) > 0:
            mtx1[idx, i] = value
            data_1[species[i]] = mtx1[:, i]
            mtx_1[idx, 0], mtx_1[idx, 1] = x1, x2
    data_1[0], data_1[1], data_1[2] = np.int64(mtx_1[:, 0]), np.int64(mtx_1[:, 1]), np.int64(serial1)
    if output_filename != "":
        data_1.to_csv(output_filename, index=False, sep='\t')
    return data_1


def quantile_contact(chrom, resolution, ref_filename, filename_list, species):
    """utility.py:2475-2505: per species [five percentiles of the values >= 0 (at 0.05 .. 0.95 PERCENT, as
    the reference calls numpy.percentile), smallest positive value, maximum, maximum / 0.95-th percentile,
    number of positive values, number of non-negative values]; NaN counts as -1."""
    import pandas as pd
    data1 = pd.read_csv(ref_filename, header=None, sep='\t')
    if not (np.asarray(data1[0]) == "chr%s" % chrom).any():
        raise IOError("chromosome size of chr%s not found" % chrom)      # utility.py:2584-2586 returns -1
    m_vec = np.zeros((len(species), 10))
    for i, input_path in enumerate(filename_list[:len(species)]):
        filename1 = "%s/chr%s.%dK.txt" % (input_path, chrom, int(resolution // 1000))
        if not os.path.exists(filename1):
            raise IOError("File %s does not exist" % filename1)
        values = np.array(pd.read_csv(filename1, header=None, sep='\t')[2], dtype=np.float64)
        values[np.isnan(values)] = -1
        b1, b2 = np.where(values > 0)[0], np.where(values >= 0)[0]
        m_vec[i, 0:5] = np.percentile(values[b2], [0.05, 0.25, 0.50, 0.75, 0.95])
        m_vec[i, 5] = np.min(values[b1])
        m_vec[i, 6] = np.max(values)
        m_vec[i, 7] = np.max(values) / (m_vec[i, 4] + 1e-16)
        m_vec[i, 8], m_vec[i, 9] = len(b1), len(b2)
    return m_vec


def quantile_contact_vec(chrom_vec, resolution, ref_filename, filename_list, species):
    """utility.py:2463-2473: the rows of quantile_contact over the chromosomes."""
    m = [quantile_contact(c, resolution, ref_filename, filename_list, species) for c in chrom_vec]
    return np.concatenate(m, axis=0)


def subregion1(filename, chrom_id, resolution, region_points, type_id):
    """utility.py:2111-2189: synteny blocks (start, stop, length per line) -> (blocks, regions).  A block
    that spans a listed centromere (by more than two bins either side) is cut into the part before and the
    part after it, both keeping the block's id; every block id then yields its diagonal region, or -- when
    it was cut -- every ordered pair (i <= j) of its parts.  A region row is
    [start, stop, start_a, stop_a, length, length_a, block id, running region number, chrom_id]."""
    table = np.atleast_2d(np.loadtxt(filename, dtype='int', delimiter='\t'))
    blocks = [[int(r[0]), int(r[1]), int(r[2]), k] for k, r in enumerate(table)]
    margin = 2 * resolution
    for centro_start, centro_stop in ((p[0], p[1]) for p in region_points):
        hit = next((k for k, b in enumerate(blocks) if b[0] < centro_start - margin and b[1] > centro_stop + margin),
                   None)
        if hit is None:
            continue
        start, stop, _, block_id = blocks[hit]
        blocks[hit:hit + 1] = [[start, centro_start, centro_start - start, block_id],
                               [centro_stop, stop, stop - centro_stop, block_id]]
    regions = []
    for block_id in sorted({b[3] for b in blocks}):
        parts = [b for b in blocks if b[3] == block_id]
        for i, first in enumerate(parts):
            for second in parts[i:]:
                regions.append([first[0], first[1], second[0], second[1], first[2], second[2], block_id,
                                len(regions), chrom_id])
    return blocks, regions


# centromere positions of chr3 and chr6 in hg38 (utility.py:383)
_REGION_POINTS_HG38 = np.asarray([[3, 90279522, 93797661], [6, 57542947, 61520508]])


def load_data_chromosome_sub1_2(chrom_id, x_max, x_min, resolution, num_neighbor, filter_mode, sigma, diagonal_typeId,
                                ref_filename, filename_list, species, data_path, m_queue, device=0):
    """utility.py:335-468: one chromosome -> (chrom_id, samples, len_vec, edge_list_vec) on ``m_queue``."""
    chrom = str(chrom_id)
    data_ori = multi_contact_matrix3A(chrom, resolution, ref_filename, filename_list, species, "", 0)
    if data_ori is False or isinstance(data_ori, int):
        raise IOError("could not load the contact files of chr%s" % chrom)
    colnames = list(data_ori)
    position = np.asarray(data_ori.loc[:, colnames[0:3]])
    x1 = np.ascontiguousarray(np.asarray(data_ori.loc[:, colnames[3:]], dtype=np.float64))
    x, _, x_min, x_max = utility.normalize_log_feature(x1, x_min, x_max, device=device)   # :358-362
    region_points = [_REGION_POINTS_HG38[i, 1:] for i in np.where(_REGION_POINTS_HG38[:, 0] == chrom_id)[0]]
    filename3 = "%s/chr%s.synteny.txt" % (data_path, chrom)
    _, region_list2_ori = subregion1(filename3, chrom_id, resolution, region_points, 0)
    if diagonal_typeId == 1:   # diagonal blocks only (utility.py:398-402)
        region_list2 = [r for r in region_list2_ori if r[0] == r[2] and r[1] == r[3]]
    else:
        region_list2 = region_list2_ori
    filter_param1, filter_param2 = (5, 50) if filter_mode == 0 else (-1, -1)
    param_vec = [resolution, num_neighbor, filter_mode, filter_param1, filter_param2, sigma]
    q = _ListQueue()
    for region_id in range(len(region_list2)):
        utility.load_data_chromosome_sub3(region_id, chrom_id, region_list2, x, position, param_vec, q, device=device)
    results = sorted(q.items, key=lambda v: v[0])
    samples, len_vec, edge_list_vec = [], [], []
    id_1 = 0
    for vec1 in results:
        t_samples, t_lenvec, t_edgelist = vec1[1], list(vec1[2]), vec1[3]
        id_2 = id_1 + t_samples.shape[0]
        t_lenvec.insert(1, id_2)
        t_lenvec.insert(1, id_1)
        id_1 = id_2
        samples.append(t_samples)
        len_vec.append(t_lenvec)
        edge_list_vec.append(t_edgelist)
    samples = np.concatenate(samples, axis=0) if samples else np.zeros((0, len(species)))
    m_queue.put((chrom_id, samples, len_vec, edge_list_vec))
    return True


def load_data_chromosome2(chrom_vec, x_max, x_min, resolution, num_neighbor, filter_mode, sigma, diagonal_typeId,
                          ref_filename, filename_list, species, data_path, annotation="", device=0):
    """utility.py:267-333: (samples, len_vec, edge_list_vec) over the chromosomes, ordered by chromosome id."""
    q = _ListQueue()
    for chrom_id in chrom_vec:
        load_data_chromosome_sub1_2(chrom_id, x_max, x_min, resolution, num_neighbor, filter_mode, sigma,
                                    diagonal_typeId, ref_filename, filename_list, species, data_path, q, device=device)
    results = sorted(q.items, key=lambda v: v[0])
    samples, len_vec, edge_list_vec = [], [], []
    n_acc = 0
    for vec1 in results:
        t_samples, t_lenvec, t_edges = vec1[1], vec1[2], vec1[3]
        samples.append(t_samples)
        for temp1 in t_lenvec:
            temp1[1] += n_acc
            temp1[2] += n_acc
            len_vec.append(temp1)
        n_acc += t_samples.shape[0]
        edge_list_vec.extend(t_edges)
    return np.concatenate(samples, axis=0), len_vec, edge_list_vec
