"""Preprocessing stage on the GPU against the oracle and the reference fixtures.

Bit-exact: rescale (no log), image scatter, 3x3 median hole fill (the wavefront order must reproduce
the reference's sequential raster scan), node order.  Tolerance: the log transform (CUDA log vs
glibc log, <= 2 ulp: rtol 1e-14) and the diffusion (float32 arithmetic like medpy, CUDA expf vs
NumPy's float32 exp: rtol 2e-5, atol 2e-6 on the float32 image)."""
import os

import numpy as np
import pytest

from oracle import prep_oracle as po

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prep_cases.npz"))
NAMES = [str(n) for n in G["names"]]


@pytest.fixture(scope="module")
def util():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from phylo_hmrf_b200 import utility
    return utility


@pytest.mark.parametrize("name", NAMES)
def test_normalise_against_reference_fixture(util, name):
    p = "prep_%s_" % name
    for tag, xmin, xmax in (("auto", -1, -1), ("fixed", 0.0, 40.0)):
        x1, vec1, lo, hi = util.normalize_feature(G[p + "value"].copy(), xmin, xmax)
        np.testing.assert_array_equal(x1, G[p + "norm_" + tag])
        np.testing.assert_array_equal(vec1, G[p + "vec1_" + tag])
        np.testing.assert_array_equal([lo, hi], G[p + "lim_" + tag])
    x, _, _, _ = util.normalize_log_feature(G[p + "value"].copy(), -1, -1)
    np.testing.assert_allclose(x, G[p + "x"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("name", NAMES)
def test_image_holefill_nodes_against_reference_fixture(util, name):
    p = "prep_%s_" % name
    data1, mtx1, pos_idx, edges = util.write_matrix_image_Ctrl_unsym1(G[p + "x"], G[p + "pos"], "", "", 8, 0, 1,
                                                                       -1, -1, -1)  # no filter
    np.testing.assert_array_equal(mtx1, G[p + "filled"])
    np.testing.assert_array_equal(data1, G[p + "data1"])
    np.testing.assert_array_equal(pos_idx, G[p + "pos_idx"])
    assert edges.shape[1] == 3


def _random_region(seed, W, d, fill):
    rng = np.random.default_rng(seed)
    ii, jj = np.triu_indices(W)
    keep = rng.random(len(ii)) < fill
    ii, jj = ii[keep], jj[keep]
    val = np.log1p((40.0 / (1.0 + (jj - ii)))[:, None] * rng.gamma(2.0, 0.5, size=(len(ii), d)))
    val[rng.random(val.shape) < 0.05] = 0.0
    return val, np.stack([ii + 11, jj + 11], axis=1).astype(np.int64)


@pytest.mark.parametrize("W,d,fill", [(5, 1, 0.9), (64, 3, 0.6), (257, 2, 0.35)])
def test_full_pipeline_diag(util, W, d, fill):
    val, pos = _random_region(W, W, d, fill)
    ref_nofilter = po.image_pipeline_diag(val, pos, filter_mode=-1)
    got = util.write_matrix_image_Ctrl_unsym1(val, pos, "", "", 8, 0, 1, -1, -1, -1)
    np.testing.assert_array_equal(got[1], ref_nofilter[1])   # hole-filled image, bit for bit
    np.testing.assert_array_equal(got[0], ref_nofilter[0])
    ref = po.image_pipeline_diag(val, pos, filter_mode=0, niter=5, kappa=50, gamma=0.1)
    got = util.write_matrix_image_Ctrl_unsym1(val, pos, "", "", 8, 0, 1, 0, 5, 50)
    np.testing.assert_allclose(got[1], ref[1], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(got[0], ref[0], rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(got[2], ref[2])
    np.testing.assert_array_equal(got[1], np.transpose(got[1], (1, 0, 2)))  # stays symmetric
    # default filter parameters (filter_param1 < 0): niter=10, kappa=50
    ref10 = po.image_pipeline_diag(val, pos, filter_mode=0, niter=10, kappa=50, gamma=0.1)
    got10 = util.write_matrix_image_Ctrl_unsym1(val, pos, "", "", 8, 0, 1, 0, -1, -1, want_image=False)
    assert got10[1] is None
    np.testing.assert_allclose(got10[0], ref10[0], rtol=4e-5, atol=4e-6)


def test_off_diagonal_block(util):
    rng = np.random.default_rng(5)
    n1, n2, d = 37, 52, 2
    ii, jj = np.meshgrid(np.arange(n1), np.arange(n2), indexing="ij")
    keep = rng.random(ii.size) < 0.55
    ii, jj = ii.ravel()[keep], jj.ravel()[keep]
    # the extremes must be present so that the window is n1 x n2
    ii = np.concatenate([ii, [0, n1 - 1]])
    jj = np.concatenate([jj, [0, n2 - 1]])
    val = rng.gamma(2.0, 0.7, size=(len(ii), d))
    pos = np.stack([ii + 100, jj + 300], axis=1).astype(np.int64)
    mtx = np.zeros((n1, n2, d))
    for k in range(len(ii)):           # utility.py:2351-2354 (later rows overwrite)
        mtx[ii[k], jj[k]] = val[k]
    # drop duplicates for the device scatter (which of two values survives is unspecified there)
    _, first = np.unique(ii * n2 + jj, return_index=True)
    last = {}
    for k in range(len(ii)):
        last[(ii[k], jj[k])] = k
    sel = np.array(sorted(last.values()))
    ref = np.stack([po.near_interpolation1a(mtx[:, :, c].copy()) for c in range(d)], axis=2)
    got = util.write_matrix_image_Ctrl_sym1(val[sel], pos[sel], "", "", 8, 0, 0, -1, -1, -1)
    np.testing.assert_array_equal(got[1], ref)
    np.testing.assert_array_equal(got[0], ref.reshape(n1 * n2, d))
    assert got[2][0].tolist() == [100, 300] and got[2][-1].tolist() == [100 + n1 - 1, 300 + n2 - 1]
    refd = np.stack([po.anisotropic_diffusion(ref[:, :, c], 5, 50, 0.1) for c in range(d)], axis=2)
    gotd = util.write_matrix_image_Ctrl_sym1(val[sel], pos[sel], "", "", 8, 0, 0, 0, 5, 50)
    np.testing.assert_allclose(gotd[1], refd, rtol=2e-5, atol=2e-6)


def test_rejects_pairs_outside_the_window_and_unbuilt_filters(util):
    val = np.ones((3, 1))
    pos = np.array([[0, 0], [1, 2], [2, 2]], dtype=np.int64)
    with pytest.raises(NotImplementedError):
        util.write_matrix_image_Ctrl_unsym1(val, pos, "", "", 8, 0, 1, 1, -1, -1)
    with pytest.raises(ValueError):
        util.normalize_feature(np.zeros((0, 3)), -1, -1)


@pytest.mark.parametrize("sigma", [0.25, 1.5, 7.0])
def test_gaussian_filter_mode_matches_scipy(util, sigma):
    """filter_mode 2 (utility.py:1584-1589): scipy.ndimage.gaussian_filter on the hole-filled image;
    sigma = 7 makes the kernel radius (28) exceed the 20-bin window, i.e. repeated reflections."""
    val, pos = _random_region(21, 20, 2, 0.7)
    ref = po.image_pipeline_diag(val, pos, filter_mode=2, sigma=sigma)
    got = util.write_matrix_image_Ctrl_unsym1(val, pos, "", "", 8, sigma, 1, 2, -1, -1)
    np.testing.assert_allclose(got[1], ref[1], rtol=1e-13, atol=1e-300)
    np.testing.assert_allclose(got[0], ref[0], rtol=1e-13, atol=1e-300)
    # sigma = 0 leaves the image unfiltered (utility.py:1585)
    none = util.write_matrix_image_Ctrl_unsym1(val, pos, "", "", 8, 0.0, 1, 2, -1, -1)
    np.testing.assert_array_equal(none[1], po.image_pipeline_diag(val, pos, filter_mode=-1)[1])


def test_region_loader_end_to_end(util):
    """utility.load_data_chromosome_sub3 (utility.py:470-534) through the GPU image pipeline, against the
    oracle composition: select -> image -> hole fill -> diffusion -> node list; len_vec entry; edge list."""
    rng = np.random.default_rng(12)
    W = 40
    ii, jj = np.triu_indices(W)
    keep = rng.random(len(ii)) < 0.7
    ii, jj = ii[keep], jj[keep]
    position = np.stack([ii + 3, jj + 3, np.arange(len(ii))], axis=1).astype(np.int64)
    x = np.log1p(rng.gamma(2.0, 3.0, size=(len(ii), 3)))
    res = 10000
    region_list = [[100000, 300000, 100000, 300000, 0, 0, 5, 0]]

    class Q:
        def put(self, item):
            self.item = item

    q = Q()
    assert util.load_data_chromosome_sub3(0, 21, region_list, x, position, [res, 8, 0, 5, 50, 0.0], q)
    rid, samples, lenvec, edges = q.item
    xs, idx = util.select_valuesPosition1_2(position, x, "", 100000, 300000, 100000, 300000, res, 0)
    ref = po.image_pipeline_diag(xs, position[idx, :2], filter_mode=0, niter=5, kappa=50, gamma=0.1)
    np.testing.assert_allclose(samples, ref[0], rtol=2e-5, atol=2e-6)
    n1 = ref[1].shape[0]
    assert rid == 0 and lenvec == [ref[0].shape[0], n1, n1, position[idx].min(), position[idx].min(), 5, 1, 21]
    from phylo_hmrf_b200 import utility
    assert edges.shape == (utility._lib.lib().phmrf_grid_edge_count(1, n1, n1, 8), 3)


def test_chromosome_loader_end_to_end(util, tmp_path):
    """loader.load_data_chromosome2 on the mini chr3 data set of tests/golden/loader_cases.npz (aligned by
    the reference itself): alignment -> rescale + log -> regions (one synteny block split at the
    centromere: two diagonal parts and their off-diagonal pair; one plain block) -> image pipeline on the
    GPU, against the oracle composition."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden_prep import write_loader_inputs
    from phylo_hmrf_b200 import loader
    L = np.load(os.path.join(os.path.dirname(__file__), "golden", "loader_cases.npz"))
    dirs = write_loader_inputs(str(tmp_path), L)
    res = int(L["resolution"])
    species = ["sp0", "sp1", "sp2"]
    samples, len_vec, edges = loader.load_data_chromosome2([3], 60.0, 0, res, 8, 0, 0.0, 0, str(tmp_path / "chrom.sizes"),
                                                           dirs, species, str(tmp_path))
    x, _, _, _ = po.normalize_feature(L["aligned_x"], 0, 60.0)
    x = po.log_transform(x)
    position = L["aligned_position"]
    ref_blocks = []
    for row in L["list1"]:
        xs, idx = util.select_valuesPosition1_2(position, x, "", row[0], row[1], row[2], row[3], res, 0)
        p = position[idx, :2]
        if row[0] == row[2] and row[1] == row[3]:
            ref_blocks.append(po.image_pipeline_diag(xs, p, filter_mode=0, niter=5, kappa=50, gamma=0.1)[0])
        else:
            s1, s2 = p[:, 0].min(), p[:, 1].min()
            n1, n2 = p[:, 0].max() - s1 + 1, p[:, 1].max() - s2 + 1
            mtx = np.zeros((n1, n2, xs.shape[1]))
            mtx[p[:, 0] - s1, p[:, 1] - s2] = xs
            for c in range(xs.shape[1]):
                plane = po.near_interpolation1a(mtx[:, :, c].copy())
                mtx[:, :, c] = po.anisotropic_diffusion(plane, 5, 50, 0.1)
            ref_blocks.append(mtx.reshape(n1 * n2, -1))
    ref = np.concatenate(ref_blocks, axis=0)
    assert samples.shape == ref.shape
    np.testing.assert_allclose(samples, ref, rtol=2e-5, atol=2e-6)
    assert [lv[8] for lv in len_vec] == [1, 0, 1, 1] and len_vec[-1][2] == len(samples) and len(edges) == 4
    for lv, blk in zip(len_vec, ref_blocks):
        assert lv[0] == len(blk) and lv[2] - lv[1] == len(blk)
