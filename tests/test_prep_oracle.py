"""The preprocessing oracle (oracle/prep_oracle.py) against fixtures produced by the reference's own
functions (tests/golden/make_golden_prep.py): bit-identical, both are NumPy float64 in the same
operation order."""
import os

import numpy as np
import pytest

from oracle import prep_oracle as po

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "prep_cases.npz"))
NAMES = [str(n) for n in G["names"]]


@pytest.mark.parametrize("name", NAMES)
def test_normalize_feature(name):
    p = "prep_%s_" % name
    for tag, xmin, xmax in (("auto", -1, -1), ("fixed", 0.0, 40.0)):
        x1, vec1, lo, hi = po.normalize_feature(G[p + "value"], xmin, xmax)
        np.testing.assert_array_equal(x1, G[p + "norm_" + tag])
        np.testing.assert_array_equal(vec1, G[p + "vec1_" + tag])
        np.testing.assert_array_equal([lo, hi], G[p + "lim_" + tag])
    np.testing.assert_array_equal(po.log_transform(G[p + "norm_auto"]), G[p + "x"])


@pytest.mark.parametrize("name", NAMES)
def test_image_holefill_flatten(name):
    p = "prep_%s_" % name
    mtx, start = po.matrix_image(G[p + "x"], G[p + "pos"])
    np.testing.assert_array_equal(mtx, G[p + "mtx"])
    assert start == int(G[p + "start"][0])
    filled = np.stack([po.near_interpolation1(mtx[:, :, c].copy()) for c in range(mtx.shape[2])], axis=2)
    np.testing.assert_array_equal(filled, G[p + "filled"])
    assert (filled != mtx).any(), "fixture must exercise the hole fill"
    np.testing.assert_array_equal(po.near_interpolation1a(G[p + "rect"].copy()), G[p + "rect_filled"])
    data1, pos_idx, serial = po.matrix_array(filled, start, 1)
    np.testing.assert_array_equal(data1, G[p + "data1"])
    np.testing.assert_array_equal(pos_idx, G[p + "pos_idx"])
    np.testing.assert_array_equal(serial, G[p + "serial"])


def test_diffusion_restatement_properties():
    """medpy is absent (parity unpinned): check what the published scheme guarantees -- a constant
    image is a fixed point, the total is conserved (zero-flux borders), symmetric in -> symmetric
    out, and the result is float32."""
    rng = np.random.default_rng(0)
    a = rng.random((17, 17)) * 5
    a = a + a.T
    out = po.anisotropic_diffusion(a, niter=5, kappa=50, gamma=0.1)
    assert out.dtype == np.float32
    np.testing.assert_array_equal(out, out.T)
    assert abs(out.astype(np.float64).sum() - a.astype(np.float32).astype(np.float64).sum()) < 1e-2
    c = np.full((9, 11), 3.25)
    np.testing.assert_array_equal(po.anisotropic_diffusion(c, niter=3), c.astype(np.float32))
    # smoothing: the variance does not grow
    assert out.var() <= a.var()


def _wavefront_fill(m, symmetric):
    """The schedule csrc/kernels_prep.cu (holefill_kernel) uses, restated with NumPy: all cells with
    the same t = 2i + j are updated together from a snapshot; cells below the diagonal are read
    through their mirror and rewritten at the end."""
    m = m.copy()
    n1, n2 = m.shape
    i_hi, j_hi = n1 - 2, n2 - 2
    if i_hi < 2 or j_hi < 2:
        return m
    for t in range(2 * 2 + 2, 2 * i_hi + j_hi + 1):
        snap = m.copy()
        for i in range(2, i_hi + 1):
            j = t - 2 * i
            if j < (i if symmetric else 2) or j > j_hi:
                continue
            if snap[i, j] < po.THRESH1:
                w = []
                for di in (-1, 0, 1):
                    for dj in (-1, 0, 1):
                        if di == 0 and dj == 0:
                            continue
                        a, b = i + di, j + dj
                        if symmetric and a > b:
                            a, b = b, a
                        w.append(snap[a, b])
                med = np.median(w)
                if med > po.THRESH1:
                    m[i, j] = med
    if symmetric:
        iu = np.triu_indices(n1, 1)
        m[(iu[1], iu[0])] = m[iu]
    return m


@pytest.mark.parametrize("name", NAMES)
def test_wavefront_schedule_equals_the_sequential_scan(name):
    """The GPU hole fill's 2i+j wavefront must give the reference's in-place raster-scan result."""
    p = "prep_%s_" % name
    mtx = G[p + "mtx"]
    for c in range(mtx.shape[2]):
        np.testing.assert_array_equal(_wavefront_fill(mtx[:, :, c], True), G[p + "filled"][:, :, c])
    np.testing.assert_array_equal(_wavefront_fill(G[p + "rect"], False), G[p + "rect_filled"])
    rng = np.random.default_rng(7)   # sparse image: long fill cascades along rows and across the diagonal
    s = rng.random((31, 31)) * (rng.random((31, 31)) < 0.45)
    s = np.triu(s) + np.triu(s, 1).T
    np.testing.assert_array_equal(_wavefront_fill(s, True), po.near_interpolation1(s.copy()))
    r = rng.random((19, 27)) * (rng.random((19, 27)) < 0.5)
    np.testing.assert_array_equal(_wavefront_fill(r, False), po.near_interpolation1a(r.copy()))


def _tiled_fill(m, symmetric, R=4, C=6, reverse=False):
    """The schedule of holefill_tile_kernel restated with NumPy: the skewed plane (i, s = i + j) cut into
    R x C tiles, tile anti-diagonals I + S in turn, the tiles of one anti-diagonal in ANY order (they must be
    independent: `reverse` flips it), the cells of a tile by their own anti-diagonals ri + si from a snapshot."""
    m = m.copy()
    n1, n2 = m.shape
    i_hi, j_hi = n1 - 2, n2 - 2
    NI, NS = -(-n1 // R), -(-(n1 + n2 - 1) // C)
    for w in range(NI + NS - 1):
        tiles = [(I, w - I) for I in range(max(0, w - NS + 1), min(NI - 1, w) + 1)]
        for I, S in (reversed(tiles) if reverse else tiles):
            for d in range(R + C - 1):
                snap = m.copy()
                for ri in range(R):
                    si = d - ri
                    i = I * R + ri
                    if not (0 <= si < C) or i < 2 or i > i_hi:
                        continue
                    j = S * C + si - i
                    if j < (i if symmetric else 2) or j > j_hi or snap[i, j] >= po.THRESH1:
                        continue
                    w8 = []
                    for di in (-1, 0, 1):
                        for dj in (-1, 0, 1):
                            if di or dj:
                                a, b = i + di, j + dj
                                if symmetric and a > b:
                                    a, b = b, a
                                w8.append(snap[a, b])
                    med = np.median(w8)
                    if med > po.THRESH1:
                        m[i, j] = med
    if symmetric:
        iu = np.triu_indices(n1, 1)
        m[(iu[1], iu[0])] = m[iu]
    return m


@pytest.mark.parametrize("reverse", [False, True])
def test_tiled_skewed_schedule_equals_the_sequential_scan(reverse):
    """holefill_tile_kernel's order (tiles of the skewed plane, anti-diagonal by anti-diagonal) must give the
    reference's in-place raster-scan result, whichever way the tiles of one anti-diagonal are ordered."""
    rng = np.random.default_rng(11)
    for n, dens in ((31, 0.45), (23, 0.3), (40, 0.6)):   # sparse images: long fill cascades, also across the diagonal
        s = rng.random((n, n)) * (rng.random((n, n)) < dens)
        s = np.triu(s) + np.triu(s, 1).T
        np.testing.assert_array_equal(_tiled_fill(s, True, reverse=reverse), po.near_interpolation1(s.copy()))
    for shape, dens in (((19, 27), 0.5), ((33, 14), 0.35)):
        r = rng.random(shape) * (rng.random(shape) < dens)
        np.testing.assert_array_equal(_tiled_fill(r, False, reverse=reverse), po.near_interpolation1a(r.copy()))
    for name in NAMES:                                    # and the reference's own fixtures
        p = "prep_%s_" % name
        np.testing.assert_array_equal(_tiled_fill(G[p + "rect"], False, R=3, C=5, reverse=reverse), G[p + "rect_filled"])


def test_host_side_argument_checks_need_no_gpu():
    """The utility.py mirrors validate their arguments before any device call."""
    from phylo_hmrf_b200 import utility
    val = np.ones((3, 2))
    pos = np.array([[0, 0], [0, 1], [1, 1]], dtype=np.int64)
    with pytest.raises(NotImplementedError):   # bilateral filter (skimage) is not built
        utility.write_matrix_image_Ctrl_unsym1(val, pos, "", "", 8, 0.0, 1, 1, -1, -1)
    with pytest.raises(ValueError):
        utility.write_matrix_image_Ctrl_unsym1(val, pos[:2], "", "", 8, 0.0, 1, 0, 5, 50)
    with pytest.raises(ValueError):
        utility.normalize_feature(np.zeros((0, 3)), -1, -1)


def test_select_values_position_against_reference_fixture():
    from phylo_hmrf_b200 import utility
    S = np.load(os.path.join(os.path.dirname(__file__), "golden", "select_cases.npz"))
    for k in range(4):
        bt, p1, p2, p1a, p2a, res = (int(v) for v in S["sel%d_args" % k])
        xs, idx = utility.select_valuesPosition1_2(S["position"], S["x"], "", p1, p2, p1a, p2a, res, bt)
        np.testing.assert_array_equal(idx, S["sel%d_idx" % k])
        np.testing.assert_array_equal(xs, S["sel%d_x" % k])


def test_region_loader_glue(monkeypatch):
    """utility.load_data_chromosome_sub3 (utility.py:470-534) with the two GPU-backed image builders
    replaced by oracle stand-ins: region typing, window shape, start bins and the queue tuple."""
    from phylo_hmrf_b200 import utility

    def fake_unsym(value, pos, f1, f2, nn, sigma, type_id, fm, fp1, fp2, device=0, want_image=True):
        data1, mtx1, pos_idx, _ = po.image_pipeline_diag(value, pos[:, :2], filter_mode=-1)
        return data1, None, pos_idx, np.zeros((1, 3))

    def fake_sym(value, pos, f1, f2, nn, sigma, type_id, fm, fp1, fp2, device=0, want_image=True):
        return value.copy(), None, pos[:, :2].copy(), np.zeros((2, 3))

    monkeypatch.setattr(utility, "write_matrix_image_Ctrl_unsym1", fake_unsym)
    monkeypatch.setattr(utility, "write_matrix_image_Ctrl_sym1", fake_sym)
    S = np.load(os.path.join(os.path.dirname(__file__), "golden", "select_cases.npz"))
    position, x = S["position"], S["x"]
    region_list = [[80000, 250000, 80000, 250000, 0, 0, 7, 0],      # diagonal block
                   [60000, 150000, 200000, 330000, 0, 0, 8, 1]]     # off-diagonal block

    class Q:
        def put(self, item):
            self.item = item

    q = Q()
    assert utility.load_data_chromosome_sub3(0, 21, region_list, x, position, [10000, 8, -1, -1, -1, 0.0], q)
    rid, samples, lenvec, edges = q.item
    idx = S["sel0_idx"]
    lo, hi = position[idx, :2].min(), position[idx, :2].max()
    W = hi - lo + 1
    assert rid == 0 and samples.shape == (W * (W + 1) // 2, 3)
    assert lenvec == [samples.shape[0], W, W, np.min(position[idx]), np.min(position[idx]), 7, 1, 21]
    assert utility.load_data_chromosome_sub3(1, 21, region_list, x, position, [10000, 8, -1, -1, -1, 0.0], q)
    rid, samples, lenvec, edges = q.item
    idx = S["sel1_idx"]
    p = position[idx]
    assert rid == 1 and lenvec[1:] == [p[:, 0].max() - p[:, 0].min() + 1, p[:, 1].max() - p[:, 1].min() + 1,
                                       p[:, 0].min(), p[:, 1].min(), 8, 0, 21]
