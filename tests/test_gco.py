"""Host graph-cut step: the vendored GCO v3.0 behind include/phmrf_gco.h.

Known answers: gco_source/example.cpp (the only golden vectors in the reference tree):
six scenarios, 250 -> 44/44/44/170/44/244."""
import itertools
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "oracle", "_ref", "gco_example")


@pytest.fixture(scope="module")
def engine():
    from phylo_hmrf_b200 import build, engine
    build.build_gco()
    return engine


def test_vendored_gco_example_known_answers():
    if not os.path.exists(EXAMPLE):
        if not os.path.isdir("/root/reference/gco_source"):
            pytest.skip("reference tree not present and oracle/_ref/gco_example not built")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    out = subprocess.run([EXAMPLE], capture_output=True, text=True, check=True).stdout
    before = [int(x.split()[-1]) for x in out.splitlines() if x.startswith("Before")]
    after = [int(x.split()[-1]) for x in out.splitlines() if x.startswith("After")]
    assert before == [250] * 6
    assert after == [44, 44, 44, 170, 44, 244]


def _example_general_graph():
    """GeneralGraph_DArraySArraySpatVarying of example.cpp:335-395 (10x5 grid, 7 labels)."""
    width, height, L = 10, 5, 7
    n = width * height
    data = np.full((n, L), 10, dtype=np.int32)
    data[:25, 0] = 0
    data[25:, 5] = 0
    l = np.arange(L)
    smooth = np.minimum((l[:, None] - l[None, :]) ** 2, 4).astype(np.int32)
    e, w = [], []
    for y in range(height):
        for x in range(1, width):
            p1, p2 = x - 1 + y * width, x + y * width
            e.append((p1, p2)); w.append(p1 + p2)
    for y in range(1, height):
        for x in range(width):
            p1, p2 = x + (y - 1) * width, x + y * width
            e.append((p1, p2)); w.append(p1 * p2)
    return data, np.asarray(e, dtype=np.int64), np.asarray(w, dtype=np.int32), smooth


def _energy(unary, e, w, V, lab):
    return int(unary[np.arange(len(lab)), lab].sum() + (V[lab[e[:, 0]], lab[e[:, 1]]].astype(np.int64) * w).sum())


def test_wrapper_reproduces_example_energies(engine):
    data, e, w, smooth = _example_general_graph()
    lab, en, en0 = engine.gco_cut_int(data, e, w, smooth, n_iter=2, algorithm='expansion', return_energy=True)
    assert en0 == 250 and en == 244
    assert _energy(data, e, w, smooth, lab) == 244


def test_swap_from_init_labels_is_deterministic_and_monotone(engine):
    rng = np.random.default_rng(5)
    n, K = 60, 5
    unary = rng.integers(0, 1000, size=(n, K)).astype(np.int32)
    e = np.asarray([(i, j) for i in range(n) for j in range(i + 1, min(n, i + 4))], dtype=np.int64)
    w = rng.integers(1, 50, size=len(e)).astype(np.int32)
    V = (30 * (1 - np.eye(K))).astype(np.int32)
    init = rng.integers(0, K, size=n).astype(np.int32)
    lab1, en1, en0 = engine.gco_cut_int(unary, e, w, V, n_iter=5000, algorithm='swap', init_labels=init,
                                        return_energy=True)
    lab2 = engine.gco_cut_int(unary, e, w, V, n_iter=5000, algorithm='swap', init_labels=init)
    assert np.array_equal(lab1, lab2)
    assert en0 == _energy(unary, e, w, V, init)
    assert en1 == _energy(unary, e, w, V, lab1) <= en0
    # a swap-converged labelling is single-site optimal
    for i in range(n):
        for k in range(K):
            trial = lab1.copy(); trial[i] = k
            assert _energy(unary, e, w, V, trial) >= en1


def test_two_label_swap_is_globally_optimal(engine):
    rng = np.random.default_rng(9)
    n = 10
    unary = rng.integers(0, 100, size=(n, 2)).astype(np.int32)
    e = np.asarray([(i, i + 1) for i in range(n - 1)] + [(0, 5), (2, 7)], dtype=np.int64)
    w = rng.integers(1, 40, size=len(e)).astype(np.int32)
    V = np.asarray([[0, 1], [1, 0]], dtype=np.int32)
    lab, en, _ = engine.gco_cut_int(unary, e, w, V, n_iter=100, algorithm='swap',
                                    init_labels=np.zeros(n, np.int32), return_energy=True)
    best = min(_energy(unary, e, w, V, np.asarray(c)) for c in itertools.product((0, 1), repeat=n))
    assert en == best


def test_errors_come_back_as_exceptions_not_exits(engine):
    unary = np.zeros((4, 3), dtype=np.int32)
    V = np.zeros((3, 3), dtype=np.int32)
    with pytest.raises(RuntimeError, match="id1 < id2"):
        engine.gco_cut_int(unary, np.asarray([[2, 1]]), np.asarray([1], np.int32), V, algorithm='swap')
    with pytest.raises(RuntimeError, match="init label"):
        engine.gco_cut_int(unary, np.asarray([[0, 1]]), np.asarray([1], np.int32), V, algorithm='swap',
                           init_labels=np.asarray([0, 1, 2, 3]))
    # GCException (smoothness term beyond GCO_MAX_ENERGYTERM, GCoptimization.cpp:288-310) is caught
    # inside the wrapper instead of terminating the process
    big = (5000 * (1 - np.eye(3))).astype(np.int32)
    with pytest.raises(RuntimeError, match="GCO"):
        engine.gco_cut_int(unary, np.asarray([[0, 1]]), np.asarray([20000000], np.int32), big, algorithm='swap',
                           init_labels=np.asarray([0, 1, 2, 0]))
    with pytest.raises(TypeError):
        engine.cut_general_graph(np.asarray([[0, 1]]), np.asarray([0.5]), unary, V, algorithm='swap')
