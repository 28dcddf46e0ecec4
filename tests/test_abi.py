"""CPU checks of the drop-in boundary: the C-ABI libraries load, export every symbol the
headers declare, and refuse to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from phylo_hmrf_b200 import build
    return build.build_all()


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(phmrf_[A-Za-z0-9_]+)\s*\(", src)))


def test_libphmrf_exports_every_declared_symbol(built):
    from phylo_hmrf_b200 import _lib
    names = _declared("phmrf.h")
    assert len(names) >= 20
    lib = C.CDLL(built[0])
    for n in names:
        assert hasattr(lib, n), "libphmrf.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table out of step with include/phmrf.h"
    assert _lib.lib().phmrf_abi_version() == 1


def test_libphmrf_gco_exports_every_declared_symbol(built):
    from phylo_hmrf_b200 import _lib
    names = _declared("phmrf_gco.h")
    lib = C.CDLL(built[1])
    for n in names:
        assert hasattr(lib, n)
    assert sorted(_lib.GCO_SIGNATURES) == names


def test_probe_library_is_separate_from_the_product(built):
    """The pipe probes (roofline denominators for bench.py) live in their own library; the product library
    exports none of them."""
    from phylo_hmrf_b200 import _lib
    names = _declared("phmrf_probe.h")
    lib = C.CDLL(built[2])
    for n in names:
        assert hasattr(lib, n)
    assert sorted(_lib.PROBE_SIGNATURES) == names
    product = C.CDLL(built[0])
    assert not any(hasattr(product, n) for n in names)


def test_headers_are_plain_c():
    """The boundary is a C ABI: every header under include/ must compile as C99 (no C++ types in a signature)."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):
        subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", h)], check=True)


def test_no_cpu_fallback_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import phylo_hmrf_b200 as ph
    with pytest.raises(ph.PhmrfError) as ei:
        ph.Model(4, 3)
    assert ei.value.code == -2
    with pytest.raises(ph.PhmrfError):
        ph.log_multivariate_normal_density(np.zeros((4, 3)), np.zeros((2, 3)), np.stack([np.eye(3)] * 2))


def test_host_staging_needs_a_device_and_rejects_nonsense(built):
    """phmrf_host_alloc: error code (never a crash) without a CUDA device; bad arguments are rejected."""
    import torch
    from phylo_hmrf_b200 import _lib, engine
    p = C.c_void_p()
    assert _lib.lib().phmrf_host_alloc(-1, C.byref(p)) == -1
    assert _lib.lib().phmrf_host_free(None) == 0
    if not torch.cuda.is_available():
        assert _lib.lib().phmrf_host_alloc(1024, C.byref(p)) == -2 and not p.value
        with pytest.raises(_lib.PhmrfError):
            engine.pinned_empty((4, 4), np.int32)


def test_invalid_arguments_are_rejected(built):
    from phylo_hmrf_b200 import _lib
    h = C.c_void_p()
    assert _lib.lib().phmrf_ctx_create(0, 0, 3, C.byref(h)) == -1
    assert _lib.lib().phmrf_ctx_create(0, 4, 13, C.byref(h)) == -5
    assert b"n_features" in _lib.lib().phmrf_last_error()


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from phylo_hmrf_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.lib()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under phylo_hmrf_b200/ may reference it."""
    pkg = os.path.join(ROOT, "phylo_hmrf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "oracle" not in text.replace("test infrastructure", ""), "%s mentions the oracle" % f
