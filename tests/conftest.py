import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["case_diag_w", "case_two_regions", "case_unweighted_iso", "case_d9_k30"]


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    import numpy as np
    return np.load(os.path.join(GOLDEN_DIR, request.param + ".npz"), allow_pickle=False)
