"""Shared pytest configuration: the ``gpu`` marker and golden-fixture loading."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
DEC_CASES = ["c1", "k5", "d32", "alpha2", "alpha05", "relu", "tie"]
GMM_CASES = ["c1", "k16", "d32", "relu"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(kind, name):
    with np.load(os.path.join(GOLDEN_DIR, f"{kind}_{name}.npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(params=DEC_CASES)
def dec_golden(request):
    return request.param, load_golden("dec", request.param)


@pytest.fixture(params=GMM_CASES)
def gmm_golden(request):
    return request.param, load_golden("gmm", request.param)


def rel_err(a, b):
    """max |a-b| / max |b| — the max-normalised relative error used for the 1e-5 bar."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))
