"""CPU checks of the C-ABI boundary: the shared library loads, exports every symbol that
include/scc_b200.h declares, and validates arguments without touching a GPU."""
import ctypes
import os
import re

import pytest

from spectrogram_cube_clustering_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "scc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 14
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} declared in include/scc_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes signature table out of sync with the header"


def test_version_and_status_strings(lib):
    assert lib.scc_abi_version() == 3
    assert lib.scc_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert lib.scc_status_string(code) not in (b"ok", b"unknown status")


def test_supported_shapes(lib):
    for d in (4, 8, 9, 10, 12, 16, 20, 24, 32):
        assert lib.scc_supported(d, 8) == 1
    assert lib.scc_supported(9, 17) == 0 and lib.scc_supported(33, 8) == 0 and lib.scc_supported(7, 8) == 0
    assert lib.scc_gmm_supported(9, 16) == 1
    assert lib.scc_workspace_bytes(9, 8) > 0 and lib.scc_workspace_bytes(0, 8) == 0


def test_argument_validation_without_gpu(lib):
    """Invalid calls are rejected before any CUDA work (so this runs on a CPU-only box)."""
    buf = (ctypes.c_double * 64)()
    z = (ctypes.c_float * 64)()
    p = ctypes.addressof
    # K out of range
    assert lib.scc_dec_assign(p(z), 4, 9, p(z), 17, 1.0, 0, None, None, None, p(buf), p(buf), 1 << 20, None) == -1
    # alpha <= 0
    assert lib.scc_dec_assign(p(z), 4, 9, p(z), 8, 0.0, 0, None, None, None, p(buf), p(buf), 1 << 20, None) == -1
    # unsupported d
    assert lib.scc_dec_assign(p(z), 4, 7, p(z), 8, 1.0, 0, None, None, None, p(buf), p(buf), 1 << 20, None) == -2
    # workspace too small
    assert lib.scc_dec_assign(p(z), 4, 9, p(z), 8, 1.0, 0, None, None, None, p(buf), p(buf), 16, None) == -4
    # misaligned z
    assert lib.scc_dec_assign(p(z) + 4, 4, 9, p(z), 8, 1.0, 0, None, None, None, p(buf), p(buf), 1 << 30, None) == -3
    # bad rounding flag
    assert lib.scc_dec_target(p(z), 4, 8, p(buf), 3, p(z), None) == -1
    # kl_grad needs p or f
    assert lib.scc_dec_kl_grad(p(z), 4, 9, p(z), 8, 1.0, None, None, 0, 1.0, None, p(buf), p(buf), 1 << 30, None) == -1
    assert lib.scc_gmm_em_step(p(z), 4, 9, 8, None, p(buf), None, None, None, 1, p(buf), 1 << 30, None) == -1
    # one-pass target + gradient: needs the column sums; bad rounding flag
    assert lib.scc_dec_target_kl_grad(p(z), 4, 9, p(z), 8, 1.0, None, 0, 1.0, None, None, p(buf), p(buf), 1 << 30,
                                      None, None, None) == -1
    assert lib.scc_dec_target_kl_grad(p(z), 4, 9, p(z), 8, 1.0, p(buf), 3, 1.0, None, None, p(buf), p(buf), 1 << 30,
                                      None, None, None) == -1
    # one-kernel step: f_stats is mandatory; bad rounding flag; unsupported d; an empty batch is a no-op
    assert lib.scc_dec_step(p(z), 4, 9, p(z), 8, 1.0, 0, 1.0, None, None, None, None, None, None, p(buf), p(buf),
                            1 << 30, None) == -1
    assert lib.scc_dec_step(p(z), 4, 9, p(z), 8, 1.0, 4, 1.0, None, None, None, p(buf), None, None, p(buf), p(buf),
                            1 << 30, None) == -1
    assert lib.scc_dec_step(p(z), 4, 7, p(z), 8, 1.0, 0, 1.0, None, None, None, p(buf), None, None, p(buf), p(buf),
                            1 << 30, None) == -2
    assert lib.scc_debug_set_timeline(None) == -2          # production build: the profiling stamps are compiled out


def test_dec_step_supported_shapes():
    from spectrogram_cube_clustering_b200 import ops
    assert ops.dec_step_supported(9, 8) and ops.dec_step_supported(32, 4) and ops.dec_step_supported(10, 16)
    assert not ops.dec_step_supported(32, 16) and not ops.dec_step_supported(12, 16) and not ops.dec_step_supported(7, 8)


def test_ops_refuse_cpu_tensors():
    import torch
    from spectrogram_cube_clustering_b200 import ops
    with pytest.raises(_lib.SccError):
        ops.dec_assign(torch.zeros(4, 9), torch.zeros(8, 9))
