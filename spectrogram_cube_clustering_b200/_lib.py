"""ctypes binding of ``libscc_b200.so`` (the C ABI in ``include/scc_b200.h``).

The library is built in-tree by ``csrc/Makefile`` (``__graft_entry__.build()``)
and is the ONLY compute backend: if it is missing or fails to load, importing
the ops raises — there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# SCC_LIB selects another build of the same ABI (the `make timeline` profiling variant); never a fallback.
LIB_PATH = os.environ.get("SCC_LIB") or os.path.join(_HERE, "libscc_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "scc_b200.h")

SCC_OK = 0
MAX_D, MAX_K = 32, 16

_lib = None


class SccError(RuntimeError):
    pass


def build(jobs: int | None = None, force: bool = False) -> str:
    """Compile the CUDA sources for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=True)
    jobs = jobs or max(1, (os.cpu_count() or 2))
    res = subprocess.run(["make", "-C", CSRC, f"-j{jobs}"], capture_output=True, text=True)
    if res.returncode != 0:
        raise SccError("building libscc_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    return LIB_PATH


_FP = POINTER(c_float)
_DP = POINTER(c_double)
_IP = POINTER(c_int32)

class SccExchange(ctypes.Structure):
    """Mirror of ``scc_exchange`` (include/scc_b200.h)."""
    _fields_ = [("windows", c_void_p), ("rank", c_int), ("world", c_int), ("max_len", c_int)]


_SIGNATURES = {
    "scc_abi_version": (c_int, []),
    "scc_status_string": (c_char_p, [c_int]),
    "scc_last_cuda_error": (c_char_p, []),
    "scc_supported": (c_int, [c_int, c_int]),
    "scc_debug_set_timeline": (c_int, [c_void_p]),
    "scc_gmm_supported": (c_int, [c_int, c_int]),
    "scc_workspace_bytes": (c_size_t, [c_int, c_int]),
    "scc_workspace_init": (c_int, [c_void_p, c_size_t, c_void_p]),
    "scc_dec_assign": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scc_dec_target": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "scc_colsum": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scc_dec_kl_grad": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int,
                                c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scc_dec_backward": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_size_t, c_void_p]),
    "scc_kmeans_step": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_size_t, c_void_p]),
    "scc_kmeans_batch_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "scc_kmeans_batch_step": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "scc_kmeans_batch_update": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "scc_dec_distances": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "scc_dec_assign_f64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_double, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scc_dec_target_f64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                   c_void_p]),
    "scc_dec_grad_f64": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_double, c_void_p, c_void_p, c_int,
                                 c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scc_dec_assign_ex": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "scc_dec_target_ex": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "scc_dec_kl_grad_ex": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int,
                                   c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "scc_dec_target_kl_grad": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_void_p, c_int, c_float,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "scc_dec_assign_u": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_int, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "scc_dec_target_kl_grad_u": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_int,
                                         c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                         c_void_p]),
    "scc_dec_step": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_int, c_float, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scc_dec_step_ex": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_float, c_int, c_float, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "scc_peer_finish": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "scc_peer_window_bytes": (c_size_t, [c_int]),
    "scc_peer_allreduce": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "scc_gmm_em_step": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "scc_gmm_finalize": (c_int, [c_void_p, c_double, c_int, c_int, c_double, c_double, c_double, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scc_gmm_em_iteration": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_int, c_double, c_double,
                                     c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_size_t, c_void_p, c_void_p]),
    "scc_gmm_pack_params": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load():
    """Load the shared library (once).  Raises SccError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SccError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise SccError(f"cannot load {LIB_PATH}: {exc}. There is no CPU fallback.") from exc
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == SCC_OK:
        return
    lib = load()
    msg = lib.scc_status_string(rc).decode()
    if rc == -5:
        msg += ": " + lib.scc_last_cuda_error().decode()
    raise SccError(f"{what} failed: {msg} (status {rc})")
