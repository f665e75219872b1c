"""Device operators of the clustering hot path: torch tensors in, C-ABI launches out.

Every function takes CUDA tensors, launches the hand-written sm_100a kernels of
``libscc_b200.so`` on the CURRENT torch stream through the C ABI
(``include/scc_b200.h``) and returns freshly allocated CUDA tensors.  They are
also registered as PyTorch custom ops under ``torch.ops.scc_b200``.
There is no CPU path: a non-CUDA tensor raises.
"""
from __future__ import annotations

import torch

from . import _lib

SUPPORTED_DIMS = (4, 8, 9, 10, 12, 16, 20, 24, 32)
MAX_K = _lib.MAX_K

_workspaces: dict = {}


def _ptr(t):
    return None if t is None else t.data_ptr()


def _ex(desc):
    """ctypes pointer to an ``scc_exchange`` descriptor (or NULL)."""
    import ctypes
    return None if desc is None else ctypes.addressof(desc)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require(t: torch.Tensor, name: str, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.SccError(f"{name} must be a CUDA tensor: the clustering ops have no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def workspace(device, d: int, K: int) -> torch.Tensor:
    """Per-(device, stream, d, K) scratch buffer for the grid reductions."""
    lib = _lib.load()
    key = (torch.device(device).index, _stream(), d, K)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = lib.scc_workspace_bytes(d, K)
        if nbytes == 0:
            raise ValueError(f"unsupported shape d={d}, K={K}")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(lib.scc_workspace_init(ws.data_ptr(), nbytes, _stream()), "scc_workspace_init")
        _workspaces[key] = ws
    return ws


def padded_dim(d: int) -> int:
    """Smallest instantiated latent dimension >= d (zero-padding is exact for DEC)."""
    for s in SUPPORTED_DIMS:
        if s >= d:
            return s
    raise ValueError(f"latent dimension {d} exceeds {_lib.MAX_D}")


# --------------------------------------------------------------------------- DEC, float64 precision path
def dec_assign_f64(z, mu, alpha=1.0, round_decimals=0, want_q=True, want_labels=True, labels_prev=None):
    """float64 :func:`dec_assign` (the reference's own dtype, ``models.py:965``): q / labels / f in IEEE float64 and the
    reference's operation order.  z [n,d], mu [K,d] float64; any d <= 32, K <= 16."""
    lib = _lib.load()
    _require(z, "z", torch.float64); _require(mu, "mu", torch.float64)
    n, d = z.shape
    K = mu.shape[0]
    if mu.shape[1] != d or d > _lib.MAX_D or K > MAX_K:
        raise ValueError("mu must be [K <= 16, d] with the latent dimension d <= 32 of z")
    q = torch.empty(n, K, dtype=torch.float64, device=z.device) if want_q else None
    labels = torch.empty(n, dtype=torch.int32, device=z.device) if want_labels else None
    if labels_prev is not None:
        _require(labels_prev, "labels_prev", torch.int32)
    stats = torch.empty(K + 1, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, padded_dim(d), K)
    rc = lib.scc_dec_assign_f64(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), int(round_decimals), _ptr(q),
                                _ptr(labels), _ptr(labels_prev), stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "scc_dec_assign_f64")
    return q, labels, stats


def dec_target_f64(q, f=None, round_decimals=0):
    """float64 ``target_distribution`` (``models.py:1320-1322``): q [n,K] float64 -> (p [n,K] float64, f [K]).  ``f`` =
    column sums of q; computed from q when not given."""
    lib = _lib.load()
    _require(q, "q", torch.float64)
    n, K = q.shape
    have_f = f is not None
    if have_f:
        _require(f, "f", torch.float64)
    else:
        f = torch.empty(K, dtype=torch.float64, device=q.device)
    p = torch.empty_like(q)
    ws = workspace(q.device, 4, K)
    rc = lib.scc_dec_target_f64(q.data_ptr(), n, K, f.data_ptr(), int(have_f), int(round_decimals), p.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "scc_dec_target_f64")
    return p, f


def dec_grad_f64(z, mu, alpha=1.0, p=None, f=None, grad_q=None, round_decimals=0, scale=1.0, want_p=False, want_dz=True):
    """float64 gradients: exactly one of ``p`` (target), ``f`` (target rebuilt from the column sums) or ``grad_q``
    (generic upstream dL/dq).  -> (stats float64 [K*d+2] = loss, sum_i s_i, dmu; dz [n,d] | None; p_out | None)."""
    lib = _lib.load()
    _require(z, "z", torch.float64); _require(mu, "mu", torch.float64)
    n, d = z.shape
    K = mu.shape[0]
    for t, nm in ((p, "p"), (f, "f"), (grad_q, "grad_q")):
        if t is not None:
            _require(t, nm, torch.float64)
    dz = torch.empty_like(z) if want_dz else None
    p_out = torch.empty(n, K, dtype=torch.float64, device=z.device) if (want_p and f is not None) else None
    stats = torch.empty(K * d + 2, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, padded_dim(d), K)
    rc = lib.scc_dec_grad_f64(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), _ptr(p), _ptr(f), int(round_decimals),
                              _ptr(grad_q), float(scale), _ptr(p_out), _ptr(dz), stats.data_ptr(), ws.data_ptr(),
                              ws.numel(), _stream())
    _lib.check(rc, "scc_dec_grad_f64")
    return stats, dz, p_out


# --------------------------------------------------------------------------- DEC
def dec_assign(z, mu, alpha=1.0, round_decimals=0, want_q=True, want_labels=True, labels_prev=None,
               out_q=None, out_labels=None, out_stats=None, push=None):
    """z [n,d], mu [K,d] -> (q [n,K] | None, labels int32 [n] | None, stats float64 [K+1]).

    stats = (f_0..f_{K-1}, number of labels != labels_prev).  networks.py:279-288,
    models.py:92,94,1098-1099,1320.
    """
    lib = _lib.load()
    _require(z, "z"); _require(mu, "mu")
    n, d = z.shape
    K = mu.shape[0]
    if mu.shape[1] != d:
        raise ValueError("mu and z disagree on the latent dimension")
    q = out_q if out_q is not None else (torch.empty(n, K, dtype=torch.float32, device=z.device) if want_q else None)
    labels = out_labels if out_labels is not None else (
        torch.empty(n, dtype=torch.int32, device=z.device) if want_labels else None)
    if labels_prev is not None:
        _require(labels_prev, "labels_prev", torch.int32)
    stats = out_stats if out_stats is not None else torch.empty(K + 1, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, d, K)
    rc = lib.scc_dec_assign_ex(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), int(round_decimals), _ptr(q),
                               _ptr(labels), _ptr(labels_prev), stats.data_ptr(), ws.data_ptr(), ws.numel(),
                               _ex(push), _stream())
    _lib.check(rc, "scc_dec_assign")
    return q, labels, stats


def dec_target(q, f, round_decimals=0, out=None, pull=None):
    """q [n,K], f float64 [>=K] -> p [n,K].  models.py:1320-1322.  With ``pull`` (an exchange
    descriptor) f [K+1] is first filled with the all-reduced sums pushed by the preceding
    ``dec_assign(push=...)`` on every rank."""
    lib = _lib.load()
    _require(q, "q"); _require(f, "f", torch.float64)
    n, K = q.shape
    p = out if out is not None else torch.empty_like(q)
    rc = lib.scc_dec_target_ex(q.data_ptr(), n, K, f.data_ptr(), int(round_decimals), p.data_ptr(), _ex(pull),
                               _stream())
    _lib.check(rc, "scc_dec_target")
    return p


def colsum(q):
    """f_j = sum_i q_ij (float64 [K]) of a caller-supplied q.  models.py:1320."""
    lib = _lib.load()
    _require(q, "q")
    n, K = q.shape
    f = torch.empty(K, dtype=torch.float64, device=q.device)
    ws = workspace(q.device, 4, K)
    _lib.check(lib.scc_colsum(q.data_ptr(), n, K, f.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "scc_colsum")
    return f


def dec_kl_grad(z, mu, alpha=1.0, p=None, f=None, round_decimals=0, scale=1.0, want_dz=True,
                out_dz=None, out_stats=None, pull_f=None, push=None):
    """-> (stats float64 [K*d+2] = (loss, sum_i s_i, dmu[K,d]), dz [n,d] | None).  models.py:1124-1127."""
    lib = _lib.load()
    _require(z, "z"); _require(mu, "mu")
    n, d = z.shape
    K = mu.shape[0]
    if p is not None:
        _require(p, "p")
        if tuple(p.shape) != (n, K):
            raise ValueError("p must be [n, K]")
    if f is not None:
        _require(f, "f", torch.float64)
    if p is None and f is None and pull_f is None:
        raise ValueError("need the target p or the column sums f")
    dz = out_dz if out_dz is not None else (torch.empty_like(z) if want_dz else None)
    stats = out_stats if out_stats is not None else torch.empty(K * d + 2, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, d, K)
    rc = lib.scc_dec_kl_grad_ex(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), _ptr(p), _ptr(f),
                                int(round_decimals), float(scale), _ptr(dz), stats.data_ptr(), ws.data_ptr(),
                                ws.numel(), _ex(pull_f), _ex(push), _stream())
    _lib.check(rc, "scc_dec_kl_grad")
    return stats, dz


def dec_target_kl_grad(z, mu, f, alpha=1.0, round_decimals=0, scale=1.0, want_p=True, want_dz=True,
                       out_p=None, out_dz=None, out_stats=None, pull_f=None, push=None):
    """target_distribution + KL loss + gradients in one pass over z (models.py:1302-1322 + 1124-1127).

    -> (stats float64 [K*d+2] = (loss, sum_i s_i, dmu[K,d]), p [n,K] | None, dz [n,d] | None).
    ``f`` are the column sums from :func:`dec_assign` for the same z and mu."""
    lib = _lib.load()
    _require(z, "z"); _require(mu, "mu")
    n, d = z.shape
    K = mu.shape[0]
    if f is not None:
        _require(f, "f", torch.float64)
    elif pull_f is None:
        raise ValueError("need the column sums f")
    p = out_p if out_p is not None else (torch.empty(n, K, dtype=torch.float32, device=z.device) if want_p else None)
    dz = out_dz if out_dz is not None else (torch.empty_like(z) if want_dz else None)
    stats = out_stats if out_stats is not None else torch.empty(K * d + 2, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, d, K)
    rc = lib.scc_dec_target_kl_grad(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), _ptr(f), int(round_decimals),
                                    float(scale), _ptr(p), _ptr(dz), stats.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _ex(pull_f), _ex(push), _stream())
    _lib.check(rc, "scc_dec_target_kl_grad")
    return stats, p, dz


def dec_assign_u(z, mu, u_out, alpha=1.0, round_decimals=0, want_q=False, want_labels=True, labels_prev=None,
                 out_q=None, out_labels=None, out_stats=None, push=None):
    """:func:`dec_assign` that also hands u_ij = 1 / (1 + d_ij / alpha) ([n, K] float32) to the gradient pass
    (:func:`dec_target_kl_grad_u`) — the two-launch step of the shapes the one-kernel step does not cover."""
    lib = _lib.load()
    _require(z, "z"); _require(mu, "mu"); _require(u_out, "u_out")
    n, d = z.shape
    K = mu.shape[0]
    if tuple(u_out.shape) != (n, K):
        raise ValueError("u_out must be [n, K]")
    q = out_q if out_q is not None else (torch.empty(n, K, dtype=torch.float32, device=z.device) if want_q else None)
    labels = out_labels if out_labels is not None else (
        torch.empty(n, dtype=torch.int32, device=z.device) if want_labels else None)
    if labels_prev is not None:
        _require(labels_prev, "labels_prev", torch.int32)
    stats = out_stats if out_stats is not None else torch.empty(K + 1, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, d, K)
    rc = lib.scc_dec_assign_u(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), int(round_decimals), _ptr(q),
                              _ptr(labels), _ptr(labels_prev), u_out.data_ptr(), stats.data_ptr(), ws.data_ptr(),
                              ws.numel(), _ex(push), _stream())
    _lib.check(rc, "scc_dec_assign_u")
    return q, labels, stats


def dec_target_kl_grad_u(z, mu, u, f, alpha=1.0, round_decimals=0, scale=1.0, want_p=False, want_dz=False,
                         out_p=None, out_dz=None, out_stats=None, pull_f=None, push=None):
    """:func:`dec_target_kl_grad` with the Student's-t u_ij streamed from ``u`` (written by :func:`dec_assign_u` for
    the same z and mu) instead of recomputed."""
    lib = _lib.load()
    _require(z, "z"); _require(mu, "mu"); _require(u, "u")
    n, d = z.shape
    K = mu.shape[0]
    if f is not None:
        _require(f, "f", torch.float64)
    elif pull_f is None:
        raise ValueError("need the column sums f")
    p = out_p if out_p is not None else (torch.empty(n, K, dtype=torch.float32, device=z.device) if want_p else None)
    dz = out_dz if out_dz is not None else (torch.empty_like(z) if want_dz else None)
    stats = out_stats if out_stats is not None else torch.empty(K * d + 2, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, d, K)
    rc = lib.scc_dec_target_kl_grad_u(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), u.data_ptr(), _ptr(f),
                                      int(round_decimals), float(scale), _ptr(p), _ptr(dz), stats.data_ptr(),
                                      ws.data_ptr(), ws.numel(), _ex(pull_f), _ex(push), _stream())
    _lib.check(rc, "scc_dec_target_kl_grad_u")
    return stats, p, dz


def dec_step_supported(d: int, K: int) -> bool:
    """Shapes the one-kernel step exists for (the register-blocked gradient kernel: K_padded * d <= 160)."""
    kp = 4 if K <= 4 else (8 if K <= 8 else 16)
    return d in SUPPORTED_DIMS and K <= MAX_K and kp * d <= 160


def dec_step(z, mu, alpha=1.0, round_decimals=0, scale=1.0, want_q=True, want_labels=True, want_p=True, want_dz=True,
             labels_prev=None, out_q=None, out_labels=None, out_p=None, out_dz=None, out_f=None, out_stats=None,
             exchange=None):
    """The whole DEC step of one batch in ONE (cooperative) kernel launch — assign pass, grid-wide all-reduce of
    f, target + KL-gradient pass.  -> dict(q, labels, f [K+1 float64: f_j, label changes], p, dz,
    stats [K*d+2 float64: loss, sum_i s_i, dmu]).  networks.py:279-288 + models.py:1302-1322 + 1124-1127.
    With ``exchange`` (a peer-exchange descriptor, one process per GPU) f and the gradient statistics are
    all-reduced over the GPUs INSIDE the kernel (flag-in-data exchange over NVLink peer memory, at the grid
    barrier and in the kernel's tail): ``scale`` must be gamma / N_total."""
    lib = _lib.load()
    _require(z, "z"); _require(mu, "mu")
    n, d = z.shape
    K = mu.shape[0]
    dev = z.device
    q = out_q if out_q is not None else (torch.empty(n, K, dtype=torch.float32, device=dev) if want_q else None)
    labels = out_labels if out_labels is not None else (torch.empty(n, dtype=torch.int32, device=dev) if want_labels else None)
    p = out_p if out_p is not None else (torch.empty(n, K, dtype=torch.float32, device=dev) if want_p else None)
    dz = out_dz if out_dz is not None else (torch.empty_like(z) if want_dz else None)
    if labels_prev is not None:
        _require(labels_prev, "labels_prev", torch.int32)
    f = out_f if out_f is not None else torch.empty(K + 1, dtype=torch.float64, device=dev)
    stats = out_stats if out_stats is not None else torch.empty(K * d + 2, dtype=torch.float64, device=dev)
    ws = workspace(dev, d, K)
    rc = lib.scc_dec_step_ex(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), int(round_decimals), float(scale),
                             _ptr(q), _ptr(labels), _ptr(labels_prev), f.data_ptr(), _ptr(p), _ptr(dz),
                             stats.data_ptr(), ws.data_ptr(), ws.numel(), _ex(exchange), _stream())
    _lib.check(rc, "scc_dec_step")      # with an exchange both all-reduces (f, statistics) completed inside the kernel
    return dict(q=q, labels=labels, f=f, p=p, dz=dz, stats=stats)


def dec_backward(z, mu, grad_q, alpha=1.0, want_dz=True):
    """Layer backward for an arbitrary dL/dq -> (dz [n,d] | None, dmu float64 [K,d])."""
    lib = _lib.load()
    _require(z, "z"); _require(mu, "mu"); _require(grad_q, "grad_q")
    n, d = z.shape
    K = mu.shape[0]
    dz = torch.empty_like(z) if want_dz else None
    stats = torch.empty(K * d + 2, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, d, K)
    rc = lib.scc_dec_backward(z.data_ptr(), n, d, mu.data_ptr(), K, float(alpha), grad_q.data_ptr(), _ptr(dz),
                              stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "scc_dec_backward")
    return dz, stats[2:].view(K, d)


def kmeans_step(z, centers, labels=None, mindist=None, out_stats=None):
    """One Lloyd step -> stats float64 [K*d+2+K] = (inertia, 0, shift[K,d], count[K]).  models.py:386-394."""
    lib = _lib.load()
    _require(z, "z"); _require(centers, "centers")
    n, d = z.shape
    K = centers.shape[0]
    stats = out_stats if out_stats is not None else torch.empty(K * d + 2 + K, dtype=torch.float64, device=z.device)
    ws = workspace(z.device, d, K)
    rc = lib.scc_kmeans_step(z.data_ptr(), n, d, centers.data_ptr(), K, _ptr(labels), _ptr(mindist),
                             stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "scc_kmeans_step")
    return stats


_batch_workspaces: dict = {}


def _kmeans_batch_workspace(device, d: int, K: int, R: int) -> torch.Tensor:
    lib = _lib.load()
    key = (torch.device(device).index, _stream(), d, K, R)
    ws = _batch_workspaces.get(key)
    if ws is None:
        nbytes = lib.scc_kmeans_batch_workspace_bytes(d, K, R)
        if nbytes == 0:
            raise ValueError(f"unsupported shape d={d}, K={K}, restarts={R}")
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)        # ticket counters start at zero
        _batch_workspaces[key] = ws
    return ws


def kmeans_batch_step(z, centers, done=None, labels=None, mindist=None, out_stats=None):
    """One Lloyd scan of R restarts at once: centers [R, K, d] -> stats float64 [R, K*d+2+K] (per restart:
    inertia, 0, shift[K,d], count[K]); restarts with ``done[r] != 0`` are skipped.  models.py:386-394."""
    lib = _lib.load()
    _require(z, "z"); _require(centers, "centers")
    if centers.dim() != 3 or centers.shape[2] != z.shape[1]:
        raise ValueError("centers must be [restarts, K, d]")
    n, d = z.shape
    R, K = centers.shape[0], centers.shape[1]
    if done is not None:
        _require(done, "done", torch.uint8)
    if labels is not None:
        _require(labels, "labels", torch.int32)
    if mindist is not None:
        _require(mindist, "mindist")
    stats = out_stats if out_stats is not None else torch.zeros(R, K * d + 2 + K, dtype=torch.float64, device=z.device)
    ws = _kmeans_batch_workspace(z.device, d, K, R)
    rc = lib.scc_kmeans_batch_step(z.data_ptr(), n, d, centers.data_ptr(), K, R, _ptr(done), _ptr(labels),
                                   _ptr(mindist), stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "scc_kmeans_batch_step")
    return stats


def kmeans_batch_update(centers, stats, shift_tol, done, n_iter=None, inertia=None):
    """Move every unfinished restart's centres (c += shift / count), count the iteration and raise ``done[r]`` once
    the total squared centre shift is <= shift_tol (scikit-learn's rule) — on the device, no host sync."""
    lib = _lib.load()
    _require(centers, "centers"); _require(stats, "stats", torch.float64); _require(done, "done", torch.uint8)
    R, K, d = centers.shape
    rc = lib.scc_kmeans_batch_update(centers.data_ptr(), stats.data_ptr(), d, K, R, float(shift_tol), done.data_ptr(),
                                     _ptr(n_iter), _ptr(inertia), _stream())
    _lib.check(rc, "scc_kmeans_batch_update")


def dec_distances(z, mu, p=2.0, out=None):
    """D_ij = (sum_c |z_ic - mu_jc|^p)^(1/p)  -> [n, K] float32 (utils.py:866-869 for every centroid at once)."""
    lib = _lib.load()
    _require(z, "z"); _require(mu, "mu")
    n, d = z.shape
    K = mu.shape[0]
    if mu.shape[1] != d:
        raise ValueError("mu and z disagree on the latent dimension")
    if not p > 0:
        raise ValueError("p must be positive")
    out = out if out is not None else torch.empty(n, K, dtype=torch.float32, device=z.device)
    _lib.check(lib.scc_dec_distances(z.data_ptr(), n, d, mu.data_ptr(), K, float(p), out.data_ptr(), _stream()),
               "scc_dec_distances")
    return out


# --------------------------------------------------------------------------- GMM
def gmm_param_floats(K, d):
    return K * d + K * (d * (d + 1) // 2) + K


def gmm_stat_doubles(K, d):
    return 1 + K + K * d + K * (d * (d + 1) // 2)


def gmm_supported(d, K):
    return bool(_lib.load().scc_gmm_supported(d, K))


def gmm_pack_params(weights, means, covariances, params=None, prec_chol=None, ctrl=None):
    """(pi [K], mu [K,d], Sigma [K,d,d]) float64 -> packed float32 E-step block, U, ctrl."""
    lib = _lib.load()
    _require(weights, "weights", torch.float64); _require(means, "means", torch.float64)
    _require(covariances, "covariances", torch.float64)
    K, d = means.shape
    dev = means.device
    params = params if params is not None else torch.empty(gmm_param_floats(K, d), dtype=torch.float32, device=dev)
    prec_chol = prec_chol if prec_chol is not None else torch.empty(K, d, d, dtype=torch.float64, device=dev)
    ctrl = ctrl if ctrl is not None else torch.zeros(8, dtype=torch.float64, device=dev)
    rc = lib.scc_gmm_pack_params(weights.data_ptr(), means.data_ptr(), covariances.data_ptr(), d, K,
                                 prec_chol.data_ptr(), params.data_ptr(), ctrl.data_ptr(), _stream())
    _lib.check(rc, "scc_gmm_pack_params")
    return params, prec_chol, ctrl


GMM_ESTEP_ONLY, GMM_SOFT, GMM_HARD = 0, 1, 2
GMM_NOSKIP = 16      # OR-able: keep (point, component) pairs with responsibility < 2^-30 in the M-step sums


def gmm_em_step(z, K, params, stats=None, labels=None, resp=None, ctrl=None, mode=GMM_SOFT):
    """One fused E+M statistics pass.  Returns the packed float64 stats tensor."""
    lib = _lib.load()
    _require(z, "z"); _require(params, "params")
    n, d = z.shape
    stats = stats if stats is not None else torch.empty(gmm_stat_doubles(K, d), dtype=torch.float64, device=z.device)
    ws = workspace(z.device, d, K)
    rc = lib.scc_gmm_em_step(z.data_ptr(), n, d, K, params.data_ptr(), stats.data_ptr(), _ptr(labels), _ptr(resp),
                             _ptr(ctrl), int(mode), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "scc_gmm_em_step")
    return stats


def gmm_em_iteration(z, K, params, stats, n_total, means, weights, covariances, prec_chol, ctrl, mode=GMM_SOFT,
                     reg_covar=1e-6, nk_eps=10 * 2.220446049250313e-16, tol=1e-3, exchange=None):
    """One whole EM iteration in two launches: the fused statistics pass and ONE tail kernel (grid reduction,
    cross-GPU all-reduce through ``exchange``, M-step finalisation).  In place on every state tensor."""
    lib = _lib.load()
    _require(z, "z"); _require(params, "params")
    n, d = z.shape
    ws = workspace(z.device, d, K)
    rc = lib.scc_gmm_em_iteration(z.data_ptr(), n, d, K, params.data_ptr(), stats.data_ptr(), int(mode), float(n_total),
                                  float(reg_covar), float(nk_eps), float(tol), means.data_ptr(), weights.data_ptr(),
                                  covariances.data_ptr(), prec_chol.data_ptr(), ctrl.data_ptr(), ws.data_ptr(), ws.numel(),
                                  _ex(exchange), _stream())
    _lib.check(rc, "scc_gmm_em_iteration")


def gmm_finalize(stats, n_total, means, weights, covariances, prec_chol, params, ctrl,
                 reg_covar=1e-6, nk_eps=10 * 2.220446049250313e-16, tol=1e-3):
    """M-step finalisation on the device (in place on means/weights/covariances/prec_chol/params/ctrl)."""
    lib = _lib.load()
    K, d = means.shape
    rc = lib.scc_gmm_finalize(stats.data_ptr(), float(n_total), d, K, float(reg_covar), float(nk_eps), float(tol),
                              means.data_ptr(), weights.data_ptr(), covariances.data_ptr(), prec_chol.data_ptr(),
                              params.data_ptr(), ctrl.data_ptr(), _stream())
    _lib.check(rc, "scc_gmm_finalize")


# --------------------------------------------------------------------------- multi-GPU exchange
def peer_window_bytes(max_len: int) -> int:
    return int(_lib.load().scc_peer_window_bytes(int(max_len)))


def peer_allreduce(t: torch.Tensor, windows: torch.Tensor, rank: int, world: int, max_len: int) -> torch.Tensor:
    """In-place one-shot all-reduce(sum) of a float64 vector over NVLink peer memory."""
    lib = _lib.load()
    _require(t, "t", torch.float64)
    rc = lib.scc_peer_allreduce(t.data_ptr(), t.numel(), t.data_ptr(), windows.data_ptr(), int(rank), int(world),
                                int(max_len), _stream())
    _lib.check(rc, "scc_peer_allreduce")
    return t


def peer_finish(t: torch.Tensor, desc) -> torch.Tensor:
    """Completes a fused exchange: waits for every rank's push and writes the rank-ordered sum to t."""
    lib = _lib.load()
    _require(t, "t", torch.float64)
    _lib.check(lib.scc_peer_finish(t.data_ptr(), t.numel(), _ex(desc), _stream()), "scc_peer_finish")
    return t


# --------------------------------------------------------------------------- torch.library registration
# The launches as PyTorch custom ops (torch.ops.scc_b200.*; BASELINE.json north_star: "a thin C-ABI exposed as
# PyTorch custom ops").  Operands are float32, contiguous, CUDA — the Python layer (networks.py) casts and pads.
# `soft_assign` and `dec_kl_loss` carry autograd formulas (register_autograd), so ClusteringLayer / dec_kl_loss
# are differentiable through torch.ops and traceable (fake kernels registered); the others are plain launches.
def _register_custom_ops():
    from torch.library import custom_op

    # ---- ClusteringLayer.forward / backward (networks.py:279-288 + autograd)
    @custom_op("scc_b200::soft_assign", mutates_args=(), device_types="cuda")
    def _soft_assign(z: torch.Tensor, mu: torch.Tensor, alpha: float) -> torch.Tensor:
        if z.dtype == torch.float64:                 # the reference's dtype: float64 kernels
            return dec_assign_f64(z, mu, alpha, 0, want_labels=False)[0]
        q, _, _ = dec_assign(z, mu, alpha, 0, want_labels=False)
        return q

    @_soft_assign.register_fake
    def _(z, mu, alpha):
        return z.new_empty(z.shape[0], mu.shape[0])

    @custom_op("scc_b200::soft_assign_backward", mutates_args=(), device_types="cuda")
    def _soft_assign_backward(z: torch.Tensor, mu: torch.Tensor, grad_q: torch.Tensor, alpha: float) -> tuple[
            torch.Tensor, torch.Tensor]:
        if z.dtype == torch.float64:
            stats, dz, _ = dec_grad_f64(z, mu, alpha, grad_q=grad_q)
            return dz, stats[2:].view(mu.shape).clone()
        dz, dmu = dec_backward(z, mu, grad_q, alpha)
        return dz, dmu.to(torch.float32)

    @_soft_assign_backward.register_fake
    def _(z, mu, grad_q, alpha):
        return torch.empty_like(z), torch.empty_like(mu)

    def _sa_setup(ctx, inputs, output):
        z, mu, alpha = inputs
        ctx.save_for_backward(z, mu)
        ctx.alpha = alpha

    def _sa_backward(ctx, grad_q):
        z, mu = ctx.saved_tensors
        dz, dmu = torch.ops.scc_b200.soft_assign_backward(z, mu, grad_q.contiguous(), ctx.alpha)
        return dz, dmu, None

    _soft_assign.register_autograd(_sa_backward, setup_context=_sa_setup)

    # ---- fused clustering loss: scale * KL(p || softassign(z, mu)) + gradients from one launch (models.py:1124-1127)
    @custom_op("scc_b200::dec_kl_loss", mutates_args=(), device_types="cuda")
    def _dec_kl_loss(z: torch.Tensor, mu: torch.Tensor, p: torch.Tensor, alpha: float, scale: float) -> tuple[
            torch.Tensor, torch.Tensor, torch.Tensor]:
        if z.dtype == torch.float64:
            stats, dz, _ = dec_grad_f64(z, mu, alpha, p=p, scale=scale)
            return stats[0].clone(), dz, stats[2:].view(mu.shape).clone()
        stats, dz = dec_kl_grad(z, mu, alpha, p=p, scale=scale)
        return stats[0].to(torch.float32), dz, stats[2:].view(mu.shape).to(torch.float32)

    @_dec_kl_loss.register_fake
    def _(z, mu, p, alpha, scale):
        return z.new_empty(()), torch.empty_like(z), torch.empty_like(mu)

    def _kl_setup(ctx, inputs, output):
        ctx.save_for_backward(output[1], output[2])

    def _kl_backward(ctx, g_loss, g_dz, g_dmu):
        dz, dmu = ctx.saved_tensors
        return g_loss * dz, g_loss * dmu, None, None, None

    _dec_kl_loss.register_autograd(_kl_backward, setup_context=_kl_setup)

    # ---- plain launches
    @custom_op("scc_b200::dec_assign", mutates_args=(), device_types="cuda")
    def _dec_assign(z: torch.Tensor, mu: torch.Tensor, alpha: float, round_decimals: int) -> tuple[
            torch.Tensor, torch.Tensor, torch.Tensor]:
        q, labels, stats = dec_assign(z, mu, alpha, round_decimals)
        return q, labels, stats

    @_dec_assign.register_fake
    def _(z, mu, alpha, round_decimals):
        n, K = z.shape[0], mu.shape[0]
        return (z.new_empty(n, K), z.new_empty(n, dtype=torch.int32), z.new_empty(K + 1, dtype=torch.float64))

    @custom_op("scc_b200::dec_target", mutates_args=(), device_types="cuda")
    def _dec_target(q: torch.Tensor, f: torch.Tensor, round_decimals: int) -> torch.Tensor:
        return dec_target(q, f, round_decimals)

    @_dec_target.register_fake
    def _(q, f, round_decimals):
        return torch.empty_like(q)

    @custom_op("scc_b200::dec_kl_grad", mutates_args=(), device_types="cuda")
    def _dec_kl_grad(z: torch.Tensor, mu: torch.Tensor, p: torch.Tensor, alpha: float, scale: float) -> tuple[
            torch.Tensor, torch.Tensor]:
        stats, dz = dec_kl_grad(z, mu, alpha, p=p, scale=scale)
        return stats, dz

    @_dec_kl_grad.register_fake
    def _(z, mu, p, alpha, scale):
        return z.new_empty(mu.numel() + 2, dtype=torch.float64), torch.empty_like(z)

    @custom_op("scc_b200::dec_target_kl_grad", mutates_args=(), device_types="cuda")
    def _dec_target_kl_grad(z: torch.Tensor, mu: torch.Tensor, f: torch.Tensor, alpha: float, round_decimals: int,
                            scale: float) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        stats, p, dz = dec_target_kl_grad(z, mu, f, alpha, round_decimals, scale)
        return stats, p, dz

    @_dec_target_kl_grad.register_fake
    def _(z, mu, f, alpha, round_decimals, scale):
        return (z.new_empty(mu.numel() + 2, dtype=torch.float64), z.new_empty(z.shape[0], mu.shape[0]),
                torch.empty_like(z))

    @custom_op("scc_b200::dec_step", mutates_args=(), device_types="cuda")
    def _dec_step(z: torch.Tensor, mu: torch.Tensor, alpha: float, round_decimals: int, scale: float) -> tuple[
            torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
        r = dec_step(z, mu, alpha, round_decimals, scale)
        return r["q"], r["labels"], r["f"], r["p"], r["dz"], r["stats"]

    @_dec_step.register_fake
    def _(z, mu, alpha, round_decimals, scale):
        n, K = z.shape[0], mu.shape[0]
        return (z.new_empty(n, K), z.new_empty(n, dtype=torch.int32), z.new_empty(K + 1, dtype=torch.float64),
                z.new_empty(n, K), torch.empty_like(z), z.new_empty(mu.numel() + 2, dtype=torch.float64))

    @custom_op("scc_b200::dec_backward", mutates_args=(), device_types="cuda")
    def _dec_backward(z: torch.Tensor, mu: torch.Tensor, grad_q: torch.Tensor, alpha: float) -> tuple[
            torch.Tensor, torch.Tensor]:
        dz, dmu = dec_backward(z, mu, grad_q, alpha)
        return dz, dmu.clone()

    @_dec_backward.register_fake
    def _(z, mu, grad_q, alpha):
        return torch.empty_like(z), mu.new_empty(mu.shape, dtype=torch.float64)

    @custom_op("scc_b200::gmm_em_step", mutates_args=(), device_types="cuda")
    def _gmm_em_step(z: torch.Tensor, params: torch.Tensor, K: int) -> torch.Tensor:
        return gmm_em_step(z, K, params)

    @_gmm_em_step.register_fake
    def _(z, params, K):
        return z.new_empty(gmm_stat_doubles(K, z.shape[1]), dtype=torch.float64)

    @custom_op("scc_b200::gmm_finalize", mutates_args=("means", "weights", "covariances", "prec_chol", "params", "ctrl"),
               device_types="cuda")
    def _gmm_finalize(stats: torch.Tensor, n_total: float, means: torch.Tensor, weights: torch.Tensor,
                      covariances: torch.Tensor, prec_chol: torch.Tensor, params: torch.Tensor, ctrl: torch.Tensor,
                      reg_covar: float, tol: float) -> None:
        gmm_finalize(stats, n_total, means, weights, covariances, prec_chol, params, ctrl, reg_covar=reg_covar, tol=tol)

    @_gmm_finalize.register_fake
    def _(stats, n_total, means, weights, covariances, prec_chol, params, ctrl, reg_covar, tol):
        return None


_register_custom_ops()
