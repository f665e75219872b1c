"""Oracle (TEST INFRASTRUCTURE) — full-covariance Gaussian-mixture EM, float64 numpy.

The reference's stage 2 is one call into scikit-learn
(``Cluster/models.py:403-412``: ``GaussianMixture(n_components=K, max_iter=1000,
n_init=1, weights_init=.., means_init=..).fit_predict(z)``).  scikit-learn is an
un-vendored dependency that the reference leaves UN-PINNED
(``/root/reference/setup.py:33``, ``RISCluster_CPU.yml:19``); the oracle below
restates the algorithm of scikit-learn 1.9.0 (the version in this image), whose
EM arithmetic is unchanged in substance since 0.24:

* ``precision_cholesky``   <- sklearn/mixture/_gaussian_mixture.py:323-385 (full)
* ``log_det_cholesky``     <- sklearn/mixture/_gaussian_mixture.py:448-487
* ``log_gaussian_prob``    <- sklearn/mixture/_gaussian_mixture.py:490-553
* ``e_step``               <- sklearn/mixture/_base.py:314-332, 552-582
* ``m_step``               <- sklearn/mixture/_gaussian_mixture.py:883-901, 282-320, 168-197
* ``fit``                  <- sklearn/mixture/_base.py:203-312 (loop, |delta lower bound| < tol,
                              final E-step, argmax labels)

Pinned by tests/test_oracle_golden.py against sklearn's own private steps run
from identical explicit state (fixtures from oracle/make_golden.py) — the
deterministic harness of SURVEY.md §8c, because ``models.gmm`` itself is
non-deterministic (unseeded internal KMeans).  Not imported by the product.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import linalg as _sla


def precision_cholesky(covariances):
    """P_k = (L_k^-1)^T with L_k = chol(Sigma_k) lower.  Raises ValueError if not PD."""
    covariances = np.asarray(covariances, dtype=np.float64)
    K, d, _ = covariances.shape
    out = np.empty((K, d, d))
    for k in range(K):
        try:
            L = _sla.cholesky(covariances[k], lower=True)
        except _sla.LinAlgError as exc:  # sklearn: _gaussian_mixture.py:343-367
            raise ValueError("ill-defined empirical covariance") from exc
        out[k] = _sla.solve_triangular(L, np.eye(d), lower=True).T
    return out


def log_det_cholesky(pchol):
    K, d, _ = pchol.shape
    return np.sum(np.log(pchol.reshape(K, -1)[:, :: d + 1]), axis=1)


def log_gaussian_prob(X, means, pchol):
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    K = means.shape[0]
    log_det = log_det_cholesky(pchol)
    log_prob = np.empty((n, K))
    for k in range(K):
        y = (X @ pchol[k]) - (means[k] @ pchol[k])           # :526
        log_prob[:, k] = np.sum(np.square(y), axis=1)        # :527
    return -0.5 * (d * math.log(2 * math.pi) + log_prob) + log_det


def _logsumexp(a, axis=1):
    m = np.max(a, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return (np.log(np.sum(np.exp(a - m), axis=axis, keepdims=True)) + m).squeeze(axis)


def e_step(X, weights, means, pchol):
    """Returns (mean_i log p(x_i), log_resp [N,K])."""
    weighted = log_gaussian_prob(X, means, pchol) + np.log(weights)
    log_prob_norm = _logsumexp(weighted, axis=1)
    with np.errstate(under="ignore"):
        log_resp = weighted - log_prob_norm[:, None]
    return float(np.mean(log_prob_norm)), log_resp


def m_step(X, log_resp, reg_covar=1e-6, eps=None):
    """Returns (weights, means, covariances, precisions_cholesky, nk).

    ``eps`` is the machine epsilon added (x10) to N_k — that of resp's dtype in
    sklearn (:312); float64 by default.
    """
    X = np.asarray(X, dtype=np.float64)
    resp = np.exp(log_resp)
    eps = np.finfo(np.float64).eps if eps is None else eps
    nk = resp.sum(axis=0) + 10 * eps                          # :312
    means = (resp.T @ X) / nk[:, None]                        # :313
    K, d = means.shape
    cov = np.empty((K, d, d))
    for k in range(K):                                        # :193-196
        diff = X - means[k]
        cov[k] = ((resp[:, k] * diff.T) @ diff) / nk[k]
        cov[k].flat[:: d + 1] += reg_covar
    weights = nk / nk.sum()                                   # :898
    return weights, means, cov, precision_cholesky(cov), nk


def fit(X, weights, means, covariances, max_iter=100, tol=1e-3, reg_covar=1e-6):
    """EM from explicit state (pi0, mu0, Sigma0).  Mirrors _base.py:262-312.

    Returns dict(weights, means, covariances, precisions_cholesky, lower_bounds,
    n_iter, converged, labels).
    """
    weights = np.asarray(weights, dtype=np.float64)
    means = np.asarray(means, dtype=np.float64)
    cov = np.asarray(covariances, dtype=np.float64)
    pchol = precision_cholesky(cov)
    lower_bound = -np.inf
    history = []
    converged = False
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        prev = lower_bound
        lower_bound, log_resp = e_step(X, weights, means, pchol)
        weights, means, cov, pchol, _ = m_step(X, log_resp, reg_covar)
        history.append(lower_bound)
        if abs(lower_bound - prev) < tol:
            converged = True
            break
    _, log_resp = e_step(X, weights, means, pchol)            # :307-312
    return dict(weights=weights, means=means, covariances=cov, precisions_cholesky=pchol,
                lower_bounds=np.array(history), n_iter=n_iter, converged=converged,
                labels=np.argmax(log_resp, axis=1))
