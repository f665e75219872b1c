"""TEST-ONLY stand-ins for spectrogram_cube_clustering_b200.ops, backed by the CPU oracle, so the
multi-rank HOST logic (sharding, packed statistics, all-reduce over gloo) can be exercised
without a GPU.  Never imported by the product."""
import numpy as np
import torch

from oracle import dec as odec
from oracle import gmm as ogmm

GMM_ESTEP_ONLY, GMM_SOFT, GMM_HARD = 0, 1, 2


def dec_assign(z, mu, alpha=1.0, round_decimals=0, want_q=True, want_labels=True, labels_prev=None,
               out_q=None, out_labels=None, out_stats=None):
    zn, mn = z.numpy().astype(np.float64), mu.numpy().astype(np.float64)
    q = odec.soft_assign(zn, mn, alpha)
    labels = odec.labels_from_q(q).astype(np.int32)
    if round_decimals:
        q = np.round(q, round_decimals)
    changed = 0.0 if labels_prev is None else float((labels != labels_prev.numpy()).sum())
    stats = torch.from_numpy(np.concatenate([q.sum(0), [changed]]))
    return (torch.from_numpy(q.astype(np.float32)) if want_q else None,
            torch.from_numpy(labels) if want_labels else None, stats)


def dec_kl_grad(z, mu, alpha=1.0, p=None, f=None, round_decimals=0, scale=1.0, want_dz=True,
                out_dz=None, out_stats=None):
    zn, mn = z.numpy().astype(np.float64), mu.numpy().astype(np.float64)
    K = mn.shape[0]
    if p is None:
        q = odec.soft_assign(zn, mn, alpha)
        if round_decimals:
            q = np.round(q, round_decimals)
        w = q ** 2 / f.numpy()[:K]
        pn = w / w.sum(1, keepdims=True)
        if round_decimals:
            pn = np.round(pn, round_decimals)
    else:
        pn = p.numpy().astype(np.float64)
    loss, dz, dmu = odec.kl_grads(zn, mn, pn, alpha, scale)
    stats = torch.from_numpy(np.concatenate([[loss, pn.sum()], dmu.ravel()]))
    return stats, (torch.from_numpy(dz.astype(np.float32)) if want_dz else None)


def dec_target_kl_grad(z, mu, f, alpha=1.0, round_decimals=0, scale=1.0, want_p=True, want_dz=True,
                       out_p=None, out_dz=None, out_stats=None, pull_f=None, push=None):
    stats, dz = dec_kl_grad(z, mu, alpha, p=None, f=f, round_decimals=round_decimals, scale=scale, want_dz=want_dz)
    zn, mn = z.numpy().astype(np.float64), mu.numpy().astype(np.float64)
    q = odec.soft_assign(zn, mn, alpha)
    if round_decimals:
        q = np.round(q, round_decimals)
    w = q ** 2 / f.numpy()[:mn.shape[0]]
    pn = w / w.sum(1, keepdims=True)
    if round_decimals:
        pn = np.round(pn, round_decimals)
    p = torch.from_numpy(pn.astype(np.float32))
    if out_p is not None:
        out_p.copy_(p)
        p = out_p
    return stats, (p if (want_p or out_p is not None) else None), dz


def _unpack(params, K, d):
    tri = d * (d + 1) // 2
    p = params.numpy().astype(np.float64)
    mu = p[:K * d].reshape(K, d)
    U = np.zeros((K, d, d))
    for k in range(K):
        for b in range(d):
            for a in range(b + 1):
                U[k, a, b] = p[K * d + k * tri + b * (b + 1) // 2 + a]
    cst = p[K * d + K * tri:]
    return mu, U, cst


def gmm_em_step(z, K, params, stats=None, labels=None, resp=None, ctrl=None, mode=GMM_SOFT):
    X = z.numpy().astype(np.float64)
    n, d = X.shape
    mu, U, cst = _unpack(params, K, d)
    lp = np.stack([-0.5 * (((X - mu[k]) @ U[k]) ** 2).sum(1) + cst[k] for k in range(K)], axis=1)
    lse = ogmm._logsumexp(lp, axis=1)
    r = np.exp(lp - lse[:, None])
    if labels is not None:
        labels.copy_(torch.from_numpy(np.argmax(lp, 1).astype(np.int32)))
    if mode == GMM_HARD:
        r = np.eye(K)[np.argmax(lp, 1)]
    tri_idx = [(a, b) for b in range(d) for a in range(b + 1)]
    out = [np.array([lse.sum()]), r.sum(0)]
    s1 = np.stack([(r[:, k:k + 1] * (X - mu[k])).sum(0) for k in range(K)])
    s2 = np.stack([[(r[:, k] * (X[:, a] - mu[k, a]) * (X[:, b] - mu[k, b])).sum() for (a, b) in tri_idx]
                   for k in range(K)])
    res = torch.from_numpy(np.concatenate(out + [s1.ravel(), s2.ravel()]))
    if stats is not None:
        stats.copy_(res)
        return stats
    return res


# ---- GMM state: pack / finalize / whole iteration (the device control block of gmm_api.cu, restated) ----
def gmm_param_floats(K, d):
    return K * d + K * (d * (d + 1) // 2) + K


def gmm_stat_doubles(K, d):
    return 1 + K + K * d + K * (d * (d + 1) // 2)


def gmm_supported(d, K):
    return True


def _pack(weights, means, pchol, params):
    K, d = means.shape
    tri = d * (d + 1) // 2
    U = pchol.numpy()
    flat = [means.numpy().ravel()]
    flat.append(np.concatenate([[U[k][a, b] for b in range(d) for a in range(b + 1)] for k in range(K)]))
    flat.append(ogmm.log_det_cholesky(U) + np.log(weights.numpy()) - 0.5 * d * np.log(2 * np.pi))
    params.copy_(torch.from_numpy(np.concatenate(flat).astype(np.float32)))
    assert params.numel() == K * d + K * tri + K


def _chol_or_bad(cov):
    """U = L^-T per component, or the 1-based index of the first component that is not positive definite."""
    K = cov.shape[0]
    out = np.zeros_like(cov)
    for k in range(K):
        try:
            L = np.linalg.cholesky(cov[k])
        except np.linalg.LinAlgError:
            return None, k + 1
        out[k] = np.linalg.inv(L).T
    return out, 0


def gmm_pack_params(weights, means, covariances, params=None, prec_chol=None, ctrl=None):
    K, d = means.shape
    params = params if params is not None else torch.empty(gmm_param_floats(K, d), dtype=torch.float32)
    prec_chol = prec_chol if prec_chol is not None else torch.empty(K, d, d, dtype=torch.float64)
    ctrl = ctrl if ctrl is not None else torch.zeros(8, dtype=torch.float64)
    U, bad = _chol_or_bad(covariances.numpy())
    ctrl.copy_(torch.tensor([-np.inf, -np.inf, 0.0, 0.0, float(bad), 1.0 if bad else 0.0, 0.0, 0.0], dtype=torch.float64))
    if not bad:
        prec_chol.copy_(torch.from_numpy(U))
        _pack(weights, means, prec_chol, params)
    return params, prec_chol, ctrl


def gmm_finalize(stats, n_total, means, weights, covariances, prec_chol, params, ctrl,
                 reg_covar=1e-6, nk_eps=10 * 2.220446049250313e-16, tol=1e-3):
    """sklearn _m_step from moments centred on the CURRENT means + stop rule, in place (gmm_api.cu)."""
    if ctrl[5] != 0:
        return
    K, d = means.shape
    tri_idx = [(a, b) for b in range(d) for a in range(b + 1)]
    s = stats.numpy()
    nk = s[1:1 + K] + nk_eps
    di = s[1 + K:1 + K + K * d].reshape(K, d) / nk[:, None]
    s2 = s[1 + K + K * d:].reshape(K, len(tri_idx))
    cov = np.zeros((K, d, d))
    for k in range(K):
        for e, (a, b) in enumerate(tri_idx):
            cov[k, a, b] = cov[k, b, a] = s2[k, e] / nk[k] - di[k, a] * di[k, b]
        cov[k] += reg_covar * np.eye(d)
    means.add_(torch.from_numpy(di))
    covariances.copy_(torch.from_numpy(cov))
    U, bad = _chol_or_bad(cov)
    lower, prev = s[0] / n_total, float(ctrl[0])
    ctrl[1] = prev; ctrl[0] = lower; ctrl[2] += 1.0
    if bad:
        ctrl[4] = float(bad); ctrl[5] = 1.0
        return
    weights.copy_(torch.from_numpy(nk / nk.sum()))
    prec_chol.copy_(torch.from_numpy(U))
    _pack(weights, means, prec_chol, params)
    if abs(lower - prev) < tol:
        ctrl[3] = 1.0; ctrl[5] = 1.0


def gmm_em_iteration(z, K, params, stats, n_total, means, weights, covariances, prec_chol, ctrl, mode=GMM_SOFT,
                     reg_covar=1e-6, nk_eps=10 * 2.220446049250313e-16, tol=1e-3, exchange=None):
    if ctrl[5] != 0:                    # frozen fit: both launches are no-ops
        return
    gmm_em_step(z, K, params, stats=stats, mode=mode)
    gmm_finalize(stats, n_total, means, weights, covariances, prec_chol, params, ctrl, reg_covar, nk_eps, tol)


# ---- Lloyd scans (MODE_KMEANS of the gradient kernel + the batched update kernel, restated) ----
def _lloyd_stats(X, C):
    K, d = C.shape
    d2 = ((X[:, None, :] - C[None]) ** 2).sum(-1) if len(X) else np.zeros((0, K))
    lab = d2.argmin(1) if len(X) else np.zeros(0, dtype=np.int64)
    best = d2.min(1) if len(X) else np.zeros(0)
    shift = np.zeros((K, d)); cnt = np.zeros(K)
    for j in range(K):
        m = lab == j
        cnt[j] = m.sum()
        if cnt[j]:
            shift[j] = (X[m] - C[j]).sum(0)
    return np.concatenate([[best.sum(), 0.0], shift.ravel(), cnt]), lab, best


def kmeans_step(z, centers, labels=None, mindist=None, out_stats=None):
    st, lab, best = _lloyd_stats(z.numpy().astype(np.float64), centers.numpy().astype(np.float64))
    if labels is not None:
        labels.copy_(torch.from_numpy(lab.astype(np.int32)))
    if mindist is not None:
        mindist.copy_(torch.from_numpy(best.astype(np.float32)))
    res = torch.from_numpy(st)
    if out_stats is not None:
        out_stats.copy_(res)
        return out_stats
    return res


def kmeans_batch_step(z, centers, done=None, labels=None, mindist=None, out_stats=None):
    R, K, d = centers.shape
    stats = out_stats if out_stats is not None else torch.zeros(R, K * d + 2 + K, dtype=torch.float64)
    X = z.numpy().astype(np.float64)
    for r in range(R):
        if done is not None and done[r]:
            continue                                    # finished restarts are skipped (their stats stay)
        st, lab, best = _lloyd_stats(X, centers[r].numpy().astype(np.float64))
        stats[r] = torch.from_numpy(st)
        if labels is not None:
            labels[r] = torch.from_numpy(lab.astype(np.int32))
        if mindist is not None:
            mindist[r] = torch.from_numpy(best.astype(np.float32))
    return stats


def kmeans_batch_update(centers, stats, shift_tol, done, n_iter=None, inertia=None):
    R, K, d = centers.shape
    for r in range(R):
        if done[r]:
            continue
        st = stats[r].numpy()
        cnt = st[2 + K * d:]
        step = np.where(cnt[:, None] > 0, st[2:2 + K * d].reshape(K, d) / np.maximum(cnt[:, None], 1.0), 0.0)
        centers[r] = (centers[r].double() + torch.from_numpy(step)).float()
        if inertia is not None:
            inertia[r] = st[0]
        if n_iter is not None:
            n_iter[r] += 1
        if (step ** 2).sum() <= shift_tol:
            done[r] = 1
