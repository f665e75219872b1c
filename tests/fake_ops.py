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
