"""Oracle (TEST INFRASTRUCTURE) — DEC clustering layer, float64 numpy.

Restates, op for op, the reference's stage-1 hot path:

* ``soft_assign``          <- ``Cluster/networks.py:279-288`` (ClusteringLayer.forward)
* ``labels_from_q``        <- ``Cluster/models.py:92``        (argmax over the UNROUNDED q)
* ``round_decimals``       <- ``Cluster/models.py:94,1322``   (np.round(.., 5))
* ``target_distribution``  <- ``Cluster/models.py:1320-1322``
* ``kl_loss``              <- ``Cluster/models.py:1124-1125`` (gamma * KLDivLoss('sum')(log q, p) / B)
* ``kl_grads`` / ``backward_generic`` <- the autograd graph implied by
  ``Cluster/models.py:1124-1127`` (the reference has no hand-written backward;
  closed forms derived in SURVEY.md §8 a3 and pinned against torch autograd of
  the reference layer in ``tests/golden/dec_*.npz``).
* ``delta_label``          <- ``Cluster/models.py:1098-1099``

Pinned by tests/test_oracle_golden.py against fixtures produced by the
reference code itself (oracle/make_golden.py).  Not imported by the product.
"""
from __future__ import annotations

import numpy as np


def soft_assign(z, mu, alpha=1.0):
    """q_ij of the Student's-t kernel.  networks.py:279-288, same op order."""
    z = np.asarray(z, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    x = z[:, None, :] - mu[None, :, :]          # :280  [N,K,d]
    x = x * x                                    # :281
    x = x.sum(axis=2)                            # :282  ||z_i - mu_j||^2
    x = 1.0 + (x / alpha)                        # :283
    x = 1.0 / x                                  # :284
    x = x ** ((alpha + 1.0) / 2.0)               # :285
    x = x.T / x.sum(axis=1)                      # :286
    return np.ascontiguousarray(x.T)             # :287


def labels_from_q(q):
    """models.py:92 — first index wins on ties (numpy argmax)."""
    return np.argmax(np.asarray(q), axis=1)


def round_decimals(x, decimals=5):
    """models.py:94 / models.py:1322 — numpy rounds half to even on x*10^dec."""
    return np.round(x, decimals)


def column_sums(q):
    """f_j = sum_i q_ij — the only global dependency of the DEC path (models.py:1320)."""
    return np.sum(np.asarray(q, dtype=np.float64), axis=0)


def target_distribution(q, round_to=5, f=None):
    """models.py:1320-1322.  ``round_to=None`` gives the unrounded p.  ``f``: column sums of the WHOLE
    q when only a row block is passed (chunked evaluation of large sets); default = np.sum(q, axis=0)."""
    q = np.asarray(q, dtype=np.float64)
    p = q ** 2 / (np.sum(q, axis=0) if f is None else f)         # :1320
    p = np.transpose(np.transpose(p) / np.sum(p, axis=1))        # :1321
    return p if round_to is None else np.round(p, round_to)      # :1322


def kl_loss(q, p, gamma, batch):
    """gamma * KLDivLoss(reduction='sum')(log q, p) / batch  (models.py:1124-1125).

    torch's pointwise term is xlogy(p, p) - p * log q, i.e. 0 where p == 0.
    """
    q = np.asarray(q, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        term = np.where(p > 0, p * (np.log(np.where(p > 0, p, 1.0)) - np.log(q)), 0.0)
    return gamma * term.sum() / batch


def kl_grads(z, mu, p, alpha=1.0, scale=1.0):
    """Closed-form gradients of ``scale * sum_ij p_ij (log p_ij - log q_ij)``.

    With u_ij = 1/(1+d_ij/alpha), s_i = sum_j p_ij  (SURVEY.md §8 a3):
        dL/dz_i  =  scale*(alpha+1)/alpha * sum_j (p_ij - q_ij s_i) u_ij (z_i - mu_j)
        dL/dmu_j = -scale*(alpha+1)/alpha * sum_i (p_ij - q_ij s_i) u_ij (z_i - mu_j)
    Returns (loss, dz [N,d], dmu [K,d]); ``scale`` is gamma / batch.
    """
    z = np.asarray(z, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    diff = z[:, None, :] - mu[None, :, :]
    d2 = (diff * diff).sum(axis=2)
    u = 1.0 / (1.0 + d2 / alpha)
    t = u ** ((alpha + 1.0) / 2.0)
    q = t / t.sum(axis=1, keepdims=True)
    s = p.sum(axis=1, keepdims=True)
    w = (p - q * s) * u * (scale * (alpha + 1.0) / alpha)        # [N,K]
    dz = np.einsum("nk,nkd->nd", w, diff)
    dmu = -np.einsum("nk,nkd->kd", w, diff)
    loss = kl_loss(q, p, scale, 1.0)
    return loss, dz, dmu


def backward_generic(z, mu, grad_q, alpha=1.0):
    """dz, dmu for an arbitrary upstream G = dL/dq (the autograd path of the
    literal reference loop ``metric_kld(torch.log(q), tar_dist)``).

    dL/dlog t_ik = q_ik (G_ik - sum_j G_ij q_ij);  dlog t_ik/dz_i = -(alpha+1)/alpha u_ik (z_i-mu_k).
    """
    z = np.asarray(z, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    g = np.asarray(grad_q, dtype=np.float64)
    diff = z[:, None, :] - mu[None, :, :]
    d2 = (diff * diff).sum(axis=2)
    u = 1.0 / (1.0 + d2 / alpha)
    t = u ** ((alpha + 1.0) / 2.0)
    q = t / t.sum(axis=1, keepdims=True)
    glt = q * (g - (g * q).sum(axis=1, keepdims=True))
    w = -glt * u * ((alpha + 1.0) / alpha)
    dz = np.einsum("nk,nkd->nd", w, diff)
    dmu = -np.einsum("nk,nkd->kd", w, diff)
    return dz, dmu


def delta_label(labels, labels_prev):
    """models.py:1098-1099."""
    labels = np.asarray(labels)
    return np.sum(labels != np.asarray(labels_prev)).astype(np.float32) / labels.shape[0]


def dec_step_chunked(z, mu, alpha=1.0, gamma=1e-3, round_to=5, chunk=100_000):
    """:func:`dec_step` for sets too large for the [N, K, d] temporaries of the reference formulation
    (BASELINE sizes: 1M points): the same functions applied to row blocks, with the one global
    quantity — the column sums f — accumulated over all blocks first.  Same outputs as dec_step."""
    z = np.asarray(z)
    n = z.shape[0]
    blocks = [(s, min(s + chunk, n)) for s in range(0, n, chunk)]
    q = np.concatenate([soft_assign(z[a:b], mu, alpha) for a, b in blocks])
    labels = labels_from_q(q)
    q_r = q if round_to is None else round_decimals(q, round_to)
    f = column_sums(q_r)
    p = np.concatenate([target_distribution(q_r[a:b], round_to, f=f) for a, b in blocks])
    loss, dmu, dz = 0.0, 0.0, []
    for a, b in blocks:
        l, g, m = kl_grads(z[a:b], mu, p[a:b], alpha, gamma / n)
        loss += l; dmu = dmu + m; dz.append(g)
    return dict(q=q, q_rounded=q_r, labels=labels, f=f, p=p, loss=loss, dz=np.concatenate(dz), dmu=dmu)


def dec_step(z, mu, alpha=1.0, gamma=1e-3, round_to=5):
    """One full reference DEC refinement step over a latent set, as DEC_training
    composes it (models.py:1015-1016, 1093-1099, 1121-1127) with batch == N:
    q -> round -> labels -> p -> gamma*KL/N -> grads.
    """
    q = soft_assign(z, mu, alpha)
    labels = labels_from_q(q)
    q_r = q if round_to is None else round_decimals(q, round_to)
    f = column_sums(q_r)
    p = target_distribution(q_r, round_to)
    n = q.shape[0]
    loss, dz, dmu = kl_grads(z, mu, p, alpha, gamma / n)
    return dict(q=q, q_rounded=q_r, labels=labels, f=f, p=p, loss=loss, dz=dz, dmu=dmu)


# --------------------------------------------------------------------------
# Timing mirror: the reference's own op chain in torch (CPU, autograd), used
# by bench.py's cpu_baseline / --impl reference legs.  Same calls, same order
# as networks.py:279-288 + models.py:92-94,1320-1322,1124-1127.
# --------------------------------------------------------------------------
def torch_dec_step(z_t, mu_t, alpha=1.0, gamma=1e-3):
    """z_t [N,d], mu_t [K,d] CPU torch tensors (float64 as the reference runs,
    models.py:965).  Returns (loss, dz, dmu, labels, p)."""
    import torch

    z_t = z_t.detach().clone().requires_grad_(True)
    mu_t = mu_t.detach().clone().requires_grad_(True)

    def layer(x):
        x = x.unsqueeze(1) - mu_t
        x = torch.mul(x, x)
        x = torch.sum(x, dim=2)
        x = 1.0 + (x / alpha)
        x = 1.0 / x
        x = x ** ((alpha + 1.0) / 2.0)
        x = torch.t(x) / torch.sum(x, dim=1)
        return torch.t(x)

    with torch.no_grad():
        q_np = layer(z_t).numpy()                       # batch_eval, models.py:88-90
    labels = np.argmax(q_np, axis=1)                    # models.py:92
    p = target_distribution(np.round(q_np, 5))          # models.py:94,1016
    tar = torch.from_numpy(p)                           # models.py:1114
    q = layer(z_t)                                      # models.py:1122
    loss = gamma * torch.nn.KLDivLoss(reduction="sum")(torch.log(q), tar) / z_t.shape[0]
    loss.backward()                                     # models.py:1127
    return float(loss.detach()), z_t.grad, mu_t.grad, labels, p
