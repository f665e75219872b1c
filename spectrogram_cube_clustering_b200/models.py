"""Host-side mirror of the clustering functions of ``Cluster/models.py``.

Same names, argument meaning and error behaviour as the reference for the hot
path — ``target_distribution`` (``models.py:1302-1322``), ``gmm``
(``models.py:365-413``), ``kmeans`` (``models.py:546-574``), ``batch_eval``
(``models.py:41-103``) — but every N-sized computation runs in the sm_100a
kernels over a device-resident :class:`LatentBuffer`; numpy arrays only cross
the boundary where the reference API hands them to the caller.

``GaussianMixture`` is the scikit-learn-compatible front end of the fused EM
(the reference's stage 2 is ``sklearn.mixture.GaussianMixture.fit_predict``):
the whole fit loop stays on the device — statistics kernel -> (all-reduce) ->
finalize kernel, convergence decided on the device — and the host polls the
control block every ``poll_interval`` iterations.
"""
from __future__ import annotations

import math
import os
import warnings

import numpy as np
import torch

from . import ops
from ._lib import SccError
from .latent_buffer import LatentBuffer

EPS64 = float(np.finfo(np.float64).eps)


class ConvergenceWarning(UserWarning):
    """EM hit ``max_iter`` before |delta lower bound| < tol (sklearn raises the same warning,
    sklearn/mixture/_base.py:292-300)."""


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise SccError("no CUDA device: the B200 clustering path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_buffer(z, device=None, group=None) -> LatentBuffer:
    if isinstance(z, LatentBuffer):
        return z
    if isinstance(z, torch.Tensor) and z.is_cuda:
        return LatentBuffer(z, group=group)
    return LatentBuffer.from_host(np.ascontiguousarray(z), _device(device), group=None)


# ----------------------------------------------------------------------------- DEC
def target_distribution(q, decimals=5):
    """p = normalise_rows(q**2 / q.sum(0)), rounded to 5 decimals (``models.py:1320-1322``).

    numpy in -> float64 numpy out (the reference contract); a CUDA tensor in -> CUDA float32 out.
    ``decimals=None`` skips the rounding.
    """
    rd = 0 if decimals is None else int(decimals)
    if isinstance(q, torch.Tensor) and q.is_cuda:
        qd = q.to(torch.float32).contiguous()
        return ops.dec_target(qd, ops.colsum(qd), rd)
    qd = torch.as_tensor(np.ascontiguousarray(q, dtype=np.float32)).to(_device())
    p = ops.dec_target(qd, ops.colsum(qd), rd)
    return p.cpu().numpy().astype(np.float64)


def batch_eval(dataloader, model, device, mute=True, return_buffer=False):
    """Full-dataset inference (``models.py:41-103``): encoder per batch, latents written straight
    into a device latent buffer; q and labels come from ONE fused assign pass over the buffer
    instead of per-batch layer calls + per-batch D2H copies.

    Returns ``(np.round(q, 5), labels, z)`` as host arrays like the reference, or — with
    ``return_buffer=True`` — ``(LatentBuffer, q_device, labels_device)`` without leaving the GPU.
    """
    model.eval()
    device = torch.device(device)
    n = len(dataloader.dataset)
    d = model.clustering.weights.shape[1] if hasattr(model, "clustering") else None
    z_dev = None
    off = 0
    with torch.no_grad():
        for batch in dataloader:
            x = batch[0] if isinstance(batch, (list, tuple)) else batch
            x = x.to(device, non_blocking=True)
            if x.dim() == 5:                       # Zarr loader batches, as models.py:147-148 reshapes them
                x = x.reshape(-1, *x.shape[2:])
            z = model.encoder(x.to(next(model.encoder.parameters()).dtype))
            if z_dev is None:
                d = z.shape[1]
                z_dev = torch.empty(max(n, z.shape[0]), d, dtype=torch.float32, device=device)
            if off + z.shape[0] > z_dev.shape[0]:  # dataset items that expand to several rows
                z_dev = torch.cat([z_dev, torch.empty_like(z_dev)])
            z_dev[off:off + z.shape[0]] = z.to(torch.float32)
            off += z.shape[0]
    buf = LatentBuffer(z_dev[:off].contiguous() if off != z_dev.shape[0] else z_dev)
    if not hasattr(model, "n_clusters"):
        return buf if return_buffer else buf.z.cpu().numpy().astype(np.float64)
    mu = model.clustering.weights.detach().to(device=device, dtype=torch.float32).contiguous()
    q, _ = buf.dec_assign(mu, float(model.clustering.alpha), round_decimals=5, want_q=True)
    if return_buffer:
        return buf, q, buf.labels
    return (q.cpu().numpy().astype(np.float64), buf.labels.cpu().numpy().astype(np.int64),
            buf.z.cpu().numpy().astype(np.float64))


# ----------------------------------------------------------------------------- k-means seeding
class KMeans:
    """Lloyd k-means on the device (seeding step of ``gmm``: ``models.py:386-394``; ``kmeans``:
    ``models.py:565-574``).  k-means++ initialisation, ``n_init`` restarts, best inertia wins —
    the scikit-learn semantics the reference relies on; not bit-reproducible against
    scikit-learn's RNG stream (parity is defined from identical initial centres)."""

    def __init__(self, n_clusters, max_iter=1000, n_init=100, random_state=2009, tol=1e-4):
        self.n_clusters, self.max_iter, self.n_init = int(n_clusters), int(max_iter), int(n_init)
        self.random_state, self.tol = random_state, float(tol)

    def _plusplus(self, buf: LatentBuffer, gen: torch.Generator) -> torch.Tensor:
        z, K = buf.z, self.n_clusters
        n = z.shape[0]
        centers = z[torch.randint(n, (1,), generator=gen, device=z.device)].repeat(K, 1).contiguous()
        mind = torch.empty(n, dtype=torch.float32, device=z.device)
        for k in range(1, K):
            # distances to the k centres chosen so far (the remaining rows duplicate centre 0)
            ops.kmeans_step(z, centers, mindist=mind)
            idx = torch.multinomial(mind.clamp_min(0) + 1e-30, 1, generator=gen)
            centers[k] = z[idx[0]]
        return centers

    def lloyd(self, buf: LatentBuffer, centers: torch.Tensor, poll: int = 8):
        """Iterate from the given centres.  Returns (centers, inertia, labels, n_iter)."""
        z, K, d = buf.z, self.n_clusters, buf.d
        centers = centers.to(torch.float32).contiguous().clone()
        stats = torch.empty(K * d + 2 + K, dtype=torch.float64, device=z.device)
        var = float(z.var(dim=0).mean().item()) if z.shape[0] > 1 else 1.0
        thresh = self.tol * var
        n_iter = 0
        shifts = []
        while n_iter < self.max_iter:
            ops.kmeans_step(z, centers, out_stats=stats)
            buf._allreduce(stats)
            cnt = stats[2 + K * d:].view(K, 1)
            step = torch.where(cnt > 0, stats[2:2 + K * d].view(K, d) / cnt.clamp_min(1.0), torch.zeros_like(cnt))
            centers += step.to(torch.float32)
            shifts.append((step * step).sum())
            n_iter += 1
            if n_iter % poll == 0 or n_iter == self.max_iter:
                if float(shifts[-1].item()) <= thresh:        # sklearn: total centre shift <= tol * var
                    break
        labels = torch.empty(z.shape[0], dtype=torch.int32, device=z.device)
        ops.kmeans_step(z, centers, labels=labels, out_stats=stats)
        buf._allreduce(stats)
        return centers, float(stats[0].item()), labels, n_iter

    def fit(self, z, init_centers=None):
        buf = _as_buffer(z)
        gen = torch.Generator(device=buf.z.device)
        gen.manual_seed(0 if self.random_state is None else int(self.random_state))
        best = None
        runs = 1 if init_centers is not None else self.n_init
        for _ in range(runs):
            c0 = torch.as_tensor(init_centers, device=buf.z.device) if init_centers is not None \
                else self._plusplus(buf, gen)
            res = self.lloyd(buf, c0)
            if best is None or res[1] < best[1]:
                best = res
        self._centers, self.inertia_, self._labels, self.n_iter_ = best
        self.cluster_centers_ = self._centers.cpu().numpy().astype(np.float64)
        self.labels_ = self._labels.cpu().numpy().astype(np.int64)
        return self

    def fit_predict(self, z, init_centers=None):
        return self.fit(z, init_centers).labels_


def kmeans(z_array, n_clusters):
    """``models.py:546-574``: KMeans(n_init=100, max_iter=1000, random_state=2009)."""
    km = KMeans(n_clusters=n_clusters, max_iter=1000, n_init=100, random_state=2009)
    km.fit_predict(z_array)
    return km.labels_, km.cluster_centers_


# ----------------------------------------------------------------------------- GMM
class GaussianMixture:
    """Full-covariance EM on the device, scikit-learn's attribute names and stop rule.

    Parameters follow ``sklearn.mixture.GaussianMixture`` as the reference uses it
    (``models.py:403-409``): ``n_components``, ``max_iter``, ``tol=1e-3``, ``reg_covar=1e-6``,
    ``weights_init``, ``means_init``, plus ``covariances_init`` / ``precisions_init`` for a fully
    explicit start.  When no covariance is given the initial covariances come from one-hot
    responsibilities of the nearest initial mean — what scikit-learn derives from its internal
    k-means labels (sklearn/mixture/_base.py:119-128) — with weights/means then replaced by the
    ``*_init`` values (sklearn/mixture/_gaussian_mixture.py:848-881).
    """

    def __init__(self, n_components, max_iter=100, tol=1e-3, reg_covar=1e-6, weights_init=None, means_init=None,
                 covariances_init=None, precisions_init=None, n_init=1, poll_interval=10, random_state=None,
                 group=None):
        self.n_components, self.max_iter, self.tol, self.reg_covar = int(n_components), int(max_iter), tol, reg_covar
        self.weights_init, self.means_init = weights_init, means_init
        self.covariances_init, self.precisions_init = covariances_init, precisions_init
        self.poll_interval, self.random_state, self.group = int(poll_interval), random_state, group
        if n_init != 1:
            raise ValueError("only n_init=1 is supported (the reference uses n_init=1, models.py:406)")

    # -- state helpers
    def _alloc(self, dev, K, d):
        f64 = dict(dtype=torch.float64, device=dev)
        self._means = torch.empty(K, d, **f64)
        self._weights = torch.empty(K, **f64)
        self._cov = torch.empty(K, d, d, **f64)
        self._pchol = torch.empty(K, d, d, **f64)
        self._params = torch.empty(ops.gmm_param_floats(K, d), dtype=torch.float32, device=dev)
        self._ctrl = torch.zeros(8, **f64)
        self._stats = torch.empty(ops.gmm_stat_doubles(K, d), **f64)

    def _initialize(self, buf: LatentBuffer):
        K, d, dev = self.n_components, buf.d, buf.z.device
        if buf.n_total < K:
            raise ValueError(f"Expected n_samples >= n_components but got n_components = {K}, "
                             f"n_samples = {buf.n_total}")
        if not ops.gmm_supported(d, K):
            raise SccError(f"GMM kernels are not instantiated for d={d}, K={K}")
        self._alloc(dev, K, d)
        f64 = dict(dtype=torch.float64, device=dev)
        w = torch.full((K,), 1.0 / K, **f64) if self.weights_init is None else torch.as_tensor(
            np.asarray(self.weights_init, dtype=np.float64)).to(dev)
        if self.means_init is None:
            km = KMeans(K, max_iter=300, n_init=1, random_state=self.random_state).fit(buf)
            mu = km._centers.to(torch.float64)
        else:
            mu = torch.as_tensor(np.asarray(self.means_init, dtype=np.float64)).to(dev)
        if self.precisions_init is not None:
            cov = torch.linalg.inv(torch.as_tensor(np.asarray(self.precisions_init, dtype=np.float64)).to(dev))
        elif self.covariances_init is not None:
            cov = torch.as_tensor(np.asarray(self.covariances_init, dtype=np.float64)).to(dev)
        else:
            # one-hot responsibilities of the nearest mean -> initial covariances (means/weights kept)
            eye = torch.eye(d, **f64).expand(K, d, d).contiguous()
            self._means.copy_(mu)
            ops.gmm_pack_params(torch.full((K,), 1.0 / K, **f64), self._means, eye, self._params, self._pchol,
                                self._ctrl)
            buf.gmm_em_pass(K, self._params, self._stats, mode=ops.GMM_HARD)
            ops.gmm_finalize(self._stats, buf.n_total, self._means, self._weights, self._cov, self._pchol,
                             self._params, self._ctrl, reg_covar=self.reg_covar, nk_eps=10 * EPS64, tol=0.0)
            cov = self._cov.clone()
        self._means.copy_(mu); self._weights.copy_(w); self._cov.copy_(cov)
        ops.gmm_pack_params(self._weights, self._means, self._cov, self._params, self._pchol, self._ctrl)
        self._check_pd()

    def _check_pd(self):
        bad = float(self._ctrl[4].item())
        if bad:
            raise ValueError(
                "Fitting the mixture model failed because some components have ill-defined empirical covariance "
                f"(component {int(bad) - 1} is not positive definite). Try to decrease the number of components, "
                "increase reg_covar, or scale the input data.")

    # -- public API
    def fit(self, z):
        buf = _as_buffer(z, group=self.group)
        self._initialize(buf)
        K = self.n_components
        it = 0
        while it < self.max_iter:
            chunk = min(self.poll_interval, self.max_iter - it)
            for _ in range(chunk):                       # no host sync inside: frozen fits no-op on the device
                buf.gmm_em_pass(K, self._params, self._stats, ctrl=self._ctrl)
                ops.gmm_finalize(self._stats, buf.n_total, self._means, self._weights, self._cov, self._pchol,
                                 self._params, self._ctrl, reg_covar=self.reg_covar, nk_eps=10 * EPS64,
                                 tol=self.tol)
            it += chunk
            ctrl = self._ctrl.cpu().numpy()              # one poll per chunk
            if ctrl[4]:
                self._check_pd()
            if ctrl[5]:
                break
        ctrl = self._ctrl.cpu().numpy()
        self.n_iter_ = int(ctrl[2])
        self.converged_ = bool(ctrl[3])
        self.lower_bound_ = float(ctrl[0])
        if not self.converged_ and self.max_iter > 0:
            warnings.warn("Best performing initialization did not converge. Try different init parameters, or "
                          "increase max_iter, tol, or check for degenerate data.", ConvergenceWarning)
        self.weights_ = self._weights.cpu().numpy()
        self.means_ = self._means.cpu().numpy()
        self.covariances_ = self._cov.cpu().numpy()
        self.precisions_cholesky_ = self._pchol.cpu().numpy()
        self._buf = buf
        return self

    def predict_device(self, buf: LatentBuffer | None = None) -> torch.Tensor:
        buf = self._buf if buf is None else buf
        labels = torch.empty(buf.n_local, dtype=torch.int32, device=buf.z.device)
        buf.gmm_em_pass(self.n_components, self._params, self._stats, mode=ops.GMM_ESTEP_ONLY, labels=labels)
        return labels

    def predict(self, z=None):
        buf = self._buf if z is None else _as_buffer(z, group=self.group)
        return self.predict_device(buf).cpu().numpy().astype(np.int64)

    def fit_predict(self, z):
        """Final E-step after the last M-step, argmax of the responsibilities
        (sklearn/mixture/_base.py:307-312)."""
        return self.fit(z).predict()


def gmm(z_array, n_clusters, means_init=None, weights_init=None):
    """Initialise clusters with a Gaussian mixture (``models.py:365-413``).

    KMeans(n_init=100, max_iter=1000, random_state=2009) seeds the means and weights
    (``models.py:386-401``) unless they are supplied; full-covariance EM with ``max_iter=1000``,
    ``tol=1e-3``, ``reg_covar=1e-6`` follows (``models.py:403-411``).  Returns
    ``(labels [M] int64, centroids [K, d] float64)``.
    """
    buf = _as_buffer(z_array)
    if means_init is None:
        km = KMeans(n_clusters=n_clusters, max_iter=1000, n_init=100, random_state=2009).fit(buf)
        counts = np.bincount(km.labels_, minlength=n_clusters).astype(np.float64)
        means_init, weights_init = km.cluster_centers_, counts / buf.n_total
    gm = GaussianMixture(n_components=n_clusters, max_iter=1000, weights_init=weights_init, means_init=means_init)
    with np.errstate(under="ignore"):
        labels = gm.fit_predict(buf)
    return labels, gm.means_


def save_labels(label_list, savepath, serial=None):
    """Append sample-wise labels to ``Labels.csv`` / ``Labels{serial}.csv`` (``utils.py:1181-1209``): a list of
    dicts sharing one key set; the header row is written only when the file is created."""
    import csv
    fname = os.path.join(savepath, "Labels.csv" if serial is None else f"Labels{serial}.csv")
    fresh = not os.path.exists(fname)
    with open(fname, "w" if fresh else "a", newline="") as fh:
        writer = csv.DictWriter(fh, list(label_list[0].keys()))
        if fresh:
            writer.writeheader()
        writer.writerows(label_list)
    return fname


def gmm_fit(config, z_array, n_clusters):
    """GMM initialisation plus the reference's on-disk side effects (``models.py:415-449``): ``Labels.csv``
    (columns ``idx,label``), ``labels.npy`` and ``centroids.npy`` in ``config.savepath_run``.
    Returns ``(labels [M], centroids [K, d])`` exactly as :func:`gmm`."""
    labels, centroids = gmm(z_array, n_clusters)
    rows = [{"idx": int(i), "label": int(lab)} for i, lab in enumerate(labels)]
    save_labels(rows, config.savepath_run)
    np.save(os.path.join(config.savepath_run, "labels"), labels)
    np.save(os.path.join(config.savepath_run, "centroids"), centroids)
    return labels, centroids


# ----------------------------------------------------------------------------- DEC training loop
def DEC_training(model, dataloader, optimizer, n_epochs, gamma=1e-3, tol=3e-3, update_interval_cfg=-1,
                 device=None, labels_prev=None, fused_loss=True, log_every=0):
    """DEC fine-tuning loop with the reference's cadence (``models.py:929-1231``) and none of its
    host round trips:

    * ``batch_eval`` fills a device latent buffer; q, labels and the label-change count come from one
      fused pass, ``target_distribution`` runs on the device and p STAYS on the device (the reference
      copies q to numpy, computes p there and re-uploads it slice by slice: ``models.py:89,1113-1114``);
    * every ``update_interval = ceil(M / (2B))`` batches p is refreshed and the stop rule
      ``delta_label < tol`` is evaluated (``models.py:985-989, 1093-1111``);
    * per batch: ``loss = MSE(x_rec, x) + gamma * KL(p_batch || q) / B``, backward, optimizer step
      (``models.py:1121-1128``).  ``fused_loss=True`` takes loss + dL/dz + dL/dmu from one
      ``dec_kl_grad`` launch; ``False`` runs the literal ``KLDivLoss(log q, p)`` line through autograd;
    * losses are accumulated on the device and read back once per epoch (the reference syncs three
      scalars per batch: ``models.py:1131-1133``).

    The dataloader must not shuffle (p is indexed by running offset, as in the reference).
    Returns a history dict (per-epoch MSE / KLD / loss, deltas, whether the stop rule fired).
    """
    from .networks import dec_kl_loss
    device = _device(device)
    bsz = dataloader.batch_size
    M = len(dataloader.dataset)
    upd = int(np.ceil(M / (bsz * 2))) if update_interval_cfg == -1 else int(np.ceil(M / (bsz * update_interval_cfg)))
    mse = torch.nn.MSELoss(reduction="mean")
    kld = torch.nn.KLDivLoss(reduction="sum")
    alpha = float(model.clustering.alpha)

    def refresh():
        buf, q, labels = batch_eval(dataloader, model, device, return_buffer=True)
        p = ops.dec_target(q, ops.colsum(q), 5)
        return p, labels

    p, labels = refresh()
    if labels_prev is None:
        labels_prev = labels.clone()
    else:
        labels_prev = torch.as_tensor(labels_prev).to(device=device, dtype=torch.int32)
    hist = dict(mse=[], kld=[], loss=[], deltas=[], finished=False, update_interval=upd)
    finished = False
    for epoch in range(n_epochs):
        sums = torch.zeros(3, dtype=torch.float64, device=device)
        running = 0
        for batch_num, batch in enumerate(dataloader):
            x = batch[0] if isinstance(batch, (list, tuple)) else batch
            x = x.to(device, non_blocking=True)
            if (batch_num % upd == 0) and not (batch_num == 0 and epoch == 0):
                p, labels = refresh()
                delta = float((labels != labels_prev).sum().item()) / labels.shape[0]
                hist["deltas"].append(delta)
                labels_prev = labels.clone()
                if delta < tol:
                    finished = True
                    break
            B = x.shape[0]
            tar = p[running:running + B]
            model.train()
            optimizer.zero_grad(set_to_none=True)
            z = model.encoder(x)
            x_rec = model.decoder(z)
            loss_rec = mse(x_rec, x)
            if fused_loss:
                loss_clust = dec_kl_loss(z, model.clustering.weights, tar, alpha, gamma / B)
            else:
                loss_clust = gamma * kld(torch.log(model.clustering(z)), tar) / B
            loss = loss_rec + loss_clust
            loss.backward()
            optimizer.step()
            running += B
            sums += torch.stack([loss_rec.detach(), loss_clust.detach(), loss.detach()]).double() * B
        if running:
            tot = (sums / running).cpu().numpy()
            hist["mse"].append(float(tot[0])); hist["kld"].append(float(tot[1])); hist["loss"].append(float(tot[2]))
        if log_every and (epoch + 1) % log_every == 0:
            print(f"epoch {epoch + 1}/{n_epochs}: " + ", ".join(f"{k}={hist[k][-1]:.4e}" for k in ("mse", "kld", "loss")))
        if finished:
            break
    hist["finished"] = finished
    return hist


# ----------------------------------------------------------------------------- DEC refinement
def dec_refine(buf: LatentBuffer, centroids, alpha=1.0, gamma=1e-3, lr=1e-3, tol=3e-3, max_steps=1000,
               round_decimals=5, update_every=1, betas=(0.9, 0.999), eps=1e-8):
    """Centroid-only DEC refinement over a device-resident latent set (the N-scaled form of the
    DEC loop, ``models.py:1093-1128`` with the encoder frozen): every step is
    assign -> all-reduce f -> KL gradients -> all-reduce dmu -> Adam on the centroids, all on
    the device; the label-change stop rule (``models.py:1098-1111``) is polled every
    ``update_every`` steps from the fused count.  Returns (centroids, history)."""
    dev = buf.z.device
    mu = torch.as_tensor(centroids).to(device=dev, dtype=torch.float32).contiguous().clone()
    m = torch.zeros_like(mu, dtype=torch.float64)
    v = torch.zeros_like(mu, dtype=torch.float64)
    mu64 = mu.to(torch.float64)
    history = []
    K = mu.shape[0]
    for step in range(1, max_steps + 1):
        res = buf.dec_step(mu, alpha, gamma, round_decimals)
        g = res.dmu
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        mhat = m / (1 - betas[0] ** step)
        vhat = v / (1 - betas[1] ** step)
        mu64 -= lr * mhat / (vhat.sqrt() + eps)
        mu.copy_(mu64)
        if step % update_every == 0:
            delta = float(res.n_changed.item()) / buf.n_total
            history.append(dict(step=step, loss=float(res.loss.item()), delta=delta))
            if step > 1 and delta < tol:
                break
    return mu.cpu().numpy().astype(np.float64), history
