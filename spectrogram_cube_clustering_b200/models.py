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

    numpy in -> float64 numpy out, computed in float64 like the reference (the 5-decimal rounding lands on the
    reference's side); a CUDA tensor in -> CUDA tensor out in float64 (float64 in) or float32 (otherwise).
    ``decimals=None`` skips the rounding.
    """
    rd = 0 if decimals is None else int(decimals)
    if isinstance(q, torch.Tensor) and q.is_cuda:
        if q.dtype == torch.float64:
            return ops.dec_target_f64(q.contiguous(), None, rd)[0]
        qd = q.to(torch.float32).contiguous()
        return ops.dec_target(qd, ops.colsum(qd), rd)
    # the reference's contract: float64 numpy in, float64 numpy out — computed in float64 (scc_dec_target_f64)
    qd = torch.as_tensor(np.ascontiguousarray(q, dtype=np.float64)).to(_device())
    return ops.dec_target_f64(qd, None, rd)[0].cpu().numpy()


def batch_eval(dataloader, model, device, mute=True, return_buffer=False, labels_prev=None, group=None):
    """Full-dataset inference (``models.py:41-103``): encoder per batch, latents written straight
    into a device latent buffer; q and labels come from ONE fused assign pass over the buffer
    instead of per-batch layer calls + per-batch D2H copies.

    Returns ``(np.round(q, 5), labels, z)`` as host arrays like the reference, or — with
    ``return_buffer=True`` — ``(LatentBuffer, q_device, labels_device)`` without leaving the GPU.
    ``labels_prev`` (int32 device tensor): the same pass also counts the labels that differ from it
    (``models.py:1098-1099``); the count is left in ``buffer.assign_stats[-1]`` (device, float64).
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
    buf = LatentBuffer(z_dev[:off].contiguous() if off != z_dev.shape[0] else z_dev, group=group)
    if not hasattr(model, "n_clusters"):
        return buf if return_buffer else buf.z.cpu().numpy().astype(np.float64)
    mu = model.clustering.weights.detach().to(device=device, dtype=torch.float32).contiguous()
    if labels_prev is not None:
        buf.labels = labels_prev.to(device=device, dtype=torch.int32).contiguous()
    q, buf.assign_stats = buf.dec_assign(mu, float(model.clustering.alpha), round_decimals=5, want_q=True)
    if return_buffer:
        return buf, q, buf.labels
    return (q.cpu().numpy().astype(np.float64), buf.labels.cpu().numpy().astype(np.int64),
            buf.z.cpu().numpy().astype(np.float64))


# ----------------------------------------------------------------------------- k-means seeding
class KMeans:
    """Lloyd k-means on the device (seeding step of ``gmm``: ``models.py:386-394``; ``kmeans``:
    ``models.py:565-574``).  k-means++ initialisation, ``n_init`` restarts, best inertia wins —
    the scikit-learn semantics the reference relies on; not bit-reproducible against
    scikit-learn's RNG stream (parity is defined from identical initial centres).

    The restarts advance TOGETHER: one launch per Lloyd iteration scans the latent set against
    ``restarts x K`` centres (``scc_kmeans_batch_step``), a second tiny launch moves the centres and decides
    each restart's convergence on the device (``scc_kmeans_batch_update``); the host polls the flags every
    ``poll`` iterations.  On a sharded :class:`LatentBuffer` the packed statistics of all restarts are
    all-reduced once per iteration, the stop threshold comes from all-reduced moments and the k-means++
    picks are global (the owning rank broadcasts the chosen row), so every rank holds identical centres
    and takes identical decisions."""

    def __init__(self, n_clusters, max_iter=1000, n_init=100, random_state=2009, tol=1e-4, restart_block=None):
        self.n_clusters, self.max_iter, self.n_init = int(n_clusters), int(max_iter), int(n_init)
        self.random_state, self.tol = random_state, float(tol)
        self.restart_block = restart_block

    # ---- helpers
    def _block(self, buf: LatentBuffer) -> int:
        """Restarts per launch: bounded by the [R, n] float32 k-means++ distance buffer (<= 1 GiB)."""
        if self.restart_block:
            return max(1, int(self.restart_block))
        return int(max(1, min(self.n_init, (1 << 28) // max(buf.n_local, 1), 128)))

    @staticmethod
    def _shift_tol(buf: LatentBuffer, tol: float) -> float:
        """tol * mean feature variance of the WHOLE latent set (sklearn/cluster/_kmeans.py `_tolerance`),
        from all-reduced moments so that every rank stops at the same iteration."""
        z = buf.z
        d = buf.d
        if z.shape[0]:
            shift = z[0].double()                       # shifted moments: no cancellation in the variance
        else:
            shift = torch.zeros(d, dtype=torch.float64, device=z.device)
        if buf.world > 1:
            torch.distributed.broadcast(shift, src=0, group=buf.group)
        zc = z.double() - shift
        mom = torch.cat([zc.sum(0), (zc * zc).sum(0)])
        buf._allreduce(mom)
        n = max(buf.n_total, 1)
        mean = mom[:d] / n
        var = mom[d:] / n - mean * mean
        return float(tol * var.clamp_min(0).mean().item())

    def _pick_rows(self, buf: LatentBuffer, weights: torch.Tensor | None, R: int, gen: torch.Generator):
        """One global draw per restart: row i of the WHOLE latent set with probability weights[r, i] / sum
        (uniform when weights is None).  Returns the chosen rows [R, d], identical on every rank.
        Sampling is an inverse-CDF lookup (cumulative sum + searchsorted), so sets beyond the 2^24
        categories ``torch.multinomial`` accepts work."""
        z, dev = buf.z, buf.z.device
        u = torch.rand(R, generator=gen, dtype=torch.float64).to(dev)         # CPU generator: same stream on every rank
        if weights is None:
            idx_global = (u * buf.n_total).long().clamp_(max=buf.n_total - 1)
            lo = buf.row_offset
            local = idx_global - lo
            mine = (local >= 0) & (local < buf.n_local)
        else:
            cdf = weights.double().clamp_min(0).cumsum(dim=1)                 # [R, n_local]
            mass = cdf[:, -1].clone() if buf.n_local else torch.zeros(R, dtype=torch.float64, device=dev)
            if buf.world > 1:
                all_mass = torch.zeros(buf.world, R, dtype=torch.float64, device=dev)
                all_mass[buf.rank] = mass
                buf._allreduce(all_mass)
            else:
                all_mass = mass[None]
            total = all_mass.sum(0)
            before = all_mass[:buf.rank].sum(0)
            target = u * total.clamp_min(1e-300)
            mine = (target >= before) & ((target < before + mass) | (buf.rank == buf.world - 1))
            mine &= mass > 0
            if buf.n_local:
                local = torch.searchsorted(cdf, (target - before).clamp_min(0)[:, None]).squeeze(1)
                local = local.clamp_(max=buf.n_local - 1)
            else:
                local = torch.zeros(R, dtype=torch.long, device=dev)
            degenerate = total <= 0                                            # all points coincide with the centres
            if degenerate.any():
                mine = torch.where(degenerate, torch.full_like(mine, buf.rank == 0 and buf.n_local > 0), mine)
                local = torch.where(degenerate, torch.zeros_like(local), local)
        rows = torch.zeros(R, buf.d, dtype=torch.float32, device=dev)
        if buf.n_local:
            rows = torch.where(mine[:, None], z[local.clamp(0, buf.n_local - 1)], rows)
        if buf.world > 1:                                                      # exactly one owner per restart
            rows64 = rows.double()
            buf._allreduce(rows64)
            rows = rows64.float()
        return rows

    def _plusplus(self, buf: LatentBuffer, R: int, gen: torch.Generator) -> torch.Tensor:
        """k-means++ seeding of R restarts at once -> centers [R, K, d]."""
        K, dev = self.n_clusters, buf.z.device
        centers = self._pick_rows(buf, None, R, gen)[:, None, :].repeat(1, K, 1).contiguous()
        mind = torch.empty(R, buf.n_local, dtype=torch.float32, device=dev)
        for k in range(1, K):
            # squared distance to the nearest of the k centres chosen so far (the other rows duplicate centre 0)
            ops.kmeans_batch_step(buf.z, centers, mindist=mind)
            centers[:, k] = self._pick_rows(buf, mind, R, gen)
        return centers

    def _lloyd_batch(self, buf: LatentBuffer, centers: torch.Tensor, shift_tol: float, poll: int = 8):
        """Iterate R restarts from the given centres [R, K, d] (modified in place).
        Returns (inertia [R] float64 of the final centres, n_iter [R] int32)."""
        dev = buf.z.device
        R, K, d = centers.shape
        stats = torch.zeros(R, K * d + 2 + K, dtype=torch.float64, device=dev)
        done = torch.zeros(R, dtype=torch.uint8, device=dev)
        n_iter = torch.zeros(R, dtype=torch.int32, device=dev)
        it = 0
        while it < self.max_iter:
            ops.kmeans_batch_step(buf.z, centers, done=done, out_stats=stats)
            buf._allreduce(stats)
            ops.kmeans_batch_update(centers, stats, shift_tol, done, n_iter=n_iter)
            it += 1
            if it % poll == 0 and bool(done.all().item()):
                break
        ops.kmeans_batch_step(buf.z, centers, out_stats=stats)            # inertia of the final centres
        buf._allreduce(stats)
        return stats[:, 0].clone(), n_iter

    def lloyd(self, buf: LatentBuffer, centers: torch.Tensor, poll: int = 8):
        """Iterate from the given centres [K, d].  Returns (centers, inertia, labels, n_iter)."""
        c = centers.to(device=buf.z.device, dtype=torch.float32).contiguous().clone()[None]
        inertia, n_iter = self._lloyd_batch(buf, c, self._shift_tol(buf, self.tol), poll)
        labels = torch.empty(buf.n_local, dtype=torch.int32, device=buf.z.device)
        ops.kmeans_step(buf.z, c[0], labels=labels)
        return c[0], float(inertia[0].item()), labels, int(n_iter[0].item())

    def fit(self, z, init_centers=None):
        buf = _as_buffer(z)
        dev = buf.z.device
        if buf.n_total < self.n_clusters:
            raise ValueError(f"n_samples={buf.n_total} should be >= n_clusters={self.n_clusters}.")
        gen = torch.Generator(device="cpu")
        gen.manual_seed(0 if self.random_state is None else int(self.random_state))
        shift_tol = self._shift_tol(buf, self.tol)
        if init_centers is not None:
            runs = [torch.as_tensor(np.asarray(init_centers)).to(device=dev, dtype=torch.float32).contiguous().clone()[None]]
        else:
            blk = self._block(buf)
            runs = [self._plusplus(buf, min(blk, self.n_init - r0), gen) for r0 in range(0, self.n_init, blk)]
        best = None
        for centers in runs:
            inertia, n_iter = self._lloyd_batch(buf, centers, shift_tol)
            r = int(torch.argmin(inertia).item())
            if best is None or float(inertia[r]) < best[1]:
                best = (centers[r].clone(), float(inertia[r]), int(n_iter[r]))
        self._centers, self.inertia_, self.n_iter_ = best
        self._labels = torch.empty(buf.n_local, dtype=torch.int32, device=dev)
        ops.kmeans_step(buf.z, self._centers, labels=self._labels)
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
                 group=None, use_graph=True):
        self.n_components, self.max_iter, self.tol, self.reg_covar = int(n_components), int(max_iter), tol, reg_covar
        self.weights_init, self.means_init = weights_init, means_init
        self.covariances_init, self.precisions_init = covariances_init, precisions_init
        self.poll_interval, self.random_state, self.group = int(poll_interval), random_state, group
        self.use_graph, self._graph = bool(use_graph), None
        if n_init != 1:
            raise ValueError("only n_init=1 is supported (the reference uses n_init=1, models.py:406)")

    # -- state helpers
    def _alloc(self, dev, K, d):
        key = (str(dev), K, d)
        if getattr(self, "_alloc_key", None) == key:     # same shapes as the previous fit: keep the state tensors, and
            self._ctrl.zero_()                           # with them the captured EM-iteration graph that points at them
            return
        self._alloc_key, self._graph, self._graph_key = key, None, None
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

    def _em_iteration(self, buf: LatentBuffer):
        """statistics kernel -> ONE tail kernel (fixed-order grid reduction, all-reduce of the packed vector over the
        GPUs, M-step finalisation); convergence is decided on the device (a frozen fit turns both launches into
        no-ops), so nothing here reads a result back."""
        buf.gmm_em_iteration(self.n_components, self._params, self._stats, self._means, self._weights, self._cov,
                             self._pchol, self._ctrl, reg_covar=self.reg_covar, nk_eps=10 * EPS64, tol=self.tol)

    def _capture(self, buf: LatentBuffer):
        """One EM iteration as a CUDA graph (every buffer it touches is persistent): a fit is then
        ``poll_interval`` graph replays per host poll — at C3 on 8 GPUs an iteration is ~0.2 ms of GPU work, the
        same order as four eager launches through Python.  The graph bakes in every argument of the two launches, so
        it is reused by a later fit only for the same buffer object (shard pointer, sizes, workspaces, exchange
        window) and the same tol / reg_covar; anything else recaptures."""
        key = (id(buf), buf.z.data_ptr(), buf.n_local, buf.n_total, float(self.tol), float(self.reg_covar))
        if getattr(self, "_graph", None) is not None and getattr(self, "_graph_key", None) == key \
                and self.use_graph and self._graph_buf is buf:
            return
        self._graph, self._graph_key, self._graph_buf = None, None, None
        if not (self.use_graph and buf.z.is_cuda):
            return
        try:
            side = torch.cuda.Stream(device=buf.z.device)
            side.wait_stream(torch.cuda.current_stream())
            saved = [t.clone() for t in (self._means, self._weights, self._cov, self._pchol, self._params, self._ctrl)]
            with torch.cuda.stream(side):
                self._em_iteration(buf)              # workspaces of the capture stream, NCCL / exchange warm
                side.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self._em_iteration(buf)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(buf.z.device)
            for t, v in zip((self._means, self._weights, self._cov, self._pchol, self._params, self._ctrl), saved):
                t.copy_(v)                           # the warm-up iteration must not count
            self._graph, self._graph_key, self._graph_buf = g, key, buf
        except Exception as exc:  # pragma: no cover - depends on the box
            warnings.warn(f"CUDA graph capture of the EM iteration failed ({exc}); launching eagerly")
            self._graph = None

    # -- public API
    def fit(self, z):
        buf = _as_buffer(z, group=self.group)
        self._initialize(buf)
        self._capture(buf)
        it = 0
        while it < self.max_iter:
            chunk = min(self.poll_interval, self.max_iter - it)
            for _ in range(chunk):                       # no host sync inside: frozen fits no-op on the device
                if self._graph is not None:
                    self._graph.replay()
                else:
                    self._em_iteration(buf)
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
        self._buf = buf                                  # (the captured graph is kept for a later fit of the same buffer)
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

    Known deviation: the initial covariances come from the one-hot responsibilities of the
    nearest ``means_init`` row; scikit-learn takes them from the labels of its own unseeded internal
    ``KMeans(n_init=1)`` before overriding means and weights.  The two coincide when ``means_init`` are converged
    k-means centres of the same data — the only way the reference calls it.
    """
    buf = _as_buffer(z_array)
    if means_init is None:
        km = KMeans(n_clusters=n_clusters, max_iter=1000, n_init=100, random_state=2009).fit(buf)
        counts = torch.bincount(km._labels.long(), minlength=n_clusters).double()
        buf._allreduce(counts)
        counts = counts.cpu().numpy()
        if (counts == 0).any():
            # the reference builds weights_init from np.unique(labels) (models.py:396-401): an empty k-means
            # cluster leaves it shorter than n_components and scikit-learn rejects it
            raise ValueError(f"The parameter 'weights' should have the shape of ({n_clusters},), but got "
                             f"({int((counts > 0).sum())},): k-means left a cluster empty")
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


def initialize_clusters(model, dataloader, config, n_clusters=None):
    """Select and perform the cluster initialisation (``models.py:498-543``).

    ``config.init``: ``'load'`` reads ``labels.npy`` / ``centroids.npy`` from ``<dir of config.saved_weights>/../GMM/
    n_clusters=<K>/`` — the files :func:`gmm_fit` (stage 2) wrote (``models.py:523-530``; labels are subset by
    ``config.index_tra`` when the config has one); ``'rand'`` draws random labels and uniform centroids (testing);
    ``'kmeans'`` / ``'gmm'`` run :func:`kmeans` / :func:`gmm` on the latent set of a full ``batch_eval`` — which here
    never leaves the device.  Returns ``(labels [M], centroids [K, d])`` as host arrays like the reference."""
    n_clusters = model.n_clusters if n_clusters is None else n_clusters
    init = config.init
    if init == "load":
        path = os.path.abspath(os.path.join(config.saved_weights, os.pardir))
        path = os.path.join(path, "GMM", f"n_clusters={n_clusters}")
        labels = np.load(os.path.join(path, "labels.npy"))
        index_tra = getattr(config, "index_tra", None)
        if index_tra is not None:
            labels = labels[index_tra]
        centroids = np.load(os.path.join(path, "centroids.npy"))
        return labels, centroids
    if init == "rand":
        m = len(dataloader.dataset)
        d = model.clustering.weights.shape[1]
        return np.random.randint(0, n_clusters, (m,)), np.random.uniform(size=(n_clusters, d))
    if init not in ("kmeans", "gmm"):
        raise ValueError(f"unknown cluster initialisation {init!r} (load | rand | kmeans | gmm)")
    device = getattr(config, "device", None)
    out = batch_eval(dataloader, model, _device(device), return_buffer=True)
    buf = out[0] if isinstance(out, tuple) else out
    return kmeans(buf, n_clusters) if init == "kmeans" else gmm(buf, n_clusters)


def save_dec_params(model, savepath_run, which):
    """``torch.save(model.state_dict(), DEC_Params_{Initial,Final}.pt)`` (``models.py:1009-1012, 1227-1228``) —
    the reference's keys (``encoder.encoder.N.*``, ``decoder.decoder.N.*``, ``clustering.weights``), CPU tensors."""
    fname = os.path.join(savepath_run, f"DEC_Params_{which}.pt")
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, fname)
    return fname


def save_history(history, path):
    """``utils.save_history`` (``utils.py:1158-1178``): first item of the dict is the CSV index column."""
    import csv
    keys = list(history.keys())
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(keys)
        for row in zip(*[history[k] for k in keys]):
            w.writerow(row)
    return path


# ----------------------------------------------------------------------------- distance scans
def _dist_operands(x, y):
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else (
        y.device if isinstance(y, torch.Tensor) and y.is_cuda else _device())
    as_dev = lambda t: (t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t))).to(
        device=dev, dtype=torch.float32).contiguous()
    return as_dev(x), as_dev(y)


def fractional_distance(x, y, f):
    """``utils.fractional_distance`` (``utils.py:866-869``): ``(sum |x - y|^f)^(1/f)`` along axis 1 for one
    centroid x [d] against rows y [n, d] (the reference's call order, ``plotting.py:189``).  numpy in ->
    float64 numpy out; CUDA tensor in -> CUDA float32 out."""
    host = not (isinstance(y, torch.Tensor) and y.is_cuda)
    xc, yc = _dist_operands(x, y)
    out = ops.dec_distances(yc.reshape(-1, yc.shape[-1]), xc.reshape(1, -1), float(f))[:, 0]
    return out.cpu().numpy().astype(np.float64) if host else out


def distance_matrix(x, y, f):
    """``utils.distance_matrix`` (``utils.py:635-643``): all pairwise fractional distances, x [n, d] vs y [m, d]
    -> [n, m]; columns are scanned 16 at a time."""
    host = not (isinstance(x, torch.Tensor) and x.is_cuda)
    xc, yc = _dist_operands(x, y)
    cols = [ops.dec_distances(xc, yc[j:j + ops.MAX_K].contiguous(), float(f)) for j in range(0, yc.shape[0], ops.MAX_K)]
    out = torch.cat(cols, dim=1)
    return out.cpu().numpy().astype(np.float64) if host else out


def measure_class_inertia(data, centroids, n_clusters):
    """``utils.measure_class_inertia`` (``utils.py:1024-1029``): inertia_j = sum_i ||data_i - centroid_j||^2
    over ALL rows of data, for every centroid — one scan of the latent set."""
    xc, cc = _dist_operands(data, centroids)
    dist = ops.dec_distances(xc, cc[:n_clusters].contiguous(), 2.0).double()
    return (dist * dist).sum(0).cpu().numpy()


# ----------------------------------------------------------------------------- DEC training loop
def DEC_training(model, dataloader, optimizer, n_epochs, gamma=1e-3, tol=3e-3, update_interval_cfg=-1,
                 device=None, labels_prev=None, fused_loss=True, log_every=0, config=None, n_clusters=None,
                 group=None):
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

    * with ``config`` (the reference's configuration object: ``init``, ``saved_weights``, ``savepath_run``,
      ``device``) the stage-2 -> stage-3 hand-off is reproduced: ``initialize_clusters`` supplies
      ``labels_prev`` and the centroids, which are copied into ``clustering.weights``; the state dict is saved as
      ``DEC_Params_Initial.pt`` before and ``DEC_Params_Final.pt`` after training, with ``DEC_history.csv`` /
      ``Delta_history.csv`` beside them (``models.py:1000-1012, 1200-1228``).

    * ``group`` (one process per GPU, each with its own shard of the dataset and a replica of the model): the
      column sums f and the label-change count of every refresh are all-reduced (so p is the target distribution
      of the WHOLE set, as in the reference) and the parameter gradients are averaged over the ranks after
      ``backward`` — data-parallel training with a global batch of ``world * B``.

    The stop rule uses the label-change count the refresh pass itself produces (fused into the assign kernel):
    one scalar read per refresh.  The dataloader must not shuffle (p is indexed by running offset, as in the
    reference).  Returns a history dict (per-epoch MSE / KLD / loss, deltas, whether the stop rule fired).
    """
    from .networks import dec_kl_loss
    device = _device(device)
    bsz = dataloader.batch_size
    M = len(dataloader.dataset)
    upd = int(np.ceil(M / (bsz * 2))) if update_interval_cfg == -1 else int(np.ceil(M / (bsz * update_interval_cfg)))
    mse = torch.nn.MSELoss(reduction="mean")
    kld = torch.nn.KLDivLoss(reduction="sum")
    alpha = float(model.clustering.alpha)

    savepath = getattr(config, "savepath_run", None) if config is not None else None
    if config is not None:
        labels_prev, centroids = initialize_clusters(model, dataloader, config, n_clusters=n_clusters)
        with torch.no_grad():
            model.clustering.weights.copy_(torch.as_tensor(np.asarray(centroids)).to(
                device=model.clustering.weights.device, dtype=model.clustering.weights.dtype))
        if savepath:
            save_dec_params(model, savepath, "Initial")

    def refresh(prev):
        """batch_eval + target_distribution; f and the label-change count come from the assign pass itself."""
        buf, q, labels = batch_eval(dataloader, model, device, return_buffer=True, labels_prev=prev, group=group)
        p = ops.dec_target(q, buf.assign_stats, 5)
        return p, labels, buf.assign_stats, buf.n_total

    world = 1 if group is None else torch.distributed.get_world_size(group)
    params = [q_ for q_ in model.parameters() if q_.requires_grad]

    def average_gradients():
        flat = torch.cat([(q_.grad if q_.grad is not None else torch.zeros_like(q_)).reshape(-1) for q_ in params])
        torch.distributed.all_reduce(flat, group=group)
        flat /= world
        off = 0
        for q_ in params:
            if q_.grad is not None:
                q_.grad.copy_(flat[off:off + q_.numel()].view_as(q_))
            off += q_.numel()

    p, labels, _, n_all = refresh(None)
    if labels_prev is None:
        labels_prev = labels.clone()
    else:
        labels_prev = torch.as_tensor(np.asarray(labels_prev) if not isinstance(labels_prev, torch.Tensor)
                                      else labels_prev).to(device=device, dtype=torch.int32).contiguous()
    hist = dict(mse=[], kld=[], loss=[], deltas=[], deltas_iter=[], finished=False, update_interval=upd)
    n_iter = 1
    finished = False
    for epoch in range(n_epochs):
        sums = torch.zeros(3, dtype=torch.float64, device=device)
        running = 0
        for batch_num, batch in enumerate(dataloader):
            x = batch[0] if isinstance(batch, (list, tuple)) else batch
            x = x.to(device, non_blocking=True)
            if (batch_num % upd == 0) and not (batch_num == 0 and epoch == 0):
                p, labels, st, n_all = refresh(labels_prev)
                delta = float(st[-1].item()) / n_all                 # fused count: models.py:1098-1099
                hist["deltas"].append(delta); hist["deltas_iter"].append(n_iter)
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
            if world > 1:
                average_gradients()
            optimizer.step()
            running += B
            n_iter += 1
            sums += torch.stack([loss_rec.detach(), loss_clust.detach(), loss.detach()]).double() * B
        if running:
            tot = (sums / running).cpu().numpy()
            hist["mse"].append(float(tot[0])); hist["kld"].append(float(tot[1])); hist["loss"].append(float(tot[2]))
        if log_every and (epoch + 1) % log_every == 0:
            print(f"epoch {epoch + 1}/{n_epochs}: " + ", ".join(f"{k}={hist[k][-1]:.4e}" for k in ("mse", "kld", "loss")))
        if finished:
            break
    hist["finished"] = finished
    if savepath:
        epochs = list(range(1, len(hist["loss"]) + 1))
        save_history({"Epoch": epochs, "Reconstruction Loss": hist["mse"], "Clustering Loss": hist["kld"],
                      "Total Loss": hist["loss"]}, os.path.join(savepath, "DEC_history.csv"))
        save_history({"Iteration": hist["deltas_iter"], "Delta": hist["deltas"]},
                     os.path.join(savepath, "Delta_history.csv"))
        hist["params_final"] = save_dec_params(model, savepath, "Final")
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
