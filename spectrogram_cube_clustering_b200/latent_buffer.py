"""Sharded latent-buffer manager (new subsystem named by BASELINE.json north_star).

The reference streams every batch's q and z back to host numpy
(``Cluster/models.py:84-90``), computes ``target_distribution`` there and
re-uploads p slice by slice (``models.py:1113-1114``).  Here the N latent points
live in HBM for the whole refinement: rank g of G owns the contiguous row block
``[g*N/G, (g+1)*N/G)`` as float32 ``[n_local, d]``; centroids / mixture
parameters are replicated; every pass is one fused kernel over the shard and the
only thing that crosses GPUs is the packed float64 statistics vector
(f_j; loss + dL/dmu; N_k, sum r z, sum r zz^T), one ``all_reduce(SUM)`` per pass
over NCCL/NVLink.  With ``group=None`` (single process) no collective is issued.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from . import ops


def bind_to_gpu_numa_node(device) -> dict:
    """Pin this process (and so the pinned staging buffers it allocates next: first-touch placement) to the CPU
    cores of the NUMA node the GPU hangs off.  With one process per GPU every rank otherwise inherits the same
    affinity mask and all ranks stage their uploads through one memory controller (round 1: 45 GB/s H2D on one
    GPU, 23.5 GB/s per GPU on eight).  Best effort: returns what it found and did; never raises."""
    import os
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        dev = torch.device(device)
        bus = torch.cuda.get_device_properties(dev).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(dev), "pci_bus_id") else None
        domain = getattr(torch.cuda.get_device_properties(dev), "pci_domain_id", 0)
        devid = getattr(torch.cuda.get_device_properties(dev), "pci_device_id", 0)
        if bus is None:
            return info
        path = f"/sys/bus/pci/devices/{domain:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        with open(path) as fh:
            node = int(fh.read().strip())
        info["numa_node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        info["cpus"] = len(allowed)
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
    except Exception as exc:  # pragma: no cover - depends on the box
        info["error"] = repr(exc)
    return info


def shard_bounds(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous row block of ``rank``: sizes differ by at most one row."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class DecStepResult:
    loss: float | torch.Tensor
    dmu: torch.Tensor          # float64 [K, d], already summed over all shards
    f: torch.Tensor            # float64 [K]
    n_changed: torch.Tensor    # float64 scalar tensor (label changes vs previous pass)
    dz: torch.Tensor | None
    p: torch.Tensor | None = None   # [n_local, K] target distribution of this shard (want_p=True)


class PeerExchange:
    """Latency-optimised all-reduce of the packed float64 statistics over NVLink peer memory
    (``scc_peer_allreduce``): every rank pushes its vector into a slot of every peer's exchange
    window (a ``torch.distributed._symmetric_memory`` allocation), flags it, waits for the world's
    flags and sums the slots in rank order — one small kernel instead of an NCCL launch.
    One process per GPU on one NVSwitch box.  Raises if symmetric memory cannot be set up; the
    caller then stays on NCCL."""

    def __init__(self, group, device, max_len: int):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group, self.max_len = group, int(max_len)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = ops.peer_window_bytes(self.max_len)
        self.window = symm.empty(nbytes, dtype=torch.uint8, device=device)
        self.handle = symm.rendezvous(self.window, group)
        self.window.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        from ._lib import SccExchange
        self.desc = SccExchange(self.ptrs.data_ptr(), self.rank, self.world, self.max_len)   # for the *_ex entry points

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if t.numel() > self.max_len:
            raise ValueError(f"statistics vector of {t.numel()} doubles exceeds the exchange window ({self.max_len})")
        return ops.peer_allreduce(t, self.ptrs, self.rank, self.world, self.max_len)


class LatentBuffer:
    """Device-resident shard of the latent set + the passes that run over it."""

    def __init__(self, z: torch.Tensor, n_total: int | None = None, group=None, exchange: PeerExchange | None = None):
        if z.dim() != 2:
            raise ValueError("z must be [n_local, d]")
        self.z = z.contiguous() if z.dtype == torch.float32 else z.float().contiguous()
        if self.z.data_ptr() % 16:      # a row slice of a larger buffer: the kernels need a 16-byte aligned base
            self.z = self.z.clone()
        self.n_local, self.d = self.z.shape
        self.group = group
        self.world = 1 if group is None else torch.distributed.get_world_size(group)
        self.rank = 0 if group is None else torch.distributed.get_rank(group)
        self.n_total = int(n_total) if n_total is not None else self._sum_int(self.n_local)
        self._row_offset = None     # global index of this shard's first row (lazy: one all-gather)
        self.labels = None          # int32 [n_local] of the last assign pass
        self.assign_stats = None    # float64 [K+1] of the last batch_eval pass (f_j, label changes)
        self._u_scratch = None      # [n_local, K] Student's-t hand-off between the two launches of a tiled-shape step
        self._labels_spare = None
        self.exchange = exchange    # optional NVLink peer-memory exchange (else NCCL / gloo all_reduce)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_host(cls, z_host, device, group=None, pin: bool = True):
        """Shard a host array [N, d] by rows and upload this rank's block."""
        zt = torch.as_tensor(z_host)
        world = 1 if group is None else torch.distributed.get_world_size(group)
        rank = 0 if group is None else torch.distributed.get_rank(group)
        lo, hi = shard_bounds(zt.shape[0], rank, world)
        blk = zt[lo:hi].to(torch.float32).contiguous()
        if pin and torch.cuda.is_available() and torch.device(device).type == "cuda":
            blk = blk.pin_memory()
        return cls(blk.to(device, non_blocking=True), n_total=zt.shape[0], group=group)

    @property
    def row_offset(self) -> int:
        """Global row index of the first local row (rank-ordered contiguous blocks)."""
        if self._row_offset is None:
            if self.group is None or self.world == 1:
                self._row_offset = 0
            else:
                sizes = torch.zeros(self.world, dtype=torch.int64, device=self.z.device)
                sizes[self.rank] = self.n_local
                torch.distributed.all_reduce(sizes, group=self.group)
                self._row_offset = int(sizes[:self.rank].sum().item())
        return self._row_offset

    def _sum_int(self, v: int) -> int:
        if self.group is None:
            return int(v)
        t = torch.tensor([v], dtype=torch.int64, device=self.z.device)
        torch.distributed.all_reduce(t, group=self.group)
        return int(t.item())

    def _allreduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.group is not None and self.world > 1:
            if self.exchange is not None and t.dtype == torch.float64 and t.numel() <= self.exchange.max_len:
                self.exchange.all_reduce(t.view(-1))
            else:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.SUM, group=self.group)
        return t

    def enable_peer_exchange(self, max_clusters: int = 16) -> bool:
        """Switch the statistics all-reduce from NCCL to the one-shot NVLink peer-memory exchange.
        Returns False (and stays on NCCL) if symmetric memory is unavailable."""
        if self.group is None or self.world == 1:
            return False
        d, K = self.d, max_clusters
        max_len = 1 + K + K * d + K * (d * (d + 1) // 2)
        try:
            self.exchange = PeerExchange(self.group, self.z.device, max_len)
            return True
        except Exception as exc:  # pragma: no cover - depends on the box
            import warnings
            warnings.warn(f"peer-memory exchange unavailable, staying on NCCL: {exc}")
            self.exchange = None
            return False

    # ------------------------------------------------------------------ DEC passes
    def dec_assign(self, mu: torch.Tensor, alpha: float = 1.0, round_decimals: int = 0,
                   want_q: bool = False, track_labels: bool = True):
        """Pass 1: q (optional), labels, f_j and the label-change count; f and the count are
        summed over shards.  networks.py:279-288, models.py:92-94,1098-1099,1320."""
        prev = self.labels if track_labels else None
        out_labels = None
        if track_labels:
            out_labels = self._labels_spare if self._labels_spare is not None else torch.empty(
                self.n_local, dtype=torch.int32, device=self.z.device)
        q, labels, stats = ops.dec_assign(self.z, mu, alpha, round_decimals, want_q=want_q,
                                          want_labels=track_labels, labels_prev=prev, out_labels=out_labels)
        if track_labels:
            self._labels_spare, self.labels = self.labels, labels
        self._allreduce(stats)
        return q, stats

    def dec_grad(self, mu: torch.Tensor, f_stats: torch.Tensor, alpha: float = 1.0, gamma: float = 1e-3,
                 round_decimals: int = 0, p: torch.Tensor | None = None, want_dz: bool = False,
                 out_p: torch.Tensor | None = None):
        """Pass 2: loss, dL/dmu (summed over shards) and optionally dL/dz for the encoder.
        scale = gamma / N_total, i.e. the whole latent set is one batch (models.py:1124-1125).
        With ``out_p`` (and no ``p``) the shard's target distribution rows — what
        ``target_distribution`` (models.py:1302-1322) returns for them — are written there by the
        same pass (``scc_dec_target_kl_grad``)."""
        if p is None and out_p is not None:
            stats, _, dz = ops.dec_target_kl_grad(self.z, mu, f_stats, alpha, round_decimals, gamma / self.n_total,
                                                  out_p=out_p, want_dz=want_dz)
        else:
            stats, dz = ops.dec_kl_grad(self.z, mu, alpha, p=p, f=None if p is not None else f_stats,
                                        round_decimals=round_decimals, scale=gamma / self.n_total, want_dz=want_dz)
        self._allreduce(stats)
        return stats, dz

    def dec_step(self, mu: torch.Tensor, alpha: float = 1.0, gamma: float = 1e-3, round_decimals: int = 0,
                 want_dz: bool = False, want_p: bool = False) -> DecStepResult:
        """Fused latent-buffer DEC step: assign -> (allreduce f) -> KL gradients -> (allreduce dmu).
        With a peer exchange the collectives ride on the kernels: the assign kernel's last CTA pushes f
        to every rank, the gradient kernel pulls it in its prologue and its last CTA all-reduces the gradient
        statistics in its tail — two launches per step (one for the register-blocked shapes), no separate
        collective."""
        K = mu.shape[0]
        one_kernel = getattr(ops, "dec_step_supported", lambda d, k: False)(self.d, K)
        if one_kernel and (self.world == 1 or self.exchange is not None):
            # the whole step in one cooperative kernel; with a peer exchange the all-reduce of f runs inside it
            # (between its two passes) and a one-CTA finish kernel collects the gradient statistics
            prev = self.labels
            out_labels = self._labels_spare if self._labels_spare is not None else torch.empty(
                self.n_local, dtype=torch.int32, device=self.z.device)
            r = ops.dec_step(self.z, mu, alpha, round_decimals, gamma / self.n_total, want_q=False, want_labels=True,
                             want_p=want_p, want_dz=want_dz, labels_prev=prev, out_labels=out_labels,
                             exchange=self.exchange.desc if (self.exchange is not None and self.world > 1) else None)
            self._labels_spare, self.labels = self.labels, r["labels"]
            st, stats = r["f"], r["stats"]
            return DecStepResult(loss=stats[0], dmu=stats[2:].view(K, self.d), f=st[:K], n_changed=st[K], dz=r["dz"],
                                 p=r["p"])
        if self.world == 1 or self.exchange is not None:
            # tiled shapes (K*d > 160): two launches with the u hand-off; with a peer exchange f is pushed by the assign
            # kernel's last CTA, pulled in the gradient kernel's prologue, and the gradient statistics are all-reduced
            # in that kernel's tail
            ex = self.exchange.desc if (self.exchange is not None and self.world > 1) else None
            prev = self.labels
            out_labels = self._labels_spare if self._labels_spare is not None else torch.empty(
                self.n_local, dtype=torch.int32, device=self.z.device)
            if self._u_scratch is None or tuple(self._u_scratch.shape) != (self.n_local, K):
                self._u_scratch = torch.empty(self.n_local, K, dtype=torch.float32, device=self.z.device)
            _, labels, st = ops.dec_assign_u(self.z, mu, self._u_scratch, alpha, round_decimals, want_q=False,
                                             want_labels=True, labels_prev=prev, out_labels=out_labels, push=ex)
            self._labels_spare, self.labels = self.labels, labels
            stats, p_out, dz = ops.dec_target_kl_grad_u(self.z, mu, self._u_scratch, None if ex is not None else st, alpha,
                                                        round_decimals, gamma / self.n_total, want_p=want_p,
                                                        want_dz=want_dz, pull_f=ex, push=ex)
            if ex is None:
                return DecStepResult(loss=stats[0], dmu=stats[2:].view(K, self.d), f=st[:K], n_changed=st[K], dz=dz,
                                     p=p_out)
            # (push on a gradient kernel = complete all-reduce: its last CTA also collects the world's sum)
            # st holds only this shard's f; the all-reduced f is not needed by the caller of a fused step,
            # the label-change count is: exchange it with the (tiny) standalone kernel
            self.exchange.all_reduce(st)
            return DecStepResult(loss=stats[0], dmu=stats[2:].view(K, self.d), f=st[:K], n_changed=st[K], dz=dz,
                                 p=p_out)
        _, st = self.dec_assign(mu, alpha, round_decimals)
        p_out = torch.empty(self.n_local, K, dtype=torch.float32, device=self.z.device) if want_p else None
        stats, dz = self.dec_grad(mu, st, alpha, gamma, round_decimals, want_dz=want_dz, out_p=p_out)
        return DecStepResult(loss=stats[0], dmu=stats[2:].view(K, self.d), f=st[:K], n_changed=st[K], dz=dz, p=p_out)

    def delta_label(self, assign_stats: torch.Tensor) -> float:
        """models.py:1098-1099 from the fused count (one host sync)."""
        return float(assign_stats[-1].item()) / self.n_total

    # ------------------------------------------------------------------ GMM passes
    def gmm_em_pass(self, K: int, params: torch.Tensor, stats: torch.Tensor | None = None, ctrl=None,
                    mode: int = ops.GMM_SOFT, labels: torch.Tensor | None = None):
        stats = ops.gmm_em_step(self.z, K, params, stats=stats, ctrl=ctrl, mode=mode, labels=labels)
        if mode != ops.GMM_ESTEP_ONLY:
            self._allreduce(stats)
        return stats


    def gmm_em_iteration(self, K: int, params, stats, means, weights, covariances, prec_chol, ctrl,
                         reg_covar: float = 1e-6, nk_eps: float = 10 * 2.220446049250313e-16, tol: float = 1e-3):
        """One EM iteration over the sharded set: statistics kernel + one tail kernel that reduces, all-reduces
        (peer exchange inside the kernel) and finalises.  Without a peer exchange on several ranks the three
        steps run separately with the group's all_reduce in between."""
        single = self.group is None or self.world == 1
        fits = self.exchange is not None and stats.numel() <= self.exchange.max_len
        if single or fits:
            ops.gmm_em_iteration(self.z, K, params, stats, self.n_total, means, weights, covariances, prec_chol, ctrl,
                                 reg_covar=reg_covar, nk_eps=nk_eps, tol=tol,
                                 exchange=None if single else self.exchange.desc)
        else:
            self.gmm_em_pass(K, params, stats, ctrl=ctrl)
            ops.gmm_finalize(stats, self.n_total, means, weights, covariances, prec_chol, params, ctrl,
                             reg_covar=reg_covar, nk_eps=nk_eps, tol=tol)


def update_interval(m: int, batch_size: int, config_update_interval: int = -1) -> int:
    """models.py:985-989."""
    if config_update_interval == -1:
        return int(math.ceil(m / (batch_size * 2)))
    return int(math.ceil(m / (batch_size * config_update_interval)))
