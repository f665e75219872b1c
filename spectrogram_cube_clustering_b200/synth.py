"""Synthetic latent sets and spectrograms of the reference's input shape.

There is no dataset in the reference tree (its Zarr cube path is hard-coded,
``Cluster/ZarrDataLoader.py:96``), so throughput and parity runs use the
generator fixed in SURVEY.md §8(d):

    true centres   c_k ~ U(0,4)^d               (seed 2009, shared by all ranks)
    linear maps    A_k = 0.35 * N(0,1)^{d x d}  (shared)
    component      k_i ~ U{0..K-1}              (seed 2009 + rank)
    latent point   z_i = c_{k_i} + A_{k_i} eps_i,  eps_i ~ N(0, I)

``relu=True`` clamps at 0 like the encoder's final ReLU (``networks.py:184-185``).
Initial DEC centroids / GMM means are ``c_k + 0.1 N(0,1)``; the GMM starts from
Sigma_k = I, pi_k = 1/K.
"""
from __future__ import annotations

import torch

SEED = 2009


def mixture_truth(d: int, K: int, device="cpu", dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(SEED)
    centres = 4.0 * torch.rand(K, d, generator=g, dtype=torch.float64)
    maps = 0.35 * torch.randn(K, d, d, generator=g, dtype=torch.float64)
    init = centres + 0.1 * torch.randn(K, d, generator=g, dtype=torch.float64)
    return (centres.to(device=device, dtype=dtype), maps.to(device=device, dtype=dtype),
            init.to(device=device, dtype=dtype))


def latent_points(n: int, d: int, K: int, rank: int = 0, device="cpu", relu: bool = False,
                  chunk: int = 1 << 22, out: torch.Tensor | None = None):
    """[n, d] float32 latent points on ``device`` plus the initial centroids [K, d]."""
    device = torch.device(device)
    centres, maps, init = mixture_truth(d, K, device)
    g = torch.Generator(device=device).manual_seed(SEED + 1 + rank)
    z = out if out is not None else torch.empty(n, d, device=device, dtype=torch.float32)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        comp = torch.randint(0, K, (m,), generator=g, device=device)
        eps = torch.randn(m, d, generator=g, device=device, dtype=torch.float32)
        blk = z[s:s + m]
        for k in range(K):
            sel = (comp == k).nonzero(as_tuple=True)[0]
            if sel.numel():
                blk[sel] = centres[k] + eps[sel] @ maps[k].T
        if relu:
            blk.clamp_(min=0)
    return z, init


def gmm_initial_state(d: int, K: int, device="cpu"):
    """(pi0 [K], mu0 [K,d], Sigma0 [K,d,d]) float64 — SURVEY.md §8(d)."""
    _, _, init = mixture_truth(d, K, device, torch.float64)
    w = torch.full((K,), 1.0 / K, dtype=torch.float64, device=device)
    cov = torch.eye(d, dtype=torch.float64, device=device).expand(K, d, d).contiguous()
    return w, init, cov


def spectrograms(n: int, rank: int = 0, device="cpu", dtype=torch.float32):
    """(n,1,4,101) synthetic spectrograms (shape: ``models.py:612`` + the
    ``Linear(84,9)`` arithmetic of ``networks.py:176-184``), normalised per
    sample like ``ZarrDataLoader.sample_norm_cent`` (``ZarrDataLoader.py:22-23``)."""
    g = torch.Generator(device=torch.device(device)).manual_seed(SEED + 101 + rank)
    x = torch.randn(n, 1, 4, 101, generator=g, device=device, dtype=dtype)
    amax = x.abs().amax(dim=(1, 2, 3), keepdim=True)      # of the un-centred sample, as the reference
    return (x - x.mean(dim=(1, 2, 3), keepdim=True)) / (amax + 1e-8)


class TensorBatches:
    """Minimal sequential loader over a tensor already in memory (host or device): yields contiguous
    ``batch_size`` slices in order, like ``DataLoader(dataset, batch_size, shuffle=False)``
    (``production.py:131-136``) without the per-item indexing + collate cost."""

    def __init__(self, data: torch.Tensor, batch_size: int):
        self.dataset, self.batch_size = data, int(batch_size)

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for i in range(0, len(self.dataset), self.batch_size):
            yield self.dataset[i:i + self.batch_size]
