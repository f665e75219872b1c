"""Host-side mirror of ``Cluster/networks.py`` for the clustering path.

``ClusteringLayer`` keeps the reference constructor, the ``weights`` parameter
([K, d], state-dict key ``clustering.weights``) and the ``forward(x) -> q``
contract (``networks.py:251-288``) but runs the fused sm_100a kernels through the
registered custom ops ``torch.ops.scc_b200.soft_assign`` / ``dec_kl_loss``: forward
is one ``dec_assign`` launch, backward one ``dec_backward`` launch (the
reference builds a [B,K,d] temporary and ~15 autograd nodes).  ``DEC`` wires it
behind the stock-torch encoder / decoder exactly like ``networks.py:291-323``
(the convolutional autoencoder is out of scope and stays plain torch).

There is no CPU path: calling the layer on a CPU tensor raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import SccError


def _pad_cols(t: torch.Tensor, width: int) -> torch.Tensor:
    return t if t.shape[1] == width else F.pad(t, (0, width - t.shape[1]))


def _operands(x, weights):
    """Kernel operands.  float64 input (the reference's ``model.double()``, ``models.py:965``) stays float64 and runs
    the float64 kernels; anything else becomes float32, contiguous, padded to an instantiated latent dimension
    (zero-padding z and mu is exact)."""
    if not x.is_cuda:
        raise SccError("the clustering layer needs CUDA tensors: the B200 path has no CPU fallback")
    if x.dtype == torch.float64:
        return x.contiguous(), weights.to(device=x.device, dtype=torch.float64).contiguous()
    dp = ops.padded_dim(x.shape[1])
    x32 = _pad_cols(x.to(torch.float32), dp).contiguous()
    w32 = _pad_cols(weights.to(device=x.device, dtype=torch.float32), dp).contiguous()
    return x32, w32


def soft_assign(x, weights, alpha=1.0):
    """q = Student's-t soft assignment of x [B, d] to the centroids weights [K, d] (``networks.py:279-288``) through
    ``torch.ops.scc_b200.soft_assign``: forward is one ``scc_dec_assign`` launch, backward one
    ``scc_dec_backward`` launch (q is recomputed, only x and weights are saved).  The casts / padding around the op
    are ordinary differentiable torch ops, so gradients arrive in the callers' dtypes (float64 for the reference's
    ``model.double()``)."""
    x32, w32 = _operands(x, weights)
    return torch.ops.scc_b200.soft_assign(x32, w32, float(alpha)).to(x.dtype)


def dec_kl_loss(z, weights, p, alpha=1.0, scale=1.0):
    """Fused clustering loss: ``scale * sum_ij p_ij (log p_ij - log q_ij)`` with q recomputed from
    (z, weights); differentiable w.r.t. z and weights.  ``scale = gamma / batch_size`` reproduces
    ``models.py:1124-1125``.  One ``scc_dec_kl_grad`` launch yields loss, dL/dz and dL/dweights
    (``torch.ops.scc_b200.dec_kl_loss``); backward only scales them."""
    z32, w32 = _operands(z, weights)
    p32 = p.detach().to(device=z.device, dtype=z32.dtype).contiguous()
    loss, _, _ = torch.ops.scc_b200.dec_kl_loss(z32, w32, p32, float(alpha), float(scale))
    return loss.to(z.dtype)


class ClusteringLayer(nn.Module):
    """Student's-t soft assignment (drop-in for ``Cluster.networks.ClusteringLayer``).

    Arguments (same as the reference, ``networks.py:265``):
        n_clusters, n_features=9, alpha=1.0, weights=None (initial centroids [K, d])
    Input  x [B, n_features] (CUDA).  float32 (or half) input runs the float32 throughput kernels; float64 input —
           the reference's ``model.double()`` — runs the float64 kernels (reference precision, ~10x slower)
    Output q [B, n_clusters] in x.dtype, rows sum to 1.
    """

    def __init__(self, n_clusters, n_features=9, alpha=1.0, weights=None):
        super().__init__()
        self.n_features = int(n_features)
        self.n_clusters = int(n_clusters)
        self.alpha = alpha
        if self.n_clusters < 1 or self.n_clusters > ops.MAX_K:
            raise ValueError(f"n_clusters must be in [1, {ops.MAX_K}]")
        if weights is None:
            initial_weights = torch.zeros(self.n_clusters, self.n_features, dtype=torch.float)
            nn.init.xavier_uniform_(initial_weights)
        else:
            initial_weights = torch.as_tensor(weights)
        self.weights = nn.Parameter(initial_weights)

    def forward(self, x):
        if x.dim() != 2 or x.shape[1] != self.weights.shape[1]:
            raise ValueError(f"expected x of shape [B, {self.weights.shape[1]}], got {tuple(x.shape)}")
        return soft_assign(x, self.weights, float(self.alpha))

    def extra_repr(self):
        return f"n_clusters={self.n_clusters}, n_features={self.n_features}, alpha={self.alpha}"


# --------------------------------------------------------------------------- stock-torch autoencoder
class _SpatialAttention(nn.Module):
    """Mean/max channel pooling -> 3x3 conv -> x * sigmoid(x)  (``networks.py:157-168``)."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(2, 1, kernel_size=3, padding=1, bias=False)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):
        pooled = torch.cat([x.mean(dim=1, keepdim=True), x.amax(dim=1, keepdim=True)], dim=1)
        y = self.conv(pooled)
        return y * self.sigmoid(y)


def _conv_stack(transposed: bool):
    kw = dict(kernel_size=(2, 4), stride=(1, 2), padding=1)
    return nn.ConvTranspose2d if transposed else nn.Conv2d, kw


class Encoder(nn.Module):
    """(B,1,4,101) spectrogram -> 9 ReLU latent features; layer indices match the reference's
    ``nn.Sequential`` so its checkpoints load (``networks.py:172-189``)."""

    def __init__(self):
        super().__init__()
        conv, kw = _conv_stack(False)
        layers = []
        for cin in (1, 8, 8):
            layers += [conv(cin, 8, **kw), nn.ReLU(True)]
        layers += [_SpatialAttention(), nn.Flatten(), nn.Linear(84, 9), nn.ReLU(True)]
        self.encoder = nn.Sequential(*layers)

    def forward(self, x):
        return self.encoder(x)


class Decoder(nn.Module):
    """9 latent features -> (B,1,4,101) reconstruction (``networks.py:194-214``)."""

    def __init__(self):
        super().__init__()
        convt, kw = _conv_stack(True)
        self.decoder = nn.Sequential(
            nn.Linear(9, 84), nn.ReLU(True), nn.Unflatten(1, (1, 7, 12)),
            nn.ConvTranspose2d(1, 8, kernel_size=3, padding=1, bias=False), nn.ReLU(True),
            convt(8, 8, output_padding=(0, 1), **kw), nn.ReLU(True),
            convt(8, 8, **kw), nn.ReLU(True),
            convt(8, 1, output_padding=(0, 1), **kw), nn.ReLU(True),
        )

    def forward(self, x):
        return self.decoder(x)


class DEC(nn.Module):
    """Encoder + decoder + B200 clustering layer; ``forward(x) -> (q, x_rec, z)``
    (``networks.py:291-323``).  Unlike the reference, the latent width and alpha can be set."""

    def __init__(self, n_clusters, n_features=9, alpha=1.0):
        super().__init__()
        self.n_clusters = n_clusters
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.clustering = ClusteringLayer(self.n_clusters, n_features=n_features, alpha=alpha)

    def forward(self, x):
        z = self.encoder(x)
        x_rec = self.decoder(z)
        q = self.clustering(z)
        return q, x_rec, z
