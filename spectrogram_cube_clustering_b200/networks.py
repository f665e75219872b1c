"""Host-side mirror of ``Cluster/networks.py`` for the clustering path.

``ClusteringLayer`` keeps the reference constructor, the ``weights`` parameter
([K, d], state-dict key ``clustering.weights``) and the ``forward(x) -> q``
contract (``networks.py:251-288``) but runs the fused sm_100a kernels: forward
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


class _SoftAssign(torch.autograd.Function):
    """q = softassign(x, weights); saves only x and weights (q is recomputed in backward)."""

    @staticmethod
    def forward(ctx, x, weights, alpha):
        if not x.is_cuda:
            raise SccError("ClusteringLayer needs CUDA tensors: the B200 path has no CPU fallback")
        d = x.shape[1]
        dp = ops.padded_dim(d)                      # zero-padding z and mu is exact for distances
        x32 = _pad_cols(x.detach().to(torch.float32), dp).contiguous()
        w32 = _pad_cols(weights.detach().to(device=x.device, dtype=torch.float32), dp).contiguous()
        q, _, _ = ops.dec_assign(x32, w32, alpha, 0, want_labels=False)
        ctx.save_for_backward(x32, w32)
        ctx.alpha, ctx.d = alpha, d
        ctx.x_dtype, ctx.w_dtype = x.dtype, weights.dtype
        return q.to(x.dtype)

    @staticmethod
    def backward(ctx, grad_q):
        x32, w32 = ctx.saved_tensors
        g32 = grad_q.to(torch.float32).contiguous()
        dz, dmu = ops.dec_backward(x32, w32, g32, ctx.alpha, want_dz=ctx.needs_input_grad[0])
        gx = dz[:, :ctx.d].to(ctx.x_dtype) if dz is not None else None
        gw = dmu[:, :ctx.d].to(ctx.w_dtype) if ctx.needs_input_grad[1] else None
        return gx, gw, None


class _FusedKLLoss(torch.autograd.Function):
    """scale * KL(p || softassign(z, weights)) with loss, dL/dz and dL/dweights from ONE launch
    (``scc_dec_kl_grad``) — the fused form of ``gamma * KLDivLoss('sum')(log q, p) / B`` + backward
    (``models.py:1124-1127``)."""

    @staticmethod
    def forward(ctx, z, weights, p, alpha, scale):
        if not z.is_cuda:
            raise SccError("dec_kl_loss needs CUDA tensors: the B200 path has no CPU fallback")
        d = z.shape[1]
        dp = ops.padded_dim(d)
        z32 = _pad_cols(z.detach().to(torch.float32), dp).contiguous()
        w32 = _pad_cols(weights.detach().to(device=z.device, dtype=torch.float32), dp).contiguous()
        p32 = p.detach().to(device=z.device, dtype=torch.float32).contiguous()
        stats, dz = ops.dec_kl_grad(z32, w32, alpha, p=p32, scale=scale, want_dz=True)
        K = w32.shape[0]
        ctx.save_for_backward(dz[:, :d].to(z.dtype), stats[2:].view(K, dp)[:, :d].to(weights.dtype))
        return stats[0].to(z.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        dz, dmu = ctx.saved_tensors
        return grad_out * dz, grad_out * dmu, None, None, None


def dec_kl_loss(z, weights, p, alpha=1.0, scale=1.0):
    """Fused clustering loss: ``scale * sum_ij p_ij (log p_ij - log q_ij)`` with q recomputed from
    (z, weights); differentiable w.r.t. z and weights.  ``scale = gamma / batch_size`` reproduces
    ``models.py:1124-1125``."""
    return _FusedKLLoss.apply(z, weights, p, float(alpha), float(scale))


class ClusteringLayer(nn.Module):
    """Student's-t soft assignment (drop-in for ``Cluster.networks.ClusteringLayer``).

    Arguments (same as the reference, ``networks.py:265``):
        n_clusters, n_features=9, alpha=1.0, weights=None (initial centroids [K, d])
    Input  x [B, n_features] (CUDA; float32 or float64 — the kernels compute in float32)
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
        return _SoftAssign.apply(x, self.weights, float(self.alpha))

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
