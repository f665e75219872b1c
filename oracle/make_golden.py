"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the reference ITSELF.

Run in the build container (needs /root/reference and scikit-learn):

    python -m oracle.make_golden

DEC fixtures come from the reference's own ``ClusteringLayer`` (float64, as
``models.py:965`` runs it), ``target_distribution`` and torch autograd through
``gamma * KLDivLoss('sum')(log q, p) / B`` (``models.py:1124-1127``).
GMM fixtures come from scikit-learn's private ``_e_step`` / ``_m_step`` driven
from an explicit initial state (the deterministic harness of SURVEY.md §8c),
because ``models.gmm`` builds its GaussianMixture unseeded (``models.py:403-409``).
Inputs are float32 values (what the CUDA path consumes) promoted to float64.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import refload  # noqa: E402
from spectrogram_cube_clustering_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def dec_case(networks, models, name, n, d, K, alpha, gamma=1e-3, relu=False, tie=False, seed=0):
    z32, mu32 = synth.latent_points(n, d, K, rank=seed, relu=relu)
    if tie:                                   # duplicate centroids -> exact q ties -> first index wins
        mu32[1] = mu32[0]
        mu32[K - 1] = mu32[2]
    z = z32.double()
    layer = networks.ClusteringLayer(K, d, alpha, weights=mu32.double().clone()).double()
    zt = z.clone().requires_grad_(True)
    q = layer(zt)                                                      # networks.py:279-288
    q_np = q.detach().numpy()
    labels = np.argmax(q_np, axis=1)                                   # models.py:92
    q_round = np.round(q_np, 5)                                        # models.py:94
    p = models.target_distribution(q_round)                            # models.py:1320-1322
    tar = torch.from_numpy(p)
    loss = gamma * torch.nn.KLDivLoss(reduction="sum")(torch.log(q), tar) / n   # models.py:1124-1125
    loss.backward()
    dz, dmu = zt.grad.numpy().copy(), layer.weights.grad.numpy().copy()
    # generic upstream gradient through the layer alone (autograd path of the literal loop)
    g = torch.Generator().manual_seed(7 + seed)
    G = torch.randn(n, K, generator=g, dtype=torch.float64)
    zt2 = z.clone().requires_grad_(True)
    layer.weights.grad = None
    layer(zt2).backward(G)
    out = dict(z=z32.numpy(), mu=mu32.numpy(), alpha=np.float64(alpha), gamma=np.float64(gamma),
               q=q_np, labels=labels.astype(np.int64), q_round=q_round, p=p, f=q_round.sum(0),
               loss=np.float64(loss.item()), dz=dz, dmu=dmu,
               G=G.numpy(), dz_generic=zt2.grad.numpy().copy(),
               dmu_generic=layer.weights.grad.numpy().copy())
    np.savez_compressed(os.path.join(GOLDEN, f"dec_{name}.npz"), **out)
    print(f"dec_{name}: n={n} d={d} K={K} alpha={alpha} loss={loss.item():.6e}")


def gmm_case(name, n, d, K, iters, seed=0, relu=False, tol=1e-3):
    from sklearn.mixture import GaussianMixture

    z32, _ = synth.latent_points(n, d, K, rank=10 + seed, relu=relu)
    w0, mu0, cov0 = [t.numpy() for t in synth.gmm_initial_state(d, K)]
    mu0 = mu0.astype(np.float32).astype(np.float64)
    X = z32.double().numpy()
    prec0 = np.linalg.inv(cov0)

    def fresh(max_iter):
        gm = GaussianMixture(n_components=K, weights_init=w0, means_init=mu0, precisions_init=prec0,
                             random_state=0, max_iter=max_iter, tol=tol, reg_covar=1e-6)
        return gm

    gm = fresh(1)
    try:                                                  # sklearn >= 1.8 threads an array namespace
        from sklearn.utils._array_api import get_namespace
        xp, _ = get_namespace(X)
        gm._check_parameters(X, xp=xp)
    except (ImportError, TypeError):
        gm._check_parameters(X)
    gm._initialize_parameters(X, np.random.RandomState(0))
    hist = dict(lower_bound=[], weights=[], means=[], covariances=[], pchol=[])
    log_resp0 = None
    for it in range(iters):
        lb, log_resp = gm._e_step(X)                      # sklearn _base.py:314-332
        if it == 0:
            log_resp0 = log_resp.copy()
        gm._m_step(X, log_resp)                           # sklearn _gaussian_mixture.py:883-901
        hist["lower_bound"].append(lb)
        hist["weights"].append(gm.weights_.copy())
        hist["means"].append(gm.means_.copy())
        hist["covariances"].append(gm.covariances_.copy())
        hist["pchol"].append(gm.precisions_cholesky_.copy())
    _, log_resp = gm._e_step(X)
    labels_after = np.argmax(log_resp, axis=1)

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        full = fresh(100)
        labels_fit = full.fit_predict(X)                  # the call models.py:411 makes
    out = dict(z=z32.numpy(), w0=w0, mu0=mu0, cov0=cov0, tol=np.float64(tol),
               log_resp0=log_resp0, labels_after=labels_after.astype(np.int64),
               fit_labels=labels_fit.astype(np.int64), fit_n_iter=np.int64(full.n_iter_),
               fit_converged=np.bool_(full.converged_), fit_lower_bound=np.float64(full.lower_bound_),
               fit_weights=full.weights_, fit_means=full.means_, fit_covariances=full.covariances_,
               **{f"it_{k}": np.stack(v) for k, v in hist.items()})
    np.savez_compressed(os.path.join(GOLDEN, f"gmm_{name}.npz"), **out)
    print(f"gmm_{name}: n={n} d={d} K={K} iters={iters} fit_n_iter={full.n_iter_} lb={full.lower_bound_:.6f}")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    networks, models = refload.load()
    dec_case(networks, models, "c1", 1024, 9, 8, 1.0)
    dec_case(networks, models, "k5", 515, 9, 5, 1.0, seed=1)              # DEC_train.py default K=5; ragged N
    dec_case(networks, models, "d32", 384, 32, 16, 1.0, seed=2)
    dec_case(networks, models, "alpha2", 300, 9, 8, 2.0, seed=3)
    dec_case(networks, models, "alpha05", 300, 16, 4, 0.5, seed=4)
    dec_case(networks, models, "relu", 512, 9, 5, 1.0, relu=True, seed=5)  # exact-zero features
    dec_case(networks, models, "tie", 256, 9, 8, 1.0, tie=True, seed=6)
    gmm_case("c1", 4000, 9, 8, 6)
    gmm_case("k16", 6000, 9, 16, 4, seed=1)
    gmm_case("d32", 6000, 32, 16, 3, seed=2)
    gmm_case("relu", 3000, 9, 5, 4, seed=3, relu=True)


if __name__ == "__main__":
    main()
