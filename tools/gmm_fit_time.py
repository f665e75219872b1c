"""Time per EM iteration of a graph-captured GaussianMixture.fit (tol = 0, fixed initial state):
python tools/gmm_fit_time.py [n ...]   (SCC_LIB selects the build)"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import synth
from spectrogram_cube_clustering_b200.latent_buffer import LatentBuffer
from spectrogram_cube_clustering_b200.models import GaussianMixture

d, K, iters = 9, 16, 100
dev = torch.device("cuda")
w0, mu0, cov0 = [t.numpy() for t in synth.gmm_initial_state(d, K, "cpu")]
for n in [int(x) for x in sys.argv[1:]] or [1_250_000, 10_000_000]:
    z, _ = synth.latent_points(n, d, K, rank=77, device=dev)
    buf = LatentBuffer(z)
    best = 1e9
    for rep in range(3):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gm = GaussianMixture(K, max_iter=iters, tol=0.0, weights_init=w0, means_init=mu0, covariances_init=cov0,
                                 poll_interval=iters)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            gm.fit(buf)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = min(best, dt / gm.n_iter_)
    print(f"n={n}: {best * 1e6:8.1f} us per EM iteration ({gm.n_iter_} iterations, best of 3 fits)  lower bound {gm.lower_bound_:.6f}")
