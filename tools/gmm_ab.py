"""A/B of the scalar vs packed-FP32 GMM EM kernel (SCC_GMM_VARIANT=s|p), N=4M."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import ops, synth
dev = torch.device("cuda")
for d, K in ((9, 8), (9, 16), (9, 5), (12, 8), (4, 4)):
    n = 4_000_000
    z, _ = synth.latent_points(n, d, K, device=dev)
    w0, mu0, cov0 = synth.gmm_initial_state(d, K, dev)
    params, pchol, ctrl = ops.gmm_pack_params(w0, mu0, cov0)
    stats = torch.empty(ops.gmm_stat_doubles(K, d), dtype=torch.float64, device=dev)
    for _ in range(3):
        ops.gmm_em_step(z, K, params, stats=stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gmm_em_step(z, K, params, stats=stats)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"variant={os.environ.get('SCC_GMM_VARIANT','default')} d={d} K={K}: {ms*1e3:8.1f} us  {n/ms/1e6:6.2f} G pts/s  checksum {stats.sum().item():.6e}")
