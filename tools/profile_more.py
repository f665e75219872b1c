"""One launch (x2) of the kernels that profile_step.py does not cover, for `ncu --set full`:

  ncu --set full --clock-control none --import-source on -k regex:"gmm_em|dec_grad|dec_assign" -o OUT python tools/profile_more.py

  N=4M d=9  K=16 : gmm_em (full variant)            BASELINE configs[2] shape
  N=1M d=32 K=16 : gmm_em (block variant)           configs[3] shape
  N=4M d=12 K=8  : gmm_em (packed variant)
  N=1M d=9  K=8  : dec_backward (MODE_GENERIC), kmeans_step (MODE_KMEANS)
  N=4M d=32 K=16 : dec_assign + dec_target_kl_grad with dz (tiled), dec_step if supported
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import ops, synth

dev = torch.device("cuda")
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)


def gmm(n, d, K):
    z, _ = synth.latent_points(n, d, K, device=dev)
    w0, mu0, cov0 = synth.gmm_initial_state(d, K, dev)
    params, pchol, ctrl = ops.gmm_pack_params(w0, mu0, cov0)
    stats = torch.empty(ops.gmm_stat_doubles(K, d), dtype=torch.float64, device=dev)
    means, weights, cov = mu0.clone(), w0.clone(), cov0.clone()
    # two EM iterations first so the responsibilities are as sparse as they are in a real fit
    for _ in range(2):
        ops.gmm_em_step(z, K, params, stats=stats, ctrl=ctrl)
        ops.gmm_finalize(stats, n, means, weights, cov, pchol, params, ctrl, tol=0.0)
    for rep in range(2):
        flush.zero_()
        ops.gmm_em_step(z, K, params, stats=stats)
    torch.cuda.synchronize()


def modes(n=1_000_000, d=9, K=8):
    z, mu = synth.latent_points(n, d, K, device=dev)
    g = torch.randn(n, K, device=dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    for rep in range(2):
        flush.zero_()
        ops.dec_backward(z, mu, g, 1.0)
        flush.zero_()
        ops.kmeans_step(z, mu, labels=lab)
    torch.cuda.synchronize()


def shard_d32(n=4_000_000, d=32, K=16):
    z, mu = synth.latent_points(n, d, K, device=dev)
    st1 = torch.empty(K + 1, dtype=torch.float64, device=dev)
    st2 = torch.empty(K * d + 2, dtype=torch.float64, device=dev)
    dz = torch.empty_like(z)
    for rep in range(2):
        flush.zero_()
        ops.dec_assign(z, mu, 1.0, 0, want_q=False, want_labels=False, out_stats=st1)
        ops.dec_target_kl_grad(z, mu, st1, 1.0, 0, 1e-9, want_p=False, out_dz=dz, out_stats=st2)
        if ops.dec_step_supported(d, K):
            flush.zero_()
            ops.dec_step(z, mu, 1.0, 0, 1e-9, want_q=False, want_labels=False, want_p=False, want_dz=False,
                         out_f=st1, out_stats=st2)
    torch.cuda.synchronize()


if __name__ == "__main__":
    which = sys.argv[1:] or ["gmm9", "gmm32", "gmm12", "modes", "d32"]
    if "gmm9" in which:
        gmm(4_000_000, 9, 16)
    if "gmm32" in which:
        gmm(1_000_000, 32, 16)
    if "gmm12" in which:
        gmm(4_000_000, 12, 8)
    if "modes" in which:
        modes()
    if "d32" in which:
        shard_d32()
    print("profile_more done")
