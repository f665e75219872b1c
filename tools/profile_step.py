"""One launch of every hot kernel, for `ncu --set full` (see profiles/README.md).

  ncu --set full --clock-control none --import-source on -k regex:"dec_|gmm_em" -o OUT python tools/profile_step.py

Order of the profiled launches (after two unprofiled-size warm-ups each, so run with --launch-skip if needed):
  headline shapes N=1M d=9 K=8: dec_assign (q, labels, f) | dec_target | dec_kl_grad(p) | dec_target_kl_grad | dec_step
  configs[3] shard    N=4M d=32 K=16: dec_assign (stats only) | dec_target_kl_grad (centroid-only)
  configs[2] shape    N=4M d=9 K=16: gmm_em_step
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import ops, synth

dev = torch.device("cuda")
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)


def headline(n=1_000_000, d=9, K=8):
    z, mu = synth.latent_points(n, d, K, device=dev)
    q = torch.empty(n, K, device=dev); p = torch.empty(n, K, device=dev); dz = torch.empty(n, d, device=dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    st1 = torch.empty(K + 1, dtype=torch.float64, device=dev); st2 = torch.empty(K * d + 2, dtype=torch.float64, device=dev)
    for rep in range(2):
        flush.zero_()
        ops.dec_assign(z, mu, 1.0, 5, out_q=q, out_labels=lab, out_stats=st1)
        ops.dec_target(q, st1, 5, out=p)
        ops.dec_kl_grad(z, mu, 1.0, p=p, scale=1e-9, out_dz=dz, out_stats=st2)
        flush.zero_()
        ops.dec_target_kl_grad(z, mu, st1, 1.0, 5, 1e-9, out_p=p, out_dz=dz, out_stats=st2)
        flush.zero_()
        ops.dec_step(z, mu, 1.0, 5, 1e-9, out_q=q, out_labels=lab, out_p=p, out_dz=dz, out_f=st1, out_stats=st2)
    torch.cuda.synchronize()


def shard_d32(n=4_000_000, d=32, K=16):
    z, mu = synth.latent_points(n, d, K, device=dev)
    st1 = torch.empty(K + 1, dtype=torch.float64, device=dev); st2 = torch.empty(K * d + 2, dtype=torch.float64, device=dev)
    for rep in range(2):
        flush.zero_()
        ops.dec_assign(z, mu, 1.0, 0, want_q=False, want_labels=False, out_stats=st1)
        ops.dec_target_kl_grad(z, mu, st1, 1.0, 0, 1e-9, want_p=False, want_dz=False, out_stats=st2)
    torch.cuda.synchronize()


def gmm(n=4_000_000, d=9, K=16):
    z, _ = synth.latent_points(n, d, K, device=dev)
    w0, mu0, cov0 = synth.gmm_initial_state(d, K, dev)
    params, pchol, ctrl = ops.gmm_pack_params(w0, mu0, cov0)
    stats = torch.empty(ops.gmm_stat_doubles(K, d), dtype=torch.float64, device=dev)
    for rep in range(2):
        flush.zero_()
        ops.gmm_em_step(z, K, params, stats=stats)
    torch.cuda.synchronize()


if __name__ == "__main__":
    headline()
    shard_d32()
    gmm()
    print("profile_step done")
