"""Two launches of every hot kernel (the second is the one to read), for `ncu --set full`:

  tools/ncu_capture.sh r02_kernels "dec_|gmm_em|gmm_tail|peer_" 0 40 -- python tools/profile_all.py

Order of the launches (each twice, L2 flushed before every launch):
  N=1M  d=9  K=8  rd=5 : dec_assign | dec_target | dec_kl_grad(p)=grad_reg<..,0> | dec_target_kl_grad=<..,3> | dec_step=<..,4>
  N=1M  d=9  K=8       : dec_backward=grad_reg<..,1> (MODE_GENERIC) | kmeans_step=grad_reg<..,2> (MODE_KMEANS)
  N=4M  d=32 K=16      : dec_assign (u hand-off) | dec_target_kl_grad_u = grad_tiled<..,5> | dec_target_kl_grad = grad_tiled<..,3>
  N=4M  d=9  K=16      : gmm_em_sparse (sharp responsibilities, skip on) | gmm_tail (reduce + finalize)
  N=1M  d=32 K=16      : gmm_em_block
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import ops, synth

dev = torch.device("cuda")
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)


def twice(fn):
    for _ in range(2):
        flush.zero_()
        fn()
    torch.cuda.synchronize()


def headline(n=1_000_000, d=9, K=8):
    z, mu = synth.latent_points(n, d, K, device=dev)
    q = torch.empty(n, K, device=dev); p = torch.empty(n, K, device=dev); dz = torch.empty(n, d, device=dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    st1 = torch.empty(K + 1, dtype=torch.float64, device=dev); st2 = torch.empty(K * d + 2, dtype=torch.float64, device=dev)
    g = torch.randn(n, K, device=dev)
    twice(lambda: ops.dec_assign(z, mu, 1.0, 5, out_q=q, out_labels=lab, out_stats=st1))
    twice(lambda: ops.dec_target(q, st1, 5, out=p))
    twice(lambda: ops.dec_kl_grad(z, mu, 1.0, p=p, scale=1e-9, out_dz=dz, out_stats=st2))
    twice(lambda: ops.dec_target_kl_grad(z, mu, st1, 1.0, 5, 1e-9, out_p=p, out_dz=dz, out_stats=st2))
    twice(lambda: ops.dec_step(z, mu, 1.0, 5, 1e-9, out_q=q, out_labels=lab, out_p=p, out_dz=dz, out_f=st1, out_stats=st2))
    twice(lambda: ops.dec_backward(z, mu, g, 1.0))
    twice(lambda: ops.kmeans_step(z, mu, labels=lab))


def shard_d32(n=4_000_000, d=32, K=16):
    z, mu = synth.latent_points(n, d, K, device=dev)
    u = torch.empty(n, K, device=dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    st1 = torch.empty(K + 1, dtype=torch.float64, device=dev); st2 = torch.empty(K * d + 2, dtype=torch.float64, device=dev)
    twice(lambda: ops.dec_assign_u(z, mu, u, 1.0, 0, out_labels=lab, out_stats=st1))
    twice(lambda: ops.dec_target_kl_grad_u(z, mu, u, st1, 1.0, 0, 1e-9, out_stats=st2))
    twice(lambda: ops.dec_target_kl_grad(z, mu, st1, 1.0, 0, 1e-9, want_p=False, want_dz=False, out_stats=st2))


def gmm(n, d, K, warm):
    z, _ = synth.latent_points(n, d, K, device=dev)
    w0, mu0, cov0 = synth.gmm_initial_state(d, K, dev)
    params, pchol, ctrl = ops.gmm_pack_params(w0, mu0, cov0)
    stats = torch.empty(ops.gmm_stat_doubles(K, d), dtype=torch.float64, device=dev)
    means, weights, cov = mu0.clone(), w0.clone(), cov0.clone()
    for _ in range(warm):                     # sharpen the responsibilities as a fit does
        ops.gmm_em_step(z, K, params, stats=stats, ctrl=ctrl)
        ops.gmm_finalize(stats, n, means, weights, cov, pchol, params, ctrl, tol=0.0)
    twice(lambda: ops.gmm_em_iteration(z, K, params, stats, n, means, weights, cov, pchol, ctrl, tol=0.0))


if __name__ == "__main__":
    headline()
    shard_d32()
    gmm(4_000_000, 9, 16, 5)
    gmm(1_000_000, 32, 16, 2)
    print("profile_all done")
