"""d = 32, K = 16 (BASELINE configs[3] shard shape): the two passes of the fused latent-buffer DEC step.
python tools/d32_scan.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import ops, synth

dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
d, K = 32, 16
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
z, mu = synth.latent_points(n, d, K, rank=78, device=dev)
u = torch.empty(n, K, device=dev)
dz = torch.empty_like(z)
lab = torch.empty(n, dtype=torch.int32, device=dev)
st1 = torch.empty(K + 1, dtype=torch.float64, device=dev)
st2 = torch.empty(K * d + 2, dtype=torch.float64, device=dev)


def t(fn, reps=7):
    fn(); fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


rows = [
    ("assign (stats only)", lambda: ops.dec_assign(z, mu, 1.0, 0, want_q=False, want_labels=False, out_stats=st1), 4 * d),
    ("assign (labels)", lambda: ops.dec_assign(z, mu, 1.0, 0, want_q=False, out_labels=lab, out_stats=st1), 4 * d + 4),
    ("assign_u (labels + u)", lambda: ops.dec_assign_u(z, mu, u, 1.0, 0, out_labels=lab, out_stats=st1), 4 * d + 4 + 4 * K),
    ("grad recompute (no dz)", lambda: ops.dec_target_kl_grad(z, mu, st1, 1.0, 0, 1e-9, want_p=False, want_dz=False, out_stats=st2), 4 * d),
    ("grad u hand-off (no dz)", lambda: ops.dec_target_kl_grad_u(z, mu, u, st1, 1.0, 0, 1e-9, out_stats=st2), 4 * d + 4 * K),
    ("grad recompute (dz)", lambda: ops.dec_target_kl_grad(z, mu, st1, 1.0, 0, 1e-9, want_p=False, out_dz=dz, out_stats=st2), 8 * d),
    ("grad u hand-off (dz)", lambda: ops.dec_target_kl_grad_u(z, mu, u, st1, 1.0, 0, 1e-9, out_dz=dz, out_stats=st2), 8 * d + 4 * K),
]
print(f"d={d} K={K} n={n}  (median us, L2 flushed)")
for name, fn, bpp in rows:
    us = t(fn)
    print(f"  {name:28s} {us:9.1f} us  {n / us / 1e3:7.2f} G pts/s  {bpp * n / us / 1e3:8.1f} GB/s moved  "
          f"{us * 1e-6 * 1.965e9 * 148 / n:6.2f} SM-cycles/pt", flush=True)
