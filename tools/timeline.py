"""In-kernel phase timeline of the DEC kernels (needs the profiling build: `make -C .../csrc timeline`).

Run as  SCC_LIB=spectrogram_cube_clustering_b200/libscc_b200_timeline.so python tools/timeline.py [N d K]
Every CTA's thread 0 stamps %globaltimer at: 0 after the dependency wait, 1 prologue done, 2 first
z tile landed, 3 main loop done, 4 CTA reduction done, 5 grid reduction / kernel end.  The table
shows, over all CTAs, when each phase boundary is reached relative to the earliest CTA start
(min / median / max, microseconds) plus the CUDA-event duration of the launch.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import _lib, ops, synth

NAMES = ["start", "prologue", "first tile", "main loop", "cta reduce", "end"]


def main(n=1_000_000, d=9, K=8):
    dev = torch.device("cuda")
    lib = _lib.load()
    tl = torch.zeros(2048 * 8, dtype=torch.int64, device=dev)
    z, mu = synth.latent_points(n, d, K, device=dev)
    q = torch.empty(n, K, device=dev); p = torch.empty(n, K, device=dev); dz = torch.empty(n, d, device=dev)
    lab = torch.empty(n, dtype=torch.int32, device=dev)
    st1 = torch.empty(K + 1, dtype=torch.float64, device=dev)
    st2 = torch.empty(K * d + 2, dtype=torch.float64, device=dev)
    flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
    cases = {
        "dec_assign (q, labels, f)": lambda: ops.dec_assign(z, mu, 1.0, 5, out_q=q, out_labels=lab, out_stats=st1),
        "dec_kl_grad (p given, dz)": lambda: ops.dec_kl_grad(z, mu, 1.0, p=p, scale=1e-9, out_dz=dz, out_stats=st2),
        "dec_kl_grad (fused p from f, no dz)": lambda: ops.dec_kl_grad(z, mu, 1.0, f=st1, scale=1e-9, want_dz=False,
                                                                       out_stats=st2),
        "dec_target_kl_grad (p out, dz)": lambda: ops.dec_target_kl_grad(z, mu, st1, 1.0, 5, 1e-9, out_p=p, out_dz=dz,
                                                                        out_stats=st2),
    }
    if ops.dec_step_supported(d, K):
        stf = torch.empty(K + 1, dtype=torch.float64, device=dev)
        cases["dec_step (one kernel: q, labels, f | barrier | p, dz, stats)"] = lambda: ops.dec_step(
            z, mu, 1.0, 5, 1e-9, out_q=q, out_labels=lab, out_p=p, out_dz=dz, out_f=stf, out_stats=st2)
    ops.dec_assign(z, mu, 1.0, 5, out_q=q, out_labels=lab, out_stats=st1)
    ops.dec_target(q, st1, 5, out=p)
    print(f"N={n} d={d} K={K}; times in us relative to the earliest CTA start (min / median / max over CTAs)")
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        flush.zero_()
        tl.zero_()
        _lib.check(lib.scc_debug_set_timeline(tl.data_ptr()), "scc_debug_set_timeline (is SCC_LIB the timeline build?)")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        lib.scc_debug_set_timeline(None)
        t = tl.view(-1, 8).cpu()
        t = t[t[:, 0] > 0].double()
        t0 = t[:, 0].min()
        two_pass = bool((t[:, 6] > 0).any())
        t = (t - t0) / 1e3
        print(f"\n{name}: {t.shape[0]} CTAs, event time {e0.elapsed_time(e1) * 1e3:.1f} us")
        order = [(0, "start"), (1, "prologue")] + ([(6, "pass 1 loop"), (7, "grid barrier")] if two_pass else []) + \
                [(2, "first tile"), (3, "main loop"), (4, "cta reduce"), (5, "end")]
        for k, nm in order:
            col = t[:, k]
            col = col[col >= 0]
            print(f"  {nm:<12} {col.min():8.2f} {col.median():8.2f} {col.max():8.2f}")
        busy = (t[:, 3] - t[:, 2])
        print(f"  main-loop span per CTA: min {busy.min():.2f} median {busy.median():.2f} max {busy.max():.2f}")


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*a)
