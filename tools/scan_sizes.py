"""Kernel time vs N (slope = steady-state cost per point, intercept = fixed launch/prologue/tail cost)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import ops, synth

dev = torch.device("cuda")
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3   # us


def main(d=9, K=8):
    print(f"d={d} K={K}   (median us, L2 flushed before each launch)")
    print(f"{'N':>10} {'assign':>9} {'assign_noq':>10} {'target':>8} {'grad_p':>8} {'grad_f':>8} {'grad_f_nodz':>11} {'tgrad_p_out':>11} {'dec_step':>9} {'gmm_em':>9}")
    for n in (250_000, 1_000_000, 4_000_000, 16_000_000):
        z, mu = synth.latent_points(n, d, K, device=dev)
        q = torch.empty(n, K, device=dev); p = torch.empty(n, K, device=dev); dz = torch.empty(n, d, device=dev)
        lab = torch.empty(n, dtype=torch.int32, device=dev)
        st1 = torch.empty(K + 1, dtype=torch.float64, device=dev); st2 = torch.empty(K * d + 2, dtype=torch.float64, device=dev)
        ops.dec_assign(z, mu, 1.0, 5, out_q=q, out_labels=lab, out_stats=st1)
        ops.dec_target(q, st1, 5, out=p)
        t = {}
        t["assign"] = timeit(lambda: ops.dec_assign(z, mu, 1.0, 5, out_q=q, out_labels=lab, out_stats=st1))
        t["assign_noq"] = timeit(lambda: ops.dec_assign(z, mu, 1.0, 0, want_q=False, want_labels=False, out_stats=st1))
        t["target"] = timeit(lambda: ops.dec_target(q, st1, 5, out=p))
        t["grad_p"] = timeit(lambda: ops.dec_kl_grad(z, mu, 1.0, p=p, scale=1e-9, out_dz=dz, out_stats=st2))
        t["grad_f"] = timeit(lambda: ops.dec_kl_grad(z, mu, 1.0, f=st1, scale=1e-9, out_dz=dz, out_stats=st2))
        t["grad_f_nodz"] = timeit(lambda: ops.dec_kl_grad(z, mu, 1.0, f=st1, scale=1e-9, want_dz=False, out_stats=st2))
        t["tgrad"] = timeit(lambda: ops.dec_target_kl_grad(z, mu, st1, 1.0, 5, 1e-9, out_p=p, out_dz=dz, out_stats=st2))
        t["step"] = float("nan")
        if ops.dec_step_supported(d, K):
            stf = torch.empty(K + 1, dtype=torch.float64, device=dev)
            t["step"] = timeit(lambda: ops.dec_step(z, mu, 1.0, 5, 1e-9, out_q=q, out_labels=lab, out_p=p, out_dz=dz,
                                                    out_f=stf, out_stats=st2))
        if ops.gmm_supported(d, K):
            w0, mu0, cov0 = synth.gmm_initial_state(d, K, dev)
            params, pchol, ctrl = ops.gmm_pack_params(w0, mu0, cov0)
            stats = torch.empty(ops.gmm_stat_doubles(K, d), dtype=torch.float64, device=dev)
            t["gmm"] = timeit(lambda: ops.gmm_em_step(z, K, params, stats=stats), reps=8)
        else:
            t["gmm"] = float("nan")
        print(f"{n:>10} {t['assign']:9.1f} {t['assign_noq']:10.1f} {t['target']:8.1f} {t['grad_p']:8.1f} {t['grad_f']:8.1f} {t['grad_f_nodz']:11.1f} {t['tgrad']:11.1f} {t['step']:9.1f} {t['gmm']:9.1f}")
        del z, q, p, dz


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 9, int(sys.argv[2]) if len(sys.argv) > 2 else 8)
