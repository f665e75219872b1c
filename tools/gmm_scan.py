"""GMM EM pass timing: kernel variant (SCC_GMM_VARIANT) x sparsity skip, after `warm` EM iterations so that the
responsibilities are as sharp as they are during a fit.  python tools/gmm_scan.py [d K n warm]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import ops, synth

dev = torch.device("cuda")
d, K, n, warm = (int(x) for x in (sys.argv[1:] + ["9", "16", "10000000", "5"][len(sys.argv) - 1:]))
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
z, _ = synth.latent_points(n, d, K, rank=77, device=dev)
w0, mu0, cov0 = synth.gmm_initial_state(d, K, dev)


def run(label, mode, warm_iters):
    params, pchol, ctrl = ops.gmm_pack_params(w0, mu0, cov0)
    means, weights, cov = mu0.clone(), w0.clone(), cov0.clone()
    stats = torch.empty(ops.gmm_stat_doubles(K, d), dtype=torch.float64, device=dev)
    for _ in range(warm_iters):
        ops.gmm_em_step(z, K, params, stats=stats, ctrl=ctrl, mode=mode if mode else ops.GMM_SOFT)
        ops.gmm_finalize(stats, n, means, weights, cov, pchol, params, ctrl, tol=0.0)
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gmm_em_step(z, K, params, stats=stats, mode=mode); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    flops = 2.0 * K * (d * d + 4 * d) * n
    print(f"{label:34s} d={d} K={K} n={n} warm={warm_iters}: {ms * 1e3:9.1f} us  {n / ms / 1e6:7.2f} G pts/s  "
          f"{flops / ms / 1e9:6.1f} alg TFLOP/s   lower bound {ctrl[0].item():.6f}", flush=True)
    return stats.clone()


variant = os.environ.get("SCC_GMM_VARIANT", "sparse")
run(f"{variant} E-step only (sharp state)", ops.GMM_ESTEP_ONLY, warm)
s_skip = run(f"{variant} skip=on", ops.GMM_SOFT, warm)
s_dense = run(f"{variant} skip=off", ops.GMM_SOFT | ops.GMM_NOSKIP, warm)
run(f"{variant} skip=on (first iteration)", ops.GMM_SOFT, 0)
rel = ((s_skip - s_dense).abs().max() / s_dense.abs().max()).item()
print(f"max |stats(skip) - stats(noskip)| / max |stats| = {rel:.3e}")
