"""In-kernel phase timeline of one fused EM iteration (statistics kernel + tail kernel).  Profiling build:
  make -C spectrogram_cube_clustering_b200/csrc variant VNAME=tlg VFLAGS=-DSCC_TIMELINE VOBJS="gmm_dim9.o gmm_api.o scc_api.o"
  SCC_LIB=.../libscc_b200_tlg.so python tools/gmm_timeline.py [n]
Stamps (thread 0 of every CTA, %globaltimer): statistics kernel 0 start, 1 parameters staged + first tiles requested,
2 first tile landed, 3 tile loop done, 4 accumulators flushed, 5 slot written; tail kernel 0 start, 1 slice reduced +
ticket, 2 finalisation done (last CTA); 3..6 inside the finalisation (component 0's warp): covariance built, Cholesky
done, triangular inverse done, per-component outputs written."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spectrogram_cube_clustering_b200 import _lib, ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
d, K = 9, 16
dev = torch.device("cuda")
lib = _lib.load()
tl = torch.zeros(2048 * 8, dtype=torch.int64, device=dev)
z, _ = synth.latent_points(n, d, K, rank=77, device=dev)
w0, mu0, cov0 = synth.gmm_initial_state(d, K, dev)
params, pchol, ctrl = ops.gmm_pack_params(w0, mu0, cov0)
means, weights, cov = mu0.clone(), w0.clone(), cov0.clone()
stats = torch.empty(ops.gmm_stat_doubles(K, d), dtype=torch.float64, device=dev)
it = lambda: ops.gmm_em_iteration(z, K, params, stats, n, means, weights, cov, pchol, ctrl, tol=0.0)
for _ in range(6):
    it()
torch.cuda.synchronize()
for rep in range(2):
    tl.zero_()
    _lib.check(lib.scc_debug_set_timeline(tl.data_ptr()), "timeline build?")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); it(); e1.record()
    torch.cuda.synchronize()
    lib.scc_debug_set_timeline(None)
    t = tl.view(-1, 8).cpu().double()
    em, tail = t[:1024], t[1024:]
    em = em[em[:, 0] > 0]; tail = tail[tail[:, 0] > 0]
    t0 = em[:, 0].min()
    em = (em - t0) / 1e3; tail = (tail - t0) / 1e3
    print(f"\nn={n} d={d} K={K}: event time of the iteration {e0.elapsed_time(e1) * 1e3:.1f} us; statistics kernel {em.shape[0]} CTAs, tail {tail.shape[0]} CTAs")
    for k, nm in enumerate(["start", "prologue", "first tile", "tile loop", "flush", "slot written"]):
        c = em[:, k]
        print(f"  stats  {nm:<13} {c.min():8.2f} {c.median():8.2f} {c.max():8.2f}")
    for k, nm in enumerate(["start", "slice+ticket", "finalize", "fin: cov built", "fin: cholesky", "fin: inverse", "fin: outputs"]):
        c = tail[:, k]; c = c[c > 0]
        if len(c):
            print(f"  tail   {nm:<13} {c.min():8.2f} {c.median():8.2f} {c.max():8.2f}")
