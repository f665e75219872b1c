"""Is the skew between the CTAs of the one-kernel step systematic (the same CTAs slow every launch)?
Run as  SCC_LIB=.../libscc_b200_tl9.so python tools/skew_probe.py  (profiling build, see tools/timeline.py).
Prints, over R launches, the correlation between launches of each CTA's pass-1 / pass-2 duration and the spread
that would remain if every CTA's systematic part were removed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from spectrogram_cube_clustering_b200 import _lib, ops, synth

n, d, K, R = 1_000_000, 9, 8, 8
dev = torch.device("cuda")
lib = _lib.load()
tl = torch.zeros(2048 * 8, dtype=torch.int64, device=dev)
z, mu = synth.latent_points(n, d, K, device=dev)
bufs = ops.dec_step(z, mu, 1.0, 5, 1e-9)
fn = lambda: ops.dec_step(z, mu, 1.0, 5, 1e-9, out_q=bufs["q"], out_labels=bufs["labels"], out_p=bufs["p"],
                          out_dz=bufs["dz"], out_f=bufs["f"], out_stats=bufs["stats"])
flush = torch.empty(300 << 20, dtype=torch.uint8, device=dev)
p1, p2, end = [], [], []
for r in range(R):
    for _ in range(3):
        fn()
    if os.environ.get("FLUSH"):
        flush.zero_()
    tl.zero_()
    _lib.check(lib.scc_debug_set_timeline(tl.data_ptr()), "timeline build?")
    fn()
    torch.cuda.synchronize()
    lib.scc_debug_set_timeline(None)
    t = tl.view(-1, 8).cpu().numpy().astype(np.float64)
    t = t[t[:, 0] > 0]
    p1.append((t[:, 6] - t[:, 1]) / 1e3)
    p2.append((t[:, 3] - t[:, 7]) / 1e3)
    end.append((t[:, 3] - t[:, 0].min()) / 1e3)
p1, p2, end = np.array(p1), np.array(p2), np.array(end)
np.save("gpurun_out/skew_p1.npy", p1); np.save("gpurun_out/skew_p2.npy", p2)
for name, a in (("pass 1", p1), ("pass 2", p2)):
    c = np.corrcoef(a)
    off = c[np.triu_indices(R, 1)]
    sysm = a.mean(0)                     # per-CTA systematic part
    resid = a - sysm[None, :]
    print(f"{name}: per-launch spread (max-min) {np.ptp(a, axis=1).mean():.2f} us, mean {a.mean():.2f} us; "
          f"correlation between launches {off.mean():.2f} (min {off.min():.2f}); spread of the per-CTA means {np.ptp(sysm):.2f} us, "
          f"residual spread {np.ptp(resid, axis=1).mean():.2f} us")
c12 = np.corrcoef(p1.mean(0), p2.mean(0))[0, 1]
print(f"correlation of a CTA's mean pass-1 and pass-2 durations: {c12:.2f}")
slow = np.argsort(-p2.mean(0))[:12]
print("slowest CTAs in pass 2:", slow.tolist(), np.round(p2.mean(0)[slow], 2).tolist())
print("fastest CTAs in pass 2:", np.argsort(p2.mean(0))[:12].tolist())
