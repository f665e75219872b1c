"""One-kernel step (ops.dec_step) vs the two-kernel chain at the headline shape, both replayed as 16-step CUDA graphs
over 4 rotating input sets (python tools/time_step.py)."""
import sys, os, torch
sys.path.insert(0, os.getcwd())
from spectrogram_cube_clustering_b200 import ops, synth
dev = torch.device("cuda")
n, d, K = 1_000_000, 9, 8
sets = []
for s in range(4):
    z, mu = synth.latent_points(n, d, K, rank=s, device=dev)
    sets.append(dict(z=z, q=torch.empty(n, K, device=dev), p=torch.empty(n, K, device=dev), dz=torch.empty(n, d, device=dev),
                     lab=torch.empty(n, dtype=torch.int32, device=dev), f=torch.empty(K + 1, dtype=torch.float64, device=dev),
                     st=torch.empty(K * d + 2, dtype=torch.float64, device=dev)))
def one(s):
    ops.dec_step(s["z"], mu, 1.0, 5, 1e-9, out_q=s["q"], out_labels=s["lab"], out_p=s["p"], out_dz=s["dz"], out_f=s["f"], out_stats=s["st"])
def two(s):
    ops.dec_assign(s["z"], mu, 1.0, 5, out_q=s["q"], out_labels=s["lab"], out_stats=s["f"])
    ops.dec_target_kl_grad(s["z"], mu, s["f"], 1.0, 5, 1e-9, out_p=s["p"], out_dz=s["dz"], out_stats=s["st"])
for name, fn in (("two-kernel", two), ("one-kernel", one)):
    cs = torch.cuda.Stream()
    with torch.cuda.stream(cs):
        for s in sets: fn(s)
        cs.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cs):
            for _ in range(4):
                for s in sets: fn(s)
    torch.cuda.synchronize()
    for _ in range(20): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(60): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / (60 * 16) * 1e3:.2f} us/step")
