"""Multi-GPU correctness of the sharded path on real devices (one process per GPU):

    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tools/multigpu_check.py

Checks, every rank against NCCL-reduced single-kernel results and rank 0 against the float64 oracle:
  1. peer exchange all-reduce == NCCL all-reduce, bit for bit, for flag-in-data lengths (9, 74, 514, 1024) and
     fence/flag lengths (1025, 8977), repeated (sequence/parity handling)
  2. one-kernel step with both all-reduces inside the kernel (d=9, K=8), ragged shards, an empty shard
  3. two-kernel sharded step for a tiled shape (d=32, K=16: K*d+2 = 514 statistics > one CTA's width)
  4. sharded KMeans (global k-means++ picks, all-reduced stop rule) == the single-process run on the gathered set
  5. sharded GaussianMixture.fit (graph-captured EM iteration) == the single-GPU fit
Prints one line per check on rank 0 and exits non-zero on any mismatch.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from spectrogram_cube_clustering_b200 import ops, synth
from spectrogram_cube_clustering_b200.latent_buffer import LatentBuffer, PeerExchange, shard_bounds
from spectrogram_cube_clustering_b200.models import KMeans, GaussianMixture

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
group = dist.group.WORLD


def say(msg):
    if rank == 0:
        print(msg, flush=True)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# ---------------------------------------------------------------- 1. exchange vs NCCL
ex = PeerExchange(group, dev, 8977)
for length in (9, 74, 514, 1024, 1025, 8977):
    for rep in range(5):
        g = torch.Generator(device=dev).manual_seed(1000 * length + 10 * rep + rank)
        t = torch.randn(length, dtype=torch.float64, device=dev, generator=g)
        # a sum in RANK ORDER is what the exchange promises; NCCL's order may differ in the last bit
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        ref = torch.zeros_like(t)
        for p in parts:
            ref += p
        got = ex.all_reduce(t.clone())
        torch.cuda.synchronize()
        assert torch.equal(got, ref), f"exchange len={length} rep={rep}: max diff {(got - ref).abs().max().item():.3e}"
say(f"[1] peer exchange == rank-ordered sum, bit for bit (world={world}; LL lengths 9..1024, fence lengths 1025, 8977)")

# ---------------------------------------------------------------- 2. one-kernel step, exchanges inside
from oracle import dec as odec            # checker only
for n_total, tag in ((200_003, "ragged"), (world - 1, "empty shards")):
    d, K = 9, 8
    z_all, mu = synth.latent_points(max(n_total, 1), d, K, rank=5, device=dev)
    z_all = z_all[:n_total]
    lo, hi = shard_bounds(n_total, rank, world)
    z = z_all[lo:hi].clone()            # a row slice is a view at a 36-byte offset: the kernels need 16-byte alignment
    n = hi - lo
    scale = 1e-3 / max(n_total, 1)
    out = ops.dec_step(z, mu, 1.0, 5, scale, exchange=ex.desc)
    torch.cuda.synchronize()
    # reference: the two stand-alone kernels with NCCL-reduced statistics
    _, _, f = ops.dec_assign(z, mu, 1.0, 5, want_q=False, want_labels=False)
    dist.all_reduce(f)
    st, p, dz = ops.dec_target_kl_grad(z, mu, f, 1.0, 5, scale)
    dist.all_reduce(st)
    assert rel(out["f"].cpu(), f.cpu()) < 1e-6, (tag, "f")
    if n_total > world:
        assert rel(out["stats"].cpu(), st.cpu()) < 1e-4, (tag, "stats", rel(out["stats"].cpu(), st.cpu()))
        if n:
            assert (out["p"] - p).abs().max().item() <= 1.01e-5
    # replicas identical
    gathered = [torch.empty_like(out["stats"]) for _ in range(world)]
    dist.all_gather(gathered, out["stats"])
    assert all(torch.equal(gathered[0], g_) for g_ in gathered), (tag, "replicas differ")
    if rank == 0 and n_total > world:
        ref = odec.dec_step_chunked(z_all.cpu().numpy(), mu.cpu().numpy(), 1.0, 1e-3, 5)
        s = out["stats"].cpu().numpy()
        assert abs(s[0] - ref["loss"]) < 1e-5 * abs(ref["loss"]) and rel(s[2:].reshape(K, d), ref["dmu"]) < 1e-5
        assert rel(out["f"][:K].cpu().numpy(), ref["f"]) < 1e-5
say("[2] dec_step_ex (f and gradient statistics all-reduced inside the kernel) == NCCL chain == oracle; replicas bit-identical")

# ---------------------------------------------------------------- 3. tiled shape through the latent buffer
d, K, n_total = 32, 16, 120_001
z_all, mu = synth.latent_points(n_total, d, K, rank=6, device=dev)
lo, hi = shard_bounds(n_total, rank, world)
buf = LatentBuffer(z_all[lo:hi], n_total=n_total, group=group, exchange=ex)
res = buf.dec_step(mu, 1.0, 1e-3, 0, want_dz=True)
torch.cuda.synchronize()
if rank == 0:
    ref = odec.dec_step_chunked(z_all.cpu().numpy(), mu.cpu().numpy(), 1.0, 1e-3, None, chunk=20_000)
    assert abs(res.loss.item() - ref["loss"]) < 1e-5 * abs(ref["loss"])
    assert rel(res.dmu.cpu().numpy(), ref["dmu"]) < 1e-5
    assert rel(res.dz.cpu().numpy(), ref["dz"][lo:hi]) < 1e-5
    assert rel(res.f.cpu().numpy(), ref["f"]) < 1e-5
say("[3] sharded step at d=32, K=16 (514 statistics pushed by the tiled kernel's last CTA) == oracle")

# ---------------------------------------------------------------- 4. sharded KMeans
d, K, n_total = 9, 6, 40_000
z_all, _ = synth.latent_points(n_total, d, K, rank=8, device=dev)
lo, hi = shard_bounds(n_total, rank, world)
buf = LatentBuffer(z_all[lo:hi], n_total=n_total, group=group, exchange=ex)
km = KMeans(K, n_init=4, random_state=3, max_iter=200).fit(buf)
km1 = KMeans(K, n_init=4, random_state=3, max_iter=200).fit(LatentBuffer(z_all))
c_all = [torch.empty_like(km._centers) for _ in range(world)]
dist.all_gather(c_all, km._centers)
assert all(torch.equal(c_all[0], c) for c in c_all), "KMeans: ranks hold different centres"
assert rel(km.cluster_centers_, km1.cluster_centers_) < 1e-4, rel(km.cluster_centers_, km1.cluster_centers_)
assert abs(km.inertia_ - km1.inertia_) < 1e-5 * km1.inertia_
assert np.array_equal(km.labels_, km1.labels_[lo:hi]) or (km.labels_ != km1.labels_[lo:hi]).mean() < 1e-3
say(f"[4] sharded KMeans == single-GPU KMeans on the gathered set (inertia {km.inertia_:.6e}); replicas bit-identical")

# ---------------------------------------------------------------- 5. sharded GMM fit
d, K, n_total = 9, 8, 300_000
z_all, _ = synth.latent_points(n_total, d, K, rank=9, device=dev)
lo, hi = shard_bounds(n_total, rank, world)
w0, mu0, cov0 = [t.numpy() for t in synth.gmm_initial_state(d, K, "cpu")]
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    buf = LatentBuffer(z_all[lo:hi], n_total=n_total, group=group, exchange=ex)
    gm = GaussianMixture(K, max_iter=30, tol=1e-3, weights_init=w0, means_init=mu0, covariances_init=cov0,
                         poll_interval=5, group=group).fit(buf)
    gm1 = GaussianMixture(K, max_iter=30, tol=1e-3, weights_init=w0, means_init=mu0, covariances_init=cov0,
                          poll_interval=5).fit(LatentBuffer(z_all))
assert gm.n_iter_ == gm1.n_iter_ and gm.converged_ == gm1.converged_, (gm.n_iter_, gm1.n_iter_)
assert rel(gm.means_, gm1.means_) < 1e-6 and rel(gm.covariances_, gm1.covariances_) < 1e-6
assert abs(gm.lower_bound_ - gm1.lower_bound_) < 1e-5 * abs(gm1.lower_bound_)   # fp32 per-thread partial sums regroup with the shards
m_all = [torch.empty_like(gm._means) for _ in range(world)]
dist.all_gather(m_all, gm._means)
assert all(torch.equal(m_all[0], m) for m in m_all), "GMM: ranks hold different means"
say(f"[5] sharded GaussianMixture.fit ({gm.n_iter_} iterations, graph-captured) == single-GPU fit; replicas bit-identical")

dist.barrier()
say("multigpu_check: all checks passed")
os._exit(0)
