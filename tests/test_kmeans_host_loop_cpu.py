"""CPU tests of the HOST side of models.KMeans: batched restarts, device-side convergence flags, the global
k-means++ draw over a sharded latent set and the all-reduced stop threshold (every rank must take the same
decisions or the collectives mismatch).  CUDA launches replaced by oracle-backed fakes — test-only injection."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _inject():
    import fake_ops
    import spectrogram_cube_clustering_b200.latent_buffer as lb
    import spectrogram_cube_clustering_b200.models as models
    lb.ops = fake_ops
    models.ops = fake_ops
    return lb, models


@pytest.fixture(autouse=True)
def _restore_real_ops():
    import spectrogram_cube_clustering_b200.latent_buffer as lb
    import spectrogram_cube_clustering_b200.models as models
    saved = (lb.ops, models.ops)
    yield
    lb.ops, models.ops = saved


def _blobs(n=1200, d=3, K=4, seed=2):
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(K, d)) * 6.0
    return (centres[rng.integers(0, K, n)] + rng.normal(size=(n, d))).astype(np.float32), centres


def test_lloyd_from_given_centres_matches_sklearn():
    from sklearn.cluster import KMeans as SkKMeans
    lb, models = _inject()
    z, centres = _blobs()
    init = (centres + 0.7).astype(np.float32)
    sk = SkKMeans(4, init=init, n_init=1, max_iter=300, tol=1e-4, algorithm="lloyd").fit(z.astype(np.float64))
    km = models.KMeans(4, max_iter=300, n_init=1, tol=1e-4).fit(lb.LatentBuffer(torch.from_numpy(z)), init_centers=init)
    np.testing.assert_allclose(km.cluster_centers_, sk.cluster_centers_, rtol=1e-4, atol=1e-4)
    assert abs(km.inertia_ - sk.inertia_) < 1e-4 * sk.inertia_
    assert (km.labels_ != sk.labels_).mean() < 1e-3


def test_restarts_pick_the_best_inertia_and_are_reproducible():
    lb, models = _inject()
    z, _ = _blobs()
    buf = lb.LatentBuffer(torch.from_numpy(z))
    a = models.KMeans(4, n_init=6, random_state=11, restart_block=4).fit(buf)      # two blocks of restarts
    b = models.KMeans(4, n_init=6, random_state=11, restart_block=4).fit(buf)
    one = models.KMeans(4, n_init=1, random_state=11).fit(buf)
    assert np.array_equal(a.cluster_centers_, b.cluster_centers_) and a.inertia_ == b.inertia_
    assert a.inertia_ <= one.inertia_ + 1e-9
    with pytest.raises(ValueError):
        models.KMeans(5000).fit(buf)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lb, models = _inject()
    z, _ = _blobs()
    lo, hi = lb.shard_bounds(len(z), rank, world)
    buf = lb.LatentBuffer(torch.from_numpy(z[lo:hi]).clone(), group=dist.group.WORLD)
    km = models.KMeans(4, n_init=3, random_state=7, max_iter=200).fit(buf)
    np.savez(os.path.join(out_dir, f"km{rank}.npz"), centers=km.cluster_centers_, inertia=km.inertia_,
             labels=km.labels_, n_iter=km.n_iter_, lo=lo, hi=hi)
    dist.destroy_process_group()


def test_two_rank_sharded_kmeans_equals_single_process(tmp_path):
    """Global k-means++ picks (one owner rank per draw), all-reduced statistics and stop threshold: both ranks end
    with identical centres, equal to the single-process run on the whole set."""
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"km{r}.npz") for r in range(2))
    assert np.array_equal(r0["centers"], r1["centers"]) and float(r0["inertia"]) == float(r1["inertia"])
    assert int(r0["n_iter"]) == int(r1["n_iter"])
    lb, models = _inject()
    z, _ = _blobs()
    one = models.KMeans(4, n_init=3, random_state=7, max_iter=200).fit(lb.LatentBuffer(torch.from_numpy(z)))
    np.testing.assert_allclose(r0["centers"], one.cluster_centers_, rtol=1e-5, atol=1e-5)
    assert abs(float(r0["inertia"]) - one.inertia_) < 1e-6 * one.inertia_
    labels = np.concatenate([r0["labels"], r1["labels"]])
    assert np.array_equal(labels, one.labels_)
