"""CPU tests of the HOST side of GaussianMixture.fit (models.py): polling in chunks, device-side stop rule read
back through the control block, frozen iterations, state reuse between fits, sharded fit over gloo.  The CUDA
launches are replaced by oracle-backed fakes (tests/fake_ops.py) — test-only injection; the product has no CPU path."""
import os
import socket
import sys
import warnings

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
    """The injection is per test: later tests of the session see the real (CUDA-only) ops again."""
    import spectrogram_cube_clustering_b200.latent_buffer as lb
    import spectrogram_cube_clustering_b200.models as models
    saved = (lb.ops, models.ops)
    yield
    lb.ops, models.ops = saved


def _problem(n=1500, d=3, K=3, seed=5):
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(K, d)) * 4.0
    z = (centres[rng.integers(0, K, n)] + rng.normal(size=(n, d))).astype(np.float32)
    w0 = np.full(K, 1.0 / K)
    mu0 = centres + rng.normal(size=(K, d)) * 0.5
    cov0 = np.stack([np.eye(d)] * K)
    return z, w0, mu0, cov0


@pytest.mark.parametrize("poll", [1, 3, 50])
def test_fit_host_loop_matches_sklearn(poll):
    from sklearn.mixture import GaussianMixture as SkGMM
    lb, models = _inject()
    z, w0, mu0, cov0 = _problem()
    sk = SkGMM(3, covariance_type="full", tol=1e-4, max_iter=100, weights_init=w0, means_init=mu0,
               precisions_init=np.linalg.inv(cov0)).fit(z.astype(np.float64))
    gm = models.GaussianMixture(3, max_iter=100, tol=1e-4, weights_init=w0, means_init=mu0, covariances_init=cov0,
                                poll_interval=poll)
    buf = lb.LatentBuffer(torch.from_numpy(z))
    gm.fit(buf)
    # the stop is decided by the (fake) device; iterations launched after it within a poll chunk are no-ops
    assert gm.n_iter_ == sk.n_iter_ and gm.converged_ == sk.converged_
    np.testing.assert_allclose(gm.means_, sk.means_, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gm.weights_, sk.weights_, rtol=1e-5)
    np.testing.assert_allclose(gm.lower_bound_, sk.lower_bound_, rtol=1e-6)
    labels = gm.predict()
    assert (labels != sk.predict(z.astype(np.float64))).mean() < 2e-3
    # second fit of the same buffer: same state tensors (no reallocation), same answer
    ptr = gm._means.data_ptr()
    first = (gm.means_.copy(), gm.n_iter_)
    gm.fit(buf)
    assert gm._means.data_ptr() == ptr and gm.n_iter_ == first[1] and np.array_equal(gm.means_, first[0])


def test_fit_warns_when_max_iter_is_hit_and_raises_on_bad_covariance():
    lb, models = _inject()
    z, w0, mu0, cov0 = _problem()
    gm = models.GaussianMixture(3, max_iter=2, tol=0.0, weights_init=w0, means_init=mu0, covariances_init=cov0)
    with pytest.warns(models.ConvergenceWarning):
        gm.fit(lb.LatentBuffer(torch.from_numpy(z)))
    assert gm.n_iter_ == 2 and not gm.converged_
    bad = cov0.copy(); bad[1] = -np.eye(3)
    with pytest.raises(ValueError, match="ill-defined empirical covariance"):
        models.GaussianMixture(3, weights_init=w0, means_init=mu0, covariances_init=bad).fit(
            lb.LatentBuffer(torch.from_numpy(z)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lb, models = _inject()
    z, w0, mu0, cov0 = _problem()
    lo, hi = lb.shard_bounds(len(z), rank, world)
    buf = lb.LatentBuffer(torch.from_numpy(z[lo:hi]).clone(), group=dist.group.WORLD)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm = models.GaussianMixture(3, max_iter=100, tol=1e-4, weights_init=w0, means_init=mu0, covariances_init=cov0,
                                    poll_interval=4, group=dist.group.WORLD).fit(buf)
    np.savez(os.path.join(out_dir, f"gm{rank}.npz"), means=gm.means_, cov=gm.covariances_, n_iter=gm.n_iter_,
             lower=gm.lower_bound_)
    dist.destroy_process_group()


def test_two_rank_sharded_fit_equals_single_process(tmp_path):
    """No peer exchange on CPU: statistics pass -> group all_reduce -> finalize on every rank (latent_buffer.py)."""
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"gm{r}.npz") for r in range(2))
    assert np.array_equal(r0["means"], r1["means"]) and int(r0["n_iter"]) == int(r1["n_iter"])
    lb, models = _inject()
    z, w0, mu0, cov0 = _problem()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        one = models.GaussianMixture(3, max_iter=100, tol=1e-4, weights_init=w0, means_init=mu0,
                                     covariances_init=cov0).fit(lb.LatentBuffer(torch.from_numpy(z)))
    assert one.n_iter_ == int(r0["n_iter"])
    np.testing.assert_allclose(r0["means"], one.means_, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(r0["cov"], one.covariances_, rtol=1e-8, atol=1e-12)
