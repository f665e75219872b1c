"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row sharding, packed float64
statistics, one all-reduce per pass, identical results on every rank.  The CUDA launches are
replaced by oracle-backed fakes (tests/fake_ops.py) — test-only injection; the product has no
CPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectrogram_cube_clustering_b200.latent_buffer import shard_bounds, update_interval


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, K, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fake_ops
    import spectrogram_cube_clustering_b200.latent_buffer as lb
    from spectrogram_cube_clustering_b200 import synth
    from oracle import dec as odec
    lb.ops = fake_ops                                   # test-only injection
    z, mu = synth.latent_points(n, d, K, rank=3)
    lo, hi = shard_bounds(n, rank, world)
    buf = lb.LatentBuffer(z[lo:hi].clone(), group=dist.group.WORLD)
    assert buf.n_total == n and buf.n_local == hi - lo
    res = buf.dec_step(mu, 1.0, 1e-3, round_decimals=5)
    ref = odec.dec_step(z.numpy(), mu.numpy(), 1.0, 1e-3, round_to=5)
    np.testing.assert_allclose(res.f.numpy(), ref["f"], rtol=1e-12)
    np.testing.assert_allclose(res.dmu.numpy(), ref["dmu"], rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(float(res.loss), ref["loss"], rtol=1e-10)
    # same step with the shard's target rows kept (one-pass target + gradient launch)
    res_p = buf.dec_step(mu, 1.0, 1e-3, round_decimals=5, want_p=True)
    np.testing.assert_allclose(res_p.p.numpy(), ref["p"][lo:hi], atol=1e-7)
    assert torch.equal(res_p.dmu, res.dmu)
    buf.labels = None                                   # the extra pass above must not count as a label change
    buf.dec_assign(mu, 1.0)
    # label-change counter: second pass with moved centroids, summed over ranks
    mu2 = mu.clone(); mu2[0] += 0.5
    _, st = buf.dec_assign(mu2, 1.0)
    lab2 = odec.labels_from_q(odec.soft_assign(z.numpy(), mu2.numpy()))
    assert float(st[-1]) == float((lab2 != ref["labels"]).sum())
    assert abs(buf.delta_label(st) - odec.delta_label(lab2, ref["labels"])) < 1e-7
    # GMM statistics pass: packed buffer all-reduced once
    from oracle import gmm as ogmm
    w0, mu0, cov0 = [t.numpy() for t in synth.gmm_initial_state(d, K)]
    pchol = ogmm.precision_cholesky(cov0)
    tri = d * (d + 1) // 2
    params = np.concatenate([mu0.ravel(),
                             np.concatenate([[pchol[k][a, b] for b in range(d) for a in range(b + 1)] for k in range(K)]),
                             ogmm.log_det_cholesky(pchol) + np.log(w0) - 0.5 * d * np.log(2 * np.pi)])
    stats = buf.gmm_em_pass(K, torch.from_numpy(params.astype(np.float32)))
    lb_ref, log_resp = ogmm.e_step(z.numpy().astype(np.float64), w0, mu0, pchol)
    assert stats.numel() == 1 + K + K * d + K * tri
    np.testing.assert_allclose(float(stats[0]) / n, lb_ref, rtol=1e-6)
    np.testing.assert_allclose(stats[1:1 + K].numpy(), np.exp(log_resp).sum(0), rtol=1e-5)
    torch.save(dict(dmu=res.dmu, f=res.f, stats=stats), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_latent_buffer_gloo(tmp_path):
    world, n, d, K = 2, 1001, 9, 8                     # odd n: ragged shards
    mp.spawn(_worker, args=(world, _free_port(), n, d, K, str(tmp_path)), nprocs=world, join=True)
    a, b = (torch.load(tmp_path / f"rank{r}.pt") for r in range(world))
    for key in a:                                      # replicas end bit-identical
        assert torch.equal(a[key], b[key])


@pytest.mark.parametrize("n,world", [(10, 3), (1_000_000, 8), (7, 8), (0, 2)])
def test_shard_bounds_partition(n, world):
    spans = [shard_bounds(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a1 >= a0
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def test_update_interval_matches_reference():
    # models.py:985-989
    assert update_interval(1000, 16, -1) == int(np.ceil(1000 / 32))
    assert update_interval(1000, 16, 4) == int(np.ceil(1000 / 64))
