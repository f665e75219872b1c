"""GPU tests of the reference-facing Python surface (networks.ClusteringLayer / DEC,
models.target_distribution / gmm / kmeans / GaussianMixture, LatentBuffer)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, DEC_CASES

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("case", DEC_CASES)
def test_clustering_layer_matches_reference_layer(case):
    """Same constructor / forward / autograd contract as Cluster.networks.ClusteringLayer, run the
    way the reference runs it (float64 module), against the reference's own outputs."""
    from spectrogram_cube_clustering_b200.networks import ClusteringLayer
    g = load_golden("dec", case)
    n, d = g["z"].shape
    K = g["mu"].shape[0]
    layer = ClusteringLayer(K, d, float(g["alpha"]), weights=torch.from_numpy(g["mu"]).double()).double().cuda()
    assert layer.weights.dtype == torch.float64 and tuple(layer.weights.shape) == (K, d)
    z = torch.from_numpy(g["z"]).double().cuda().requires_grad_(True)
    q = layer(z)
    assert q.dtype == torch.float64 and tuple(q.shape) == (n, K)
    assert rel_err(q.detach().cpu().numpy(), g["q"]) < TOL
    # the literal reference loss line: gamma * KLDivLoss('sum')(log q, p) / B   (models.py:1124-1125)
    tar = torch.from_numpy(g["p"]).cuda()
    loss = float(g["gamma"]) * torch.nn.KLDivLoss(reduction="sum")(torch.log(q), tar) / n
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 2 * TOL * abs(float(g["loss"]))
    assert rel_err(z.grad.cpu().numpy(), g["dz"]) < 2 * TOL
    assert rel_err(layer.weights.grad.cpu().numpy(), g["dmu"]) < 2 * TOL


def test_clustering_layer_generic_grad_and_padding():
    """d = 7 is not instantiated: the host zero-pads z and the centroids (exact)."""
    from spectrogram_cube_clustering_b200.networks import ClusteringLayer
    from oracle import dec as odec
    rng = np.random.default_rng(3)
    z = rng.normal(size=(333, 7)).astype(np.float32)
    mu = rng.normal(size=(5, 7)).astype(np.float32)
    G = rng.normal(size=(333, 5)).astype(np.float32)
    layer = ClusteringLayer(5, 7, 1.0, weights=torch.from_numpy(mu)).cuda()
    zt = torch.from_numpy(z).cuda().requires_grad_(True)
    q = layer(zt)
    q.backward(torch.from_numpy(G).cuda())
    dz, dmu = odec.backward_generic(z, mu, G, 1.0)
    assert rel_err(q.detach().cpu().numpy(), odec.soft_assign(z, mu)) < TOL
    assert rel_err(zt.grad.cpu().numpy(), dz) < TOL and rel_err(layer.weights.grad.cpu().numpy(), dmu) < TOL


def test_clustering_layer_rejects_cpu():
    from spectrogram_cube_clustering_b200.networks import ClusteringLayer
    from spectrogram_cube_clustering_b200._lib import SccError
    with pytest.raises(SccError):
        ClusteringLayer(8, 9)(torch.zeros(4, 9))


def test_dec_module_state_dict_and_forward():
    from spectrogram_cube_clustering_b200.networks import DEC
    from spectrogram_cube_clustering_b200 import synth
    model = DEC(n_clusters=5).cuda()
    keys = set(model.state_dict())
    assert "clustering.weights" in keys and "encoder.encoder.8.weight" in keys and "decoder.decoder.0.weight" in keys
    assert "encoder.encoder.6.conv.weight" in keys
    x = synth.spectrograms(32, device="cuda")
    q, x_rec, z = model(x)
    assert tuple(q.shape) == (32, 5) and tuple(z.shape) == (32, 9) and tuple(x_rec.shape) == (32, 1, 4, 101)
    torch.testing.assert_close(q.sum(1), torch.ones(32, device="cuda"), atol=1e-5, rtol=0)
    (q[:, 0].sum() + x_rec.mean()).backward()
    assert model.clustering.weights.grad is not None and model.encoder.encoder[0].weight.grad is not None


@pytest.mark.parametrize("case", ["c1", "k5", "tie"])
def test_target_distribution_numpy_api(case):
    from spectrogram_cube_clustering_b200.models import target_distribution
    g = load_golden("dec", case)
    p = target_distribution(g["q_round"])
    assert isinstance(p, np.ndarray) and p.dtype == np.float64 and p.shape == g["p"].shape
    d = np.abs(p - g["p"])
    assert d.max() <= 1.01e-5 and (d > 1e-7).mean() < 0.03


@pytest.mark.parametrize("case", ["c1", "k16", "relu"])
def test_gaussian_mixture_front_end_matches_sklearn_fit(case):
    from spectrogram_cube_clustering_b200.models import GaussianMixture
    g = load_golden("gmm", case)
    gm = GaussianMixture(g["mu0"].shape[0], max_iter=100, tol=float(g["tol"]), weights_init=g["w0"],
                         means_init=g["mu0"], covariances_init=g["cov0"], poll_interval=4)
    labels = gm.fit_predict(g["z"])
    assert gm.n_iter_ == int(g["fit_n_iter"]) and gm.converged_ == bool(g["fit_converged"])
    assert abs(gm.lower_bound_ - float(g["fit_lower_bound"])) < 1e-4 * abs(float(g["fit_lower_bound"]))
    assert rel_err(gm.means_, g["fit_means"]) < 1e-4 and rel_err(gm.weights_, g["fit_weights"]) < 1e-4
    assert labels.dtype == np.int64 and (labels != g["fit_labels"]).mean() < 2e-3


def test_gaussian_mixture_reuses_its_graph_on_the_same_buffer():
    """A second fit of the same LatentBuffer replays the EM-iteration graph captured by the first (same state
    tensors, same kernel arguments) and gives the same result; another buffer or tolerance recaptures."""
    import warnings
    from spectrogram_cube_clustering_b200 import synth
    from spectrogram_cube_clustering_b200.latent_buffer import LatentBuffer
    from spectrogram_cube_clustering_b200.models import GaussianMixture
    d, K = 9, 4
    z, _ = synth.latent_points(20_000, d, K, rank=3, device="cuda")
    w0, mu0, cov0 = [t.numpy() for t in synth.gmm_initial_state(d, K, "cpu")]
    buf = LatentBuffer(z)
    gm = GaussianMixture(K, max_iter=12, tol=0.0, weights_init=w0, means_init=mu0, covariances_init=cov0, poll_interval=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm.fit(buf)
        g1, first = gm._graph, (gm.means_.copy(), gm.covariances_.copy(), gm.lower_bound_, gm.n_iter_)
        gm.fit(buf)
        assert g1 is not None and gm._graph is g1
        assert np.array_equal(gm.means_, first[0]) and np.array_equal(gm.covariances_, first[1])
        assert gm.lower_bound_ == first[2] and gm.n_iter_ == first[3] == 12
        gm.fit(LatentBuffer(z.clone()))
        assert gm._graph is not g1 and np.array_equal(gm.means_, first[0])
        g2 = gm._graph
        gm.tol = 1e-3
        gm.fit(gm._buf)
        assert gm._graph is not g2


def test_gaussian_mixture_errors():
    from spectrogram_cube_clustering_b200.models import GaussianMixture, ConvergenceWarning
    with pytest.raises(ValueError):                       # n_samples < n_components (sklearn _base.py:232-237)
        GaussianMixture(8).fit(np.zeros((4, 9), dtype=np.float32))
    z = np.ones((256, 9), dtype=np.float32)               # collapsed data, no regularisation -> not PD
    with pytest.raises(ValueError):
        GaussianMixture(2, reg_covar=0.0, means_init=np.ones((2, 9)),
                        covariances_init=np.tile(np.eye(9), (2, 1, 1))).fit(z)
    g = load_golden("gmm", "c1")
    with pytest.warns(ConvergenceWarning):
        GaussianMixture(8, max_iter=2, tol=0.0, weights_init=g["w0"], means_init=g["mu0"],
                        covariances_init=g["cov0"]).fit(g["z"])


def test_kmeans_lloyd_matches_sklearn_from_identical_centres():
    from sklearn.cluster import KMeans as SkKMeans
    from spectrogram_cube_clustering_b200.models import KMeans
    g = load_golden("gmm", "c1")
    z = g["z"]
    init = g["mu0"].astype(np.float32)
    # tol = 0: both run to strict convergence (no label changes), so the stop iteration cannot differ
    sk = SkKMeans(n_clusters=8, init=init, n_init=1, max_iter=500, tol=0.0, algorithm="lloyd").fit(z.astype(np.float64))
    km = KMeans(8, max_iter=500, n_init=1, tol=0.0).fit(z, init_centers=init)
    assert (km.labels_ != sk.labels_).mean() < 5e-3
    assert abs(km.inertia_ - sk.inertia_) < 1e-4 * sk.inertia_
    assert rel_err(km.cluster_centers_, sk.cluster_centers_) < 1e-3


def test_gmm_function_end_to_end():
    """models.gmm(z, K): k-means++ seeding + EM, vs scikit-learn EM from the SAME seeds."""
    from sklearn.mixture import GaussianMixture as SkGM
    from spectrogram_cube_clustering_b200.models import gmm, KMeans
    from spectrogram_cube_clustering_b200 import synth
    from oracle import gmm as ogmm
    z, _ = synth.latent_points(6000, 9, 5, rank=42)
    z = z.numpy()
    km = KMeans(5, n_init=3, random_state=2009).fit(z)
    w0 = np.bincount(km.labels_, minlength=5) / z.shape[0]
    labels, cent = gmm(z, 5, means_init=km.cluster_centers_, weights_init=w0)
    assert labels.shape == (6000,) and labels.dtype == np.int64 and cent.shape == (5, 9)
    X = z.astype(np.float64)
    lab0 = np.argmin(((X[:, None] - km.cluster_centers_[None]) ** 2).sum(2), 1)
    resp = np.zeros((6000, 5)); resp[np.arange(6000), lab0] = 1
    with np.errstate(divide="ignore"):
        _, _, cov0, _, _ = ogmm.m_step(X, np.log(resp))
    sk = SkGM(5, max_iter=1000, weights_init=w0, means_init=km.cluster_centers_,
              precisions_init=np.linalg.inv(cov0)).fit(X)
    assert rel_err(cent, sk.means_) < 1e-3
    assert (labels != sk.predict(X)).mean() < 5e-3
    labels2, cent2 = gmm(z, 5)                      # full path with its own KMeans(n_init=100) seeding
    assert labels2.shape == (6000,) and np.isfinite(cent2).all() and len(np.unique(labels2)) == 5


def test_latent_buffer_dec_step_and_refine():
    from spectrogram_cube_clustering_b200.latent_buffer import LatentBuffer
    from spectrogram_cube_clustering_b200.models import dec_refine
    from spectrogram_cube_clustering_b200 import synth
    from oracle import dec as odec
    z, mu = synth.latent_points(20000, 9, 8, rank=5)
    buf = LatentBuffer.from_host(z.numpy(), "cuda")
    res = buf.dec_step(mu.cuda(), 1.0, 1e-3, round_decimals=0)
    ref = odec.dec_step(z.numpy(), mu.numpy(), 1.0, 1e-3, round_to=None)
    assert abs(res.loss.item() - ref["loss"]) < TOL * abs(ref["loss"])
    assert rel_err(res.dmu.cpu().numpy(), ref["dmu"]) < TOL and rel_err(res.f.cpu().numpy(), ref["f"]) < TOL
    res_p = buf.dec_step(mu.cuda(), 1.0, 1e-3, round_decimals=0, want_p=True)      # same pass, target rows kept
    assert rel_err(res_p.p.cpu().numpy(), ref["p"]) < 2 * TOL and torch.equal(res_p.dmu, res.dmu)
    cent, hist = dec_refine(buf, mu, lr=1e-2, max_steps=30, tol=0.0)
    assert cent.shape == (8, 9) and len(hist) == 30 and np.isfinite(cent).all()
    assert hist[-1]["delta"] <= hist[1]["delta"] + 1e-3


def test_batch_eval_matches_layerwise_eval():
    from spectrogram_cube_clustering_b200.networks import DEC
    from spectrogram_cube_clustering_b200.models import batch_eval
    from spectrogram_cube_clustering_b200 import synth
    torch.manual_seed(0)
    model = DEC(n_clusters=5).cuda()
    x = synth.spectrograms(1000)
    loader = torch.utils.data.DataLoader(x, batch_size=128, shuffle=False)
    q, labels, z = batch_eval(loader, model, "cuda")
    assert q.shape == (1000, 5) and labels.shape == (1000,) and z.shape == (1000, 9)
    with torch.no_grad():                 # same batch split: cuDNN picks algorithms per batch size
        outs = [model(xb.cuda()) for xb in loader]
    qq = torch.cat([o[0] for o in outs]); zz = torch.cat([o[2] for o in outs])
    assert rel_err(z, zz.cpu().numpy()) < 1e-6
    assert np.abs(q - np.round(qq.cpu().numpy().astype(np.float64), 5)).max() <= 1.01e-5
    assert (labels != qq.argmax(1).cpu().numpy()).mean() < 2e-3


def _reference_style_epoch(model, x, bsz, gamma, opt):
    """The reference loop verbatim in torch (models.py:1015-1016, 1089-1128) for one epoch without
    p refreshes: q via the layer, p on the host with numpy, literal KLDivLoss line."""
    from oracle import dec as odec
    dev = next(model.parameters()).device
    model.eval()
    with torch.no_grad():
        q = torch.cat([model(x[i:i + bsz].to(dev))[0] for i in range(0, len(x), bsz)]).cpu().numpy()
    p = odec.target_distribution(np.round(q.astype(np.float64), 5))
    out = []
    for i in range(0, len(x), bsz):
        xb = x[i:i + bsz].to(dev)
        tar = torch.from_numpy(p[i:i + bsz]).to(dev, torch.float32)
        model.train(); opt.zero_grad()
        qb, x_rec, _ = model(xb)
        loss = torch.nn.MSELoss()(x_rec, xb) + gamma * torch.nn.KLDivLoss(reduction="sum")(torch.log(qb), tar) / xb.shape[0]
        loss.backward(); opt.step()
        out.append(loss.item())
    return out


@pytest.mark.parametrize("fused", [True, False])
def test_dec_training_epoch_matches_reference_style_loop(fused):
    import copy
    from spectrogram_cube_clustering_b200.networks import DEC
    from spectrogram_cube_clustering_b200.models import DEC_training
    from spectrogram_cube_clustering_b200 import synth
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(1)
        model = DEC(n_clusters=5).cuda()
        with torch.no_grad():
            model.clustering.weights.mul_(0.3).add_(0.1)
        ref_model = copy.deepcopy(model)
        x = synth.spectrograms(2048)
        bsz, gamma = 256, 1e-1
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        ref_opt = torch.optim.Adam(ref_model.parameters(), lr=1e-3)
        loader = torch.utils.data.DataLoader(x, batch_size=bsz, shuffle=False)
        hist = DEC_training(model, loader, opt, n_epochs=1, gamma=gamma, tol=0.0, update_interval_cfg=1,
                            fused_loss=fused)           # update_interval = M/B batches: no refresh inside the epoch
        ref_losses = _reference_style_epoch(ref_model, x, bsz, gamma, ref_opt)
        assert hist["update_interval"] == 8 and len(hist["loss"]) == 1
        assert abs(hist["loss"][0] - np.mean(ref_losses)) < 2e-4 * abs(np.mean(ref_losses))
        w, wr = model.clustering.weights.detach().cpu().numpy(), ref_model.clustering.weights.detach().cpu().numpy()
        assert rel_err(w, wr) < 2e-3                       # Adam amplifies 1e-6 gradient differences
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_fused_kl_loss_matches_literal_line():
    from spectrogram_cube_clustering_b200.networks import ClusteringLayer, dec_kl_loss
    g = load_golden("dec", "c1")
    n = g["z"].shape[0]
    z = torch.from_numpy(g["z"]).cuda().requires_grad_(True)
    w = torch.nn.Parameter(torch.from_numpy(g["mu"]).cuda())
    tar = torch.from_numpy(g["p"]).float().cuda()
    loss = dec_kl_loss(z, w, tar, 1.0, float(g["gamma"]) / n) * 3.0
    loss.backward()
    assert abs(loss.item() / 3.0 - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert rel_err(z.grad.cpu().numpy() / 3.0, g["dz"]) < 1e-5 and rel_err(w.grad.cpu().numpy() / 3.0, g["dmu"]) < 1e-5


# ------------------------------------------------------------------------------ SURVEY §8(f) rows
def test_batch_eval_matches_oracle_soft_assignment():
    """f2: q / labels of the device batch_eval against the float64 oracle applied to the encoder's own latents
    (networks.py:279-288, models.py:92-94) — not against this package's layer."""
    from spectrogram_cube_clustering_b200.networks import DEC
    from spectrogram_cube_clustering_b200.models import batch_eval
    from spectrogram_cube_clustering_b200 import synth
    from oracle import dec as odec
    torch.manual_seed(4)
    model = DEC(n_clusters=6).cuda()
    with torch.no_grad():
        model.clustering.weights.mul_(0.5).add_(0.2)
    x = synth.spectrograms(3000)
    loader = torch.utils.data.DataLoader(x, batch_size=256, shuffle=False)
    q, labels, z = batch_eval(loader, model, "cuda")
    assert q.dtype == np.float64 and labels.dtype == np.int64 and z.dtype == np.float64
    mu = model.clustering.weights.detach().cpu().numpy().astype(np.float64)
    q_ref = odec.soft_assign(z, mu, 1.0)                    # z: the latents batch_eval itself returned
    dq = np.abs(q - np.round(q_ref, 5))
    assert dq.max() <= 1.01e-5 and (dq > 1e-7).mean() < 0.01
    lab_ref = odec.labels_from_q(q_ref)
    mism = labels != lab_ref
    if mism.any():
        qs = np.sort(q_ref[mism], axis=1)
        assert np.all(qs[:, -1] - qs[:, -2] < 1e-5)
    # and the latents are the encoder's (same batch split)
    model.eval()
    with torch.no_grad():
        zz = torch.cat([model.encoder(xb.cuda()) for xb in loader]).cpu().numpy()
    assert rel_err(z, zz) < 1e-6
    # device-resident form: buffer + fused label-change count
    prev = torch.from_numpy(lab_ref.astype(np.int32)).cuda()
    prev[:77] = (prev[:77] + 1) % 6
    buf, q_dev, lab_dev = batch_eval(loader, model, "cuda", return_buffer=True, labels_prev=prev)
    assert int(buf.assign_stats[-1].item()) == int((lab_dev != prev).sum().item())
    assert abs(int(buf.assign_stats[-1].item()) - 77) <= int(mism.sum())


def test_dec_training_refresh_and_stop_rule():
    """f3: the loop crosses update_interval boundaries (p refreshed from a device batch_eval) and the
    delta_label < tol stop fires from the fused label-change count (models.py:1093-1111)."""
    import copy
    from spectrogram_cube_clustering_b200.networks import DEC
    from spectrogram_cube_clustering_b200.models import DEC_training, batch_eval
    from spectrogram_cube_clustering_b200 import synth
    torch.manual_seed(2)
    model = DEC(n_clusters=4).cuda()
    with torch.no_grad():
        model.clustering.weights.mul_(0.4).add_(0.15)
    x = synth.spectrograms(1024)
    loader = torch.utils.data.DataLoader(x, batch_size=128, shuffle=False)
    m0 = copy.deepcopy(model)
    opt = torch.optim.Adam(model.parameters(), lr=5e-3)
    # tol = 0: never stops; 2 epochs x 8 batches, update_interval = ceil(1024 / 256) = 4 -> refreshes at
    # (epoch 0, batch 4), (epoch 1, batch 0), (epoch 1, batch 4)
    hist = DEC_training(model, loader, opt, n_epochs=2, gamma=1e-1, tol=0.0)
    assert hist["update_interval"] == 4 and len(hist["deltas"]) == 3 and not hist["finished"]
    assert hist["deltas_iter"] == [5, 9, 13] and len(hist["loss"]) == 2
    assert all(0.0 <= dl <= 1.0 for dl in hist["deltas"])
    # stop rule: the first refresh compares with the labels of the untouched model; tol = 2 fires at once
    model2 = copy.deepcopy(m0)
    opt2 = torch.optim.Adam(model2.parameters(), lr=5e-3)
    _, lab0, _ = batch_eval(loader, m0, "cuda")
    hist2 = DEC_training(model2, loader, opt2, n_epochs=3, gamma=1e-1, tol=2.0)
    assert hist2["finished"] and len(hist2["deltas"]) == 1 and len(hist2["loss"]) == 1
    _, lab1, _ = batch_eval(loader, model2, "cuda")        # the break leaves the model as the refresh saw it
    delta_ref = float((lab1 != lab0).sum()) / len(lab0)    # models.py:1098-1099
    assert abs(hist2["deltas"][0] - delta_ref) <= 2.0 / len(lab0)
    # and with a tolerance that is NOT met the loop goes on
    model3 = copy.deepcopy(m0)
    hist3 = DEC_training(model3, loader, torch.optim.Adam(model3.parameters(), lr=5e-3), n_epochs=1, gamma=1e-1,
                         tol=min(1e-9, hist2["deltas"][0] / 2))
    assert len(hist3["deltas"]) == 1 and hist3["finished"] == (hist3["deltas"][0] < min(1e-9, hist2["deltas"][0] / 2))


def test_initialize_clusters_and_dec_params_round_trip(tmp_path):
    """f4: stage 2 -> stage 3 hand-off on disk (models.py:444-449 -> 523-530 -> 1006-1012, 1227-1228)."""
    import types
    from spectrogram_cube_clustering_b200.networks import DEC
    from spectrogram_cube_clustering_b200 import models, synth
    torch.manual_seed(3)
    K = 4
    model = DEC(n_clusters=K).cuda()
    x = synth.spectrograms(1536)
    loader = torch.utils.data.DataLoader(x, batch_size=256, shuffle=False)
    weights = tmp_path / "AEC" / "AEC_Params_Final.pt"
    weights.parent.mkdir()
    torch.save(model.state_dict(), weights)
    run = tmp_path / "run"
    run.mkdir()
    # stage 2: latent set -> gmm_fit -> labels.npy / centroids.npy where stage 3 looks for them
    gmm_dir = weights.parent / "GMM" / f"n_clusters={K}"
    gmm_dir.mkdir(parents=True)
    z = models.batch_eval(loader, model, "cuda")[2]
    lab_fit, cent_fit = models.gmm_fit(types.SimpleNamespace(savepath_run=str(gmm_dir)), z, K)
    cfg = types.SimpleNamespace(init="load", saved_weights=str(weights), savepath_run=str(run), device="cuda",
                                index_tra=np.arange(1536))
    lab, cent = models.initialize_clusters(model, loader, cfg, n_clusters=K)
    assert np.array_equal(lab, lab_fit) and np.array_equal(cent, cent_fit)
    for init in ("kmeans", "gmm", "rand"):
        cfg_i = types.SimpleNamespace(init=init, saved_weights=str(weights), savepath_run=str(run), device="cuda")
        lab_i, cent_i = models.initialize_clusters(model, loader, cfg_i, n_clusters=K)
        assert lab_i.shape == (1536,) and cent_i.shape == (K, 9) and np.isfinite(cent_i).all()
        assert lab_i.min() >= 0 and lab_i.max() < K
    with pytest.raises(ValueError):
        models.initialize_clusters(model, loader, types.SimpleNamespace(init="nope"), n_clusters=K)
    # stage 3 with the config: centroids copied into clustering.weights, both parameter files written
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    hist = models.DEC_training(model, loader, opt, n_epochs=1, gamma=1e-1, tol=0.0, config=cfg, n_clusters=K)
    init_sd = torch.load(run / "DEC_Params_Initial.pt")
    final_sd = torch.load(run / "DEC_Params_Final.pt")
    assert set(init_sd) == set(model.state_dict()) and "clustering.weights" in init_sd
    np.testing.assert_allclose(init_sd["clustering.weights"].numpy(), cent_fit.astype(np.float32), rtol=1e-6)
    fresh = DEC(n_clusters=K)
    fresh.load_state_dict(final_sd, strict=True)                       # the reference's keys, loadable as they are
    assert torch.equal(fresh.clustering.weights.detach(), model.clustering.weights.detach().cpu())
    assert (run / "DEC_history.csv").exists() and (run / "Delta_history.csv").exists()
    assert hist["params_final"].endswith("DEC_Params_Final.pt")


def test_kmeans_batched_restarts():
    """f1: R restarts advance in one launch per Lloyd iteration; each equals the single-restart run from the
    same centres, converged restarts freeze, and fit() keeps the lowest inertia."""
    from spectrogram_cube_clustering_b200 import ops, synth
    from spectrogram_cube_clustering_b200.models import KMeans
    from spectrogram_cube_clustering_b200.latent_buffer import LatentBuffer
    z, _ = synth.latent_points(20_000, 9, 6, device="cuda", rank=9)
    buf = LatentBuffer(z)
    km = KMeans(6, n_init=5, random_state=7, max_iter=300)
    gen = torch.Generator(device="cpu"); gen.manual_seed(7)
    c0 = km._plusplus(buf, 5, gen)
    assert tuple(c0.shape) == (5, 6, 9)
    # every seeded centre is a row of z
    assert all((z == c0[r, k]).all(dim=1).any().item() for r in range(5) for k in range(6))
    # one batched scan == five single scans
    st_b = ops.kmeans_batch_step(z, c0)
    for r in range(5):
        st_1 = ops.kmeans_step(z, c0[r].contiguous())
        assert torch.equal(st_b[r], st_1)
    tol = km._shift_tol(buf, km.tol)
    assert abs(tol - 1e-4 * z.double().var(dim=0, unbiased=False).mean().item()) < 1e-9 * tol + 1e-15
    batch = c0.clone()
    inertia, n_iter = km._lloyd_batch(buf, batch, tol)
    for r in range(5):
        c_r, in_r, _, it_r = km.lloyd(buf, c0[r])
        assert torch.equal(c_r, batch[r]) and in_r == inertia[r].item() and it_r == int(n_iter[r].item())
    km.fit(buf)
    assert abs(km.inertia_ - inertia.min().item()) <= 1e-9 * km.inertia_
    assert km.labels_.shape == (20_000,) and km.cluster_centers_.shape == (6, 9)
    # skipped restarts keep their statistics
    done = torch.tensor([0, 1, 0, 1, 0], dtype=torch.uint8, device="cuda")
    st_keep = torch.full_like(st_b, -1.0)
    ops.kmeans_batch_step(z, c0, done=done, out_stats=st_keep)
    assert (st_keep[1] == -1).all() and (st_keep[3] == -1).all() and torch.equal(st_keep[0], st_b[0])


@pytest.mark.parametrize("p", [2.0, 1.0, 0.5, 3.0])
def test_distance_scans_match_reference_formulas(p):
    """f4 (second half): utils.fractional_distance / distance_matrix / measure_class_inertia
    (utils.py:866-869, 635-643, 1024-1029) on the device."""
    from spectrogram_cube_clustering_b200 import models
    rng = np.random.default_rng(5)
    z = rng.normal(size=(3001, 9)) * 1.3 + 0.5
    mu = rng.normal(size=(7, 9))
    for j in (0, 6):
        ref = np.sum(np.fabs(mu[j] - z) ** p, axis=1) ** (1 / p)            # utils.py:867-868
        got = models.fractional_distance(mu[j], z, p)
        assert got.dtype == np.float64 and rel_err(got, ref) < 2e-6
    dm = models.distance_matrix(mu, mu, p)
    ref_dm = np.array([[np.sum(np.fabs(mu[i] - mu[j]) ** p) ** (1 / p) for j in range(7)] for i in range(7)])
    assert dm.shape == (7, 7) and np.abs(dm - ref_dm).max() < 2e-6 * ref_dm.max()
    if p == 2.0:
        ref_in = np.array([np.sum(np.sqrt(np.sum((z - mu[j]) ** 2, axis=1)) ** 2) for j in range(7)])
        assert rel_err(models.measure_class_inertia(z, mu, 7), ref_in) < 1e-5
    big = rng.normal(size=(40, 9))
    assert models.distance_matrix(big, big, p).shape == (40, 40)            # more than 16 columns: scanned in blocks


# ------------------------------------------------------------------------------ torch.ops.scc_b200.*
def test_custom_ops_opcheck_and_layer_routes_through_them():
    """The launches are registered PyTorch custom ops (fake kernels + autograd formulas): torch.library.opcheck
    validates schema, fake tensors, autograd registration and AOT dispatch; ClusteringLayer / dec_kl_loss go
    through torch.ops.scc_b200.soft_assign / dec_kl_loss."""
    from torch.library import opcheck
    from spectrogram_cube_clustering_b200 import ops, synth  # noqa: F401  (registers the ops)
    from spectrogram_cube_clustering_b200.networks import ClusteringLayer, dec_kl_loss
    g = load_golden("dec", "c1")
    z = torch.from_numpy(g["z"]).float().cuda()
    mu = torch.from_numpy(g["mu"]).float().cuda()
    p = torch.from_numpy(g["p"]).float().cuda()
    n, K = z.shape[0], mu.shape[0]
    zg, mug = z.clone().requires_grad_(True), mu.clone().requires_grad_(True)
    opcheck(torch.ops.scc_b200.soft_assign.default, (zg, mug, 1.0))
    opcheck(torch.ops.scc_b200.dec_kl_loss.default, (zg, mug, p, 1.0, 1e-3 / n))
    opcheck(torch.ops.scc_b200.soft_assign_backward.default, (z, mu, torch.randn(n, K, device="cuda"), 1.0))
    opcheck(torch.ops.scc_b200.dec_assign.default, (z, mu, 1.0, 5))
    q, _, f = torch.ops.scc_b200.dec_assign(z, mu, 1.0, 5)
    opcheck(torch.ops.scc_b200.dec_target.default, (q, f, 5))
    opcheck(torch.ops.scc_b200.dec_target_kl_grad.default, (z, mu, f, 1.0, 5, 1e-3 / n))
    opcheck(torch.ops.scc_b200.dec_step.default, (z, mu, 1.0, 5, 1e-3 / n))
    # registered ops == direct launches
    r = torch.ops.scc_b200.dec_step(z, mu, 1.0, 5, 1e-3 / n)
    d = ops.dec_step(z, mu, 1.0, 5, 1e-3 / n)
    assert all(torch.equal(a, d[k]) for a, k in zip(r, ("q", "labels", "f", "p", "dz", "stats")))
    # GMM: functional statistics op + mutating finalize op
    gg = load_golden("gmm", "c1")
    zz = torch.from_numpy(gg["z"]).float().cuda()
    w0, m0, c0 = (torch.from_numpy(gg[k]).double().cuda() for k in ("w0", "mu0", "cov0"))
    params, pchol, ctrl = ops.gmm_pack_params(w0, m0, c0)
    Kg = m0.shape[0]
    opcheck(torch.ops.scc_b200.gmm_em_step.default, (zz, params, Kg))
    stats = torch.ops.scc_b200.gmm_em_step(zz, params, Kg)
    means, weights, cov = m0.clone(), w0.clone(), c0.clone()
    torch.ops.scc_b200.gmm_finalize(stats, float(zz.shape[0]), means, weights, cov, pchol, params, ctrl, 1e-6, 0.0)
    assert rel_err(means.cpu().numpy(), gg["it_means"][0]) < TOL
    # the layer: traced through the dispatcher, gradients in the caller's dtype
    layer = ClusteringLayer(K, 9, 1.0, weights=mu.double()).double().cuda()
    zd = z.double().requires_grad_(True)
    with torch.profiler.profile() as prof:
        qd = layer(zd)
        loss = dec_kl_loss(zd, layer.weights, p, 1.0, 1e-3 / n) + 0.0 * qd.sum()
        loss.backward()
    names = {e.name for e in prof.events()}
    assert "scc_b200::soft_assign" in names and "scc_b200::dec_kl_loss" in names
    assert zd.grad.dtype == torch.float64 and layer.weights.grad.dtype == torch.float64
    assert rel_err(zd.grad.cpu().numpy(), g["dz"]) < 1e-5 and rel_err(layer.weights.grad.cpu().numpy(), g["dmu"]) < 1e-5
