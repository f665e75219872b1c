"""Host-side glue that needs no GPU: the on-disk side effects of gmm_fit (reference models.py:415-449,
utils.py:1181-1209) with the clustering itself stubbed out."""
import csv
import types

import numpy as np


def test_gmm_fit_writes_reference_files(tmp_path, monkeypatch):
    from spectrogram_cube_clustering_b200 import models
    labels = np.array([2, 0, 1, 1, 0], dtype=np.int64)
    centroids = np.arange(27, dtype=np.float64).reshape(3, 9)
    monkeypatch.setattr(models, "gmm", lambda z, k: (labels, centroids))          # no CUDA launch in this test
    cfg = types.SimpleNamespace(savepath_run=str(tmp_path))
    out_l, out_c = models.gmm_fit(cfg, np.zeros((5, 9), dtype=np.float32), 3)
    assert out_l is labels and out_c is centroids
    assert np.array_equal(np.load(tmp_path / "labels.npy"), labels)
    assert np.array_equal(np.load(tmp_path / "centroids.npy"), centroids)
    rows = list(csv.DictReader(open(tmp_path / "Labels.csv")))
    assert [r["idx"] for r in rows] == ["0", "1", "2", "3", "4"] and [int(r["label"]) for r in rows] == labels.tolist()
    # a second call appends rows without a second header (utils.save_labels behaviour)
    models.gmm_fit(cfg, np.zeros((5, 9), dtype=np.float32), 3)
    lines = open(tmp_path / "Labels.csv").read().strip().splitlines()
    assert lines[0] == "idx,label" and len(lines) == 11 and lines.count("idx,label") == 1
    assert models.save_labels([{"idx": 0, "label": 1}], str(tmp_path), serial="_x").endswith("Labels_x.csv")


def test_initialize_clusters_load_and_rand_need_no_device(tmp_path):
    """models.py:521-533: 'load' reads <dir of saved_weights>/GMM/n_clusters=K/{labels,centroids}.npy (labels subset by
    config.index_tra); 'rand' draws without touching the model."""
    from spectrogram_cube_clustering_b200 import models
    K = 3
    wdir = tmp_path / "AEC"
    gdir = wdir / "GMM" / f"n_clusters={K}"
    gdir.mkdir(parents=True)
    labels = np.arange(10) % K
    centroids = np.arange(K * 9, dtype=np.float64).reshape(K, 9)
    np.save(gdir / "labels.npy", labels)
    np.save(gdir / "centroids.npy", centroids)
    cfg = types.SimpleNamespace(init="load", saved_weights=str(wdir / "AEC_Params_Final.pt"), index_tra=np.array([1, 3, 5]))
    lab, cent = models.initialize_clusters(types.SimpleNamespace(n_clusters=K), None, cfg, n_clusters=K)
    assert np.array_equal(lab, labels[[1, 3, 5]]) and np.array_equal(cent, centroids)
    cfg_all = types.SimpleNamespace(init="load", saved_weights=str(wdir / "AEC_Params_Final.pt"))
    assert np.array_equal(models.initialize_clusters(None, None, cfg_all, n_clusters=K)[0], labels)
    import torch
    model = types.SimpleNamespace(n_clusters=K, clustering=types.SimpleNamespace(weights=torch.zeros(K, 9)))
    loader = types.SimpleNamespace(dataset=list(range(25)))
    lab_r, cent_r = models.initialize_clusters(model, loader, types.SimpleNamespace(init="rand"), n_clusters=K)
    assert lab_r.shape == (25,) and cent_r.shape == (K, 9) and lab_r.max() < K


def test_dec_params_and_history_files(tmp_path):
    """models.py:1009-1012, 1227-1228 (state dict with the reference's keys) and utils.py:1158-1178 (history CSV)."""
    import torch
    from spectrogram_cube_clustering_b200 import models
    from spectrogram_cube_clustering_b200.networks import DEC
    model = DEC(n_clusters=5)
    fname = models.save_dec_params(model, str(tmp_path), "Initial")
    assert fname.endswith("DEC_Params_Initial.pt")
    sd = torch.load(fname)
    assert "clustering.weights" in sd and "encoder.encoder.8.weight" in sd and "decoder.decoder.0.weight" in sd
    DEC(n_clusters=5).load_state_dict(sd, strict=True)
    path = models.save_history({"Iteration": [1, 2], "Delta": [0.5, 0.25]}, str(tmp_path / "Delta_history.csv"))
    assert open(path).read().splitlines() == ["Iteration,Delta", "1,0.5", "2,0.25"]
