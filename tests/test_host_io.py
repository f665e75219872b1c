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
