"""Pin the CPU oracle (oracle/) against fixtures produced by the reference itself
(tests/golden/*.npz, written by oracle/make_golden.py from /root/reference +
scikit-learn).  float64 vs float64, so the bar is ~1e-12."""
import numpy as np

from conftest import rel_err
from oracle import dec as odec
from oracle import gmm as ogmm

TIGHT = 1e-11


def test_dec_forward_matches_reference(dec_golden):
    _, g = dec_golden
    q = odec.soft_assign(g["z"], g["mu"], float(g["alpha"]))
    assert rel_err(q, g["q"]) < TIGHT
    np.testing.assert_allclose(q.sum(1), 1.0, atol=1e-13)
    assert np.array_equal(odec.labels_from_q(q), g["labels"])
    assert np.array_equal(odec.round_decimals(q), g["q_round"])


def test_dec_target_distribution_matches_reference(dec_golden):
    _, g = dec_golden
    p = odec.target_distribution(g["q_round"])
    assert np.array_equal(p, g["p"])
    assert rel_err(odec.column_sums(g["q_round"]), g["f"]) < 1e-14


def test_dec_loss_and_closed_form_grads_match_autograd(dec_golden):
    _, g = dec_golden
    n = g["z"].shape[0]
    alpha, gamma = float(g["alpha"]), float(g["gamma"])
    loss, dz, dmu = odec.kl_grads(g["z"], g["mu"], g["p"], alpha, gamma / n)
    assert abs(loss - float(g["loss"])) <= 1e-12 * abs(float(g["loss"]))
    assert rel_err(dz, g["dz"]) < TIGHT
    assert rel_err(dmu, g["dmu"]) < TIGHT
    # translation invariance: sum_i dL/dz_i = - sum_j dL/dmu_j
    np.testing.assert_allclose(dz.sum(0), -dmu.sum(0), atol=1e-15)


def test_dec_generic_backward_matches_autograd(dec_golden):
    _, g = dec_golden
    dz, dmu = odec.backward_generic(g["z"], g["mu"], g["G"], float(g["alpha"]))
    assert rel_err(dz, g["dz_generic"]) < TIGHT
    assert rel_err(dmu, g["dmu_generic"]) < TIGHT


def test_dec_step_composition(dec_golden):
    _, g = dec_golden
    out = odec.dec_step(g["z"], g["mu"], float(g["alpha"]), float(g["gamma"]))
    assert np.array_equal(out["labels"], g["labels"])
    assert np.array_equal(out["p"], g["p"])
    assert rel_err(out["dmu"], g["dmu"]) < TIGHT


def test_dec_tie_first_index_wins():
    from conftest import load_golden
    g = load_golden("dec", "tie")
    # centroids 0 and 1 are identical: label 1 can never be chosen
    assert not np.any(g["labels"] == 1)
    assert np.array_equal(odec.labels_from_q(odec.soft_assign(g["z"], g["mu"])), g["labels"])


def test_dec_torch_mirror_matches_golden(dec_golden):
    import torch
    _, g = dec_golden
    loss, dz, dmu, labels, p = odec.torch_dec_step(
        torch.from_numpy(g["z"]).double(), torch.from_numpy(g["mu"]).double(),
        float(g["alpha"]), float(g["gamma"]))
    assert abs(loss - float(g["loss"])) <= 1e-12 * abs(float(g["loss"]))
    assert rel_err(dz.numpy(), g["dz"]) < TIGHT
    assert rel_err(dmu.numpy(), g["dmu"]) < TIGHT
    assert np.array_equal(labels, g["labels"]) and np.array_equal(p, g["p"])


def test_gmm_steps_match_sklearn(gmm_golden):
    _, g = gmm_golden
    X = g["z"].astype(np.float64)
    w, mu, cov = g["w0"], g["mu0"], g["cov0"]
    pchol = ogmm.precision_cholesky(cov)
    for it in range(g["it_lower_bound"].shape[0]):
        lb, log_resp = ogmm.e_step(X, w, mu, pchol)
        if it == 0:
            assert np.max(np.abs(np.exp(log_resp) - np.exp(g["log_resp0"]))) < 1e-12
        w, mu, cov, pchol, _ = ogmm.m_step(X, log_resp)
        assert abs(lb - g["it_lower_bound"][it]) < 1e-11
        assert rel_err(w, g["it_weights"][it]) < TIGHT
        assert rel_err(mu, g["it_means"][it]) < TIGHT
        assert rel_err(cov, g["it_covariances"][it]) < 1e-10
        assert rel_err(pchol, g["it_pchol"][it]) < 1e-9
    _, log_resp = ogmm.e_step(X, w, mu, pchol)
    assert np.array_equal(np.argmax(log_resp, 1), g["labels_after"])


def test_gmm_fit_matches_sklearn_fit_predict(gmm_golden):
    _, g = gmm_golden
    out = ogmm.fit(g["z"].astype(np.float64), g["w0"], g["mu0"], g["cov0"], max_iter=100, tol=float(g["tol"]))
    assert out["n_iter"] == int(g["fit_n_iter"])
    assert out["converged"] == bool(g["fit_converged"])
    assert abs(out["lower_bounds"][-1] - float(g["fit_lower_bound"])) < 1e-10
    assert rel_err(out["means"], g["fit_means"]) < 1e-10
    assert rel_err(out["covariances"], g["fit_covariances"]) < 1e-9
    assert rel_err(out["weights"], g["fit_weights"]) < 1e-10
    assert np.array_equal(out["labels"], g["fit_labels"])


def test_gmm_not_pd_raises():
    import pytest
    cov = np.zeros((1, 3, 3))
    with pytest.raises(ValueError):
        ogmm.precision_cholesky(cov)
