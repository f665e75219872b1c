"""GPU parity: the CUDA path (through the C ABI) vs golden vectors produced by the
reference, and vs the float64 oracle on seeded inputs.

Bars (BASELINE.json north_star): hard labels bit-identical except at exact ties;
q, p, loss, gradients, means, covariances within 1e-5 relative (max-normalised,
fp32 kernels vs float64 reference).  Outputs that pass through the reference's
5-decimal rounding are compared to one rounding quantum (1e-5 absolute): an
fp32 q within ~1e-7 of a rounding boundary legitimately lands on the other side.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, DEC_CASES, GMM_CASES

pytestmark = pytest.mark.gpu

TOL = 1e-5
QUANTUM = 1.0e-5


@pytest.fixture(scope="module")
def ops():
    from spectrogram_cube_clustering_b200 import ops as _ops
    assert torch.cuda.is_available()
    return _ops


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype).contiguous()


@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_assign_golden(ops, case):
    g = load_golden("dec", case)
    z, mu = dev(g["z"]), dev(g["mu"])
    q, labels, stats = ops.dec_assign(z, mu, float(g["alpha"]), 0)
    assert rel_err(q.cpu().numpy(), g["q"]) < TOL
    lab = labels.cpu().numpy().astype(np.int64)
    mism = lab != g["labels"]
    if mism.any():                       # only exact-tie rows may differ (fp64 ties that fp32 breaks)
        qs = np.sort(g["q"][mism], axis=1)
        assert np.all(qs[:, -1] - qs[:, -2] < 1e-6)
    assert mism.mean() < 1e-3
    assert rel_err(stats[:-1].cpu().numpy(), g["q"].sum(0)) < TOL
    assert stats[-1].item() == 0
    # rounded mode: np.round(q, 5), f over the rounded q
    qr, _, stats_r = ops.dec_assign(z, mu, float(g["alpha"]), 5)
    dq = np.abs(qr.cpu().numpy() - g["q_round"])
    assert dq.max() <= QUANTUM * 1.01 and (dq > 1e-7).mean() < 0.03
    assert rel_err(stats_r[:-1].cpu().numpy(), g["f"]) < TOL


@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_target_golden(ops, case):
    g = load_golden("dec", case)
    q_round = dev(g["q_round"])
    f = dev(g["f"], torch.float64)
    p5 = ops.dec_target(q_round, f, 5).cpu().numpy()
    d = np.abs(p5 - g["p"])
    assert d.max() <= QUANTUM * 1.01 and (d > 1e-7).mean() < 0.03
    from oracle import dec as odec
    p0 = ops.dec_target(q_round, f, 0).cpu().numpy()
    assert rel_err(p0, odec.target_distribution(g["q_round"], None)) < TOL
    assert rel_err(ops.colsum(q_round).cpu().numpy(), g["f"]) < 1e-6


@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_kl_grad_api_mode_golden(ops, case):
    """p supplied (the reference's own rounded target) -> loss, dz, dmu of autograd."""
    g = load_golden("dec", case)
    n, d = g["z"].shape
    K = g["mu"].shape[0]
    z, mu, p = dev(g["z"]), dev(g["mu"]), dev(g["p"])
    stats, dz = ops.dec_kl_grad(z, mu, float(g["alpha"]), p=p, scale=float(g["gamma"]) / n)
    stats = stats.cpu().numpy()
    assert abs(stats[0] - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    assert abs(stats[1] - g["p"].sum()) <= 1e-6 * n
    assert rel_err(dz.cpu().numpy(), g["dz"]) < TOL
    assert rel_err(stats[2:].reshape(K, d), g["dmu"]) < TOL
    stats2, none = ops.dec_kl_grad(z, mu, float(g["alpha"]), p=p, scale=float(g["gamma"]) / n, want_dz=False)
    assert none is None and np.array_equal(stats2.cpu().numpy(), stats)      # deterministic reduction


@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_kl_grad_fused_mode_oracle(ops, case):
    """p rebuilt in-kernel from the column sums, unrounded -> oracle unrounded chain."""
    from oracle import dec as odec
    g = load_golden("dec", case)
    n, d = g["z"].shape
    K = g["mu"].shape[0]
    alpha, gamma = float(g["alpha"]), float(g["gamma"])
    ref = odec.dec_step(g["z"], g["mu"], alpha, gamma, round_to=None)
    z, mu = dev(g["z"]), dev(g["mu"])
    _, _, st = ops.dec_assign(z, mu, alpha, 0, want_q=False, want_labels=False)
    stats, dz = ops.dec_kl_grad(z, mu, alpha, f=st, round_decimals=0, scale=gamma / n)
    stats = stats.cpu().numpy()
    assert abs(stats[0] - ref["loss"]) <= TOL * abs(ref["loss"])
    assert rel_err(dz.cpu().numpy(), ref["dz"]) < TOL
    assert rel_err(stats[2:].reshape(K, d), ref["dmu"]) < TOL
    # rounded fused mode reproduces the reference's quantised chain to the quantisation noise
    _, _, st5 = ops.dec_assign(z, mu, alpha, 5, want_q=False, want_labels=False)
    stats5, _ = ops.dec_kl_grad(z, mu, alpha, f=st5, round_decimals=5, scale=gamma / n, want_dz=False)
    stats5 = stats5.cpu().numpy()
    assert abs(stats5[0] - float(g["loss"])) <= 2e-4 * abs(float(g["loss"]))
    assert rel_err(stats5[2:].reshape(K, d), g["dmu"]) < 2e-4


@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_target_kl_grad_golden(ops, case):
    """One-pass target + KL gradient: p written by the kernel == target_distribution(np.round(q,5)) of the
    reference to one rounding quantum; loss / dz / dmu == the three-kernel chain fed with that same p."""
    g = load_golden("dec", case)
    n, d = g["z"].shape
    K = g["mu"].shape[0]
    alpha, gamma = float(g["alpha"]), float(g["gamma"])
    z, mu = dev(g["z"]), dev(g["mu"])
    q5, _, st5 = ops.dec_assign(z, mu, alpha, 5)
    stats, p, dz = ops.dec_target_kl_grad(z, mu, st5, alpha, 5, gamma / n)
    # here p comes from the kernel's OWN fp32 q: a q that rounds to the other side of a 5-decimal
    # boundary (one quantum) moves p = q^2/f/sum by up to 2p/q quanta before p's own rounding
    dp = np.abs(p.cpu().numpy() - g["p"])
    assert dp.max() <= 3 * QUANTUM * 1.01 and (dp > 1e-7).mean() < 0.03 and (dp > 1.5 * QUANTUM).mean() < 2e-3
    # q recomputed by the gradient kernel and the row sum of the rebuilt p follow the operation order of
    # dec_assign / dec_target, so the one-pass p is bit-identical to the two stand-alone kernels' output
    p3 = ops.dec_target(q5, st5, 5)
    assert torch.equal(p, p3)
    stats3, dz3 = ops.dec_kl_grad(z, mu, alpha, p=p, scale=gamma / n)  # same target, p read from memory
    assert torch.equal(stats, stats3) and torch.equal(dz, dz3)
    stats = stats.cpu().numpy()
    assert abs(stats[0] - float(g["loss"])) <= 2e-4 * abs(float(g["loss"]))
    assert rel_err(stats[2:].reshape(K, d), g["dmu"]) < 2e-4
    # unrounded: against the float64 oracle chain
    from oracle import dec as odec
    ref = odec.dec_step(g["z"], g["mu"], alpha, gamma, round_to=None)
    _, _, st0 = ops.dec_assign(z, mu, alpha, 0, want_q=False, want_labels=False)
    stats0, p0, dz0 = ops.dec_target_kl_grad(z, mu, st0, alpha, 0, gamma / n)
    assert rel_err(p0.cpu().numpy(), ref["p"]) < 2 * TOL
    assert abs(stats0[0].item() - ref["loss"]) <= TOL * abs(ref["loss"])
    assert rel_err(dz0.cpu().numpy(), ref["dz"]) < TOL
    assert rel_err(stats0[2:].cpu().numpy().reshape(K, d), ref["dmu"]) < TOL
    # no N-sized outputs requested: statistics unchanged
    stats_n, none_p, none_dz = ops.dec_target_kl_grad(z, mu, st0, alpha, 0, gamma / n, want_p=False, want_dz=False)
    assert none_p is None and none_dz is None and torch.equal(stats_n, stats0)


@pytest.mark.parametrize("n,d,K,alpha,rd", [(1, 9, 8, 1.0, 5), (255, 9, 5, 1.0, 0), (4097, 16, 7, 1.0, 5),
                                            (1000, 10, 16, 2.0, 5), (777, 8, 3, 0.5, 0), (300, 4, 2, 1.0, 5),
                                            (70001, 9, 8, 1.0, 5), (5000, 32, 4, 1.0, 0), (2049, 12, 8, 1.0, 5),
                                            (333, 20, 8, 1.0, 5), (640, 24, 4, 1.0, 0)])
def test_dec_step_one_kernel_matches_two_kernel_chain(ops, n, d, K, alpha, rd):
    """scc_dec_step (assign pass + grid barrier + target/gradient pass in one cooperative kernel) against
    scc_dec_assign + scc_dec_target_kl_grad, and through them against the oracle."""
    assert ops.dec_step_supported(d, K)
    rng = np.random.default_rng(7 * n + d + K)
    z = dev(rng.normal(size=(n, d)).astype(np.float32) * 1.5 + 1.0)
    mu = dev(rng.normal(size=(K, d)).astype(np.float32) + 1.0)
    prev = torch.randint(0, K, (n,), device="cuda", dtype=torch.int32)
    scale = 1e-3 / n
    q2, lab2, f2 = ops.dec_assign(z, mu, alpha, rd, labels_prev=prev)
    st2, p2, dz2 = ops.dec_target_kl_grad(z, mu, f2, alpha, rd, scale)
    out = ops.dec_step(z, mu, alpha, rd, scale, labels_prev=prev)
    torch.cuda.synchronize()
    assert torch.equal(out["q"], q2) and torch.equal(out["labels"], lab2)      # same per-point arithmetic
    assert out["f"][K].item() == f2[K].item()                                  # label-change count
    torch.testing.assert_close(out["f"][:K], f2[:K], rtol=1e-6, atol=0)        # different summation order
    dp = (out["p"] - p2).abs()
    if rd:      # f enters p through (float)(1/f): a last-bit difference may move a value across a rounding boundary
        assert dp.max().item() <= QUANTUM * 1.01 and (dp > 1e-7).float().mean().item() < 5e-3
    else:
        assert dp.max().item() <= 2e-6
    if n > 1:
        assert abs(out["stats"][0].item() - st2[0].item()) <= 2e-4 * abs(st2[0].item()) + 1e-12
        assert (out["stats"][2:] - st2[2:]).abs().max() <= 2e-4 * st2[2:].abs().max()
        assert (out["dz"] - dz2).abs().max() <= 2e-4 * dz2.abs().max()
    if not rd and n > 1:        # unrounded chain: straight against the float64 oracle
        from oracle import dec as odec
        ref = odec.dec_step(z.cpu().numpy(), mu.cpu().numpy(), alpha, 1e-3, round_to=None)
        assert rel_err(out["p"].cpu().numpy(), ref["p"]) < 2 * TOL
        assert abs(out["stats"][0].item() - ref["loss"]) <= TOL * abs(ref["loss"])
        assert rel_err(out["dz"].cpu().numpy(), ref["dz"]) < TOL
        assert rel_err(out["stats"][2:].cpu().numpy().reshape(K, d), ref["dmu"]) < TOL
    # repeat: bit-identical (fixed reduction order), and the barrier counter was reset
    out2 = ops.dec_step(z, mu, alpha, rd, scale, labels_prev=prev)
    assert all(torch.equal(out[k], out2[k]) for k in ("q", "labels", "f", "p", "dz", "stats"))
    # nothing N-sized requested
    out3 = ops.dec_step(z, mu, alpha, rd, scale, want_q=False, want_labels=False, want_p=False, want_dz=False)
    assert out3["q"] is None and out3["dz"] is None and torch.equal(out3["stats"], out["stats"])


def test_dec_step_in_cuda_graph_and_unsupported_shape(ops):
    from spectrogram_cube_clustering_b200 import _lib, synth
    z, mu = synth.latent_points(100_000, 9, 8, device="cuda", rank=3)
    eager = ops.dec_step(z, mu, 1.0, 5, 1e-8)
    bufs = ops.dec_step(z, mu, 1.0, 5, 1e-8)          # allocates the outputs the graph will write
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.dec_step(z, mu, 1.0, 5, 1e-8, out_q=bufs["q"], out_labels=bufs["labels"], out_p=bufs["p"],
                     out_dz=bufs["dz"], out_f=bufs["f"], out_stats=bufs["stats"])        # workspace of this stream
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            ops.dec_step(z, mu, 1.0, 5, 1e-8, out_q=bufs["q"], out_labels=bufs["labels"], out_p=bufs["p"],
                         out_dz=bufs["dz"], out_f=bufs["f"], out_stats=bufs["stats"])
    for k in ("p", "dz", "stats"):
        bufs[k].zero_()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(eager[k], bufs[k]) for k in ("q", "labels", "f", "p", "dz", "stats"))
    assert not ops.dec_step_supported(32, 16)
    z32, mu32 = synth.latent_points(1000, 32, 16, device="cuda")
    with pytest.raises(_lib.SccError):
        ops.dec_step(z32, mu32)


def test_dec_step_nonfinite_point_poisons_f_like_the_reference(ops):
    """A NaN (or infinite) latent point makes its whole q row NaN in the reference (networks.py:280-288), hence every
    f_j (models.py:1320), every p and the loss.  The one-kernel step sums f in fixed point: the contribution is flagged
    in the accumulator word (CountedFix) and the sums read as NaN — the barrier still completes, on every later launch too."""
    from spectrogram_cube_clustering_b200 import synth
    z, mu = synth.latent_points(50_000, 9, 8, device="cuda", rank=11)
    clean = ops.dec_step(z, mu, 1.0, 5, 1e-8)
    for bad in (float("nan"), float("inf")):
        zb = z.clone()
        zb[12_345, 3] = bad
        out = ops.dec_step(zb, mu, 1.0, 5, 1e-8)
        torch.cuda.synchronize()
        assert torch.isnan(out["f"][:8]).all() and torch.isnan(out["stats"][0])
        assert torch.isnan(out["q"][12_345]).all()
        good = torch.ones(50_000, dtype=torch.bool, device="cuda"); good[12_345] = False
        assert torch.equal(out["q"][good], clean["q"][good]) and torch.equal(out["labels"][good], clean["labels"][good])
    again = ops.dec_step(z, mu, 1.0, 5, 1e-8)          # the accumulators were reset: the next launch is clean
    torch.cuda.synchronize()
    assert all(torch.equal(again[k], clean[k]) for k in ("q", "labels", "f", "p", "dz", "stats"))


@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_backward_generic_golden(ops, case):
    g = load_golden("dec", case)
    z, mu, G = dev(g["z"]), dev(g["mu"]), dev(g["G"])
    dz, dmu = ops.dec_backward(z, mu, G, float(g["alpha"]))
    assert rel_err(dz.cpu().numpy(), g["dz_generic"]) < TOL
    assert rel_err(dmu.cpu().numpy(), g["dmu_generic"]) < TOL


@pytest.mark.parametrize("n,d,K", [(1, 9, 8), (255, 9, 5), (257, 32, 16), (4097, 16, 7), (1000, 10, 16),
                                   (777, 8, 3), (300, 4, 2), (513, 12, 16), (1025, 20, 8), (640, 24, 12),
                                   (5000, 32, 4)])          # last: large distances in the register-blocked kernel
def test_dec_shapes_vs_oracle(ops, n, d, K):
    from oracle import dec as odec
    rng = np.random.default_rng(n + d + K)
    z = rng.normal(size=(n, d)).astype(np.float32) * 1.5 + 1.0
    mu = rng.normal(size=(K, d)).astype(np.float32) + 1.0
    ref = odec.dec_step(z, mu, 1.0, 1e-3, round_to=None)
    zt, mt = dev(z), dev(mu)
    q, labels, st = ops.dec_assign(zt, mt, 1.0, 0)
    assert rel_err(q.cpu().numpy(), ref["q"]) < TOL
    assert (labels.cpu().numpy() != ref["labels"]).mean() < 2e-3
    assert rel_err(st[:-1].cpu().numpy(), ref["f"]) < TOL
    p = ops.dec_target(q, st, 0)
    assert rel_err(p.cpu().numpy(), ref["p"]) < 2 * TOL
    stats, dz = ops.dec_kl_grad(zt, mt, 1.0, f=st, scale=1e-3 / n)
    stats = stats.cpu().numpy()
    if n == 1:          # p == q exactly: loss and gradients vanish, only fp32 noise is left
        assert abs(stats[0]) < 1e-8 and np.abs(dz.cpu().numpy()).max() < 1e-8
        return
    assert abs(stats[0] - ref["loss"]) <= TOL * abs(ref["loss"])
    assert rel_err(dz.cpu().numpy(), ref["dz"]) < TOL
    assert rel_err(stats[2:].reshape(K, d), ref["dmu"]) < TOL
    # property: translation invariance  sum_i dz_i = - sum_j dmu_j
    np.testing.assert_allclose(dz.double().sum(0).cpu().numpy(), -stats[2:].reshape(K, d).sum(0),
                               atol=1e-6 * np.abs(stats[2:]).max() * K)


def test_dec_empty_and_label_changes(ops):
    z = torch.zeros(0, 9, device="cuda")
    mu = torch.randn(8, 9, device="cuda")
    q, labels, st = ops.dec_assign(z, mu)
    assert q.shape == (0, 8) and torch.all(st == 0)
    z = torch.randn(5000, 9, device="cuda")
    _, lab, _ = ops.dec_assign(z, mu)
    prev = lab.clone()
    prev[:137] = (prev[:137] + 1) % 8
    _, _, st = ops.dec_assign(z, mu, labels_prev=prev)
    assert st[-1].item() == 137


def test_dec_zero_column_gives_nan_like_reference(ops):
    q = torch.rand(64, 8, device="cuda")
    q[:, 3] = 0
    f = ops.colsum(q)
    p = ops.dec_target(q, f, 0)
    assert torch.isnan(p).all(dim=1).all()          # q^2/0 -> nan poisons every row (models.py:1320-1321)


# ------------------------------------------------------------------------------ GMM
class _GmmState:
    def __init__(self, ops, w, mu, cov):
        self.ops = ops
        self.means = dev(mu, torch.float64)
        self.weights = dev(w, torch.float64)
        self.cov = dev(cov, torch.float64)
        self.K, self.d = self.means.shape
        self.params, self.pchol, self.ctrl = ops.gmm_pack_params(self.weights, self.means, self.cov)

    def em(self, z, resp=None, tol=0.0):
        stats = self.ops.gmm_em_step(z, self.K, self.params, resp=resp, ctrl=self.ctrl)
        self.ops.gmm_finalize(stats, z.shape[0], self.means, self.weights, self.cov, self.pchol, self.params,
                              self.ctrl, tol=tol)
        c = self.ctrl.cpu().numpy()
        return dict(lb=c[0], weights=self.weights.cpu().numpy().copy(), means=self.means.cpu().numpy().copy(),
                    cov=self.cov.cpu().numpy().copy(), pchol=self.pchol.cpu().numpy().copy(), ctrl=c)

    def labels(self, z):
        lab = torch.empty(z.shape[0], dtype=torch.int32, device="cuda")
        self.ops.gmm_em_step(z, self.K, self.params, labels=lab, mode=self.ops.GMM_ESTEP_ONLY)
        return lab.cpu().numpy()


def _mat_err(a, b):
    return max(rel_err(a[k], b[k]) for k in range(a.shape[0]))


GMM_GPU_CASES = [c for c in GMM_CASES]


@pytest.mark.parametrize("case", GMM_GPU_CASES)
def test_gmm_step_from_identical_state_golden(ops, case):
    """One fused EM iteration from sklearn's own state of the previous iteration
    (parity is pinned per iteration from identical state)."""
    g = load_golden("gmm", case)
    if not ops.gmm_supported(g["z"].shape[1], g["mu0"].shape[0]):
        pytest.skip("GMM kernels not instantiated for this d")
    z = dev(g["z"])
    n, K = z.shape[0], g["mu0"].shape[0]
    iters = g["it_lower_bound"].shape[0]
    for it in range(iters):
        if it == 0:
            st = _GmmState(ops, g["w0"], g["mu0"], g["cov0"])
        else:
            st = _GmmState(ops, g["it_weights"][it - 1], g["it_means"][it - 1], g["it_covariances"][it - 1])
            assert _mat_err(st.pchol.cpu().numpy(), g["it_pchol"][it - 1]) < 1e-9      # float64 Cholesky path
        resp = torch.empty(n, K, device="cuda") if it == 0 else None
        h = st.em(z, resp=resp)
        if it == 0:
            # responsibilities are exp() of fp32 log-densities of magnitude ~d: allow 5e-5 absolute at d >= 16
            assert np.max(np.abs(resp.cpu().numpy() - np.exp(g["log_resp0"]))) < (TOL if z.shape[1] < 16 else 5 * TOL)
        assert abs(h["lb"] - g["it_lower_bound"][it]) < TOL * abs(g["it_lower_bound"][it])
        assert rel_err(h["weights"], g["it_weights"][it]) < TOL
        assert rel_err(h["means"], g["it_means"][it]) < TOL
        assert _mat_err(h["cov"], g["it_covariances"][it]) < TOL
        # U = chol(Sigma)^-T amplifies by cond(Sigma) (up to 1/reg_covar = 1e6 here): compare the
        # precision it represents through Sigma, i.e. U U^T Sigma_ref = I
        for k in range(K):
            prod = h["pchol"][k] @ h["pchol"][k].T @ g["it_covariances"][it][k]
            assert np.abs(prod - np.eye(prod.shape[0])).max() < 2e-2
    st = _GmmState(ops, g["it_weights"][-1], g["it_means"][-1], g["it_covariances"][-1])
    assert (st.labels(z) != g["labels_after"]).mean() < 2e-3


@pytest.mark.parametrize("case", GMM_GPU_CASES)
def test_gmm_fit_trajectory_golden(ops, case):
    """Whole device-resident fit (tol = 1e-3, sklearn's stop rule) vs GaussianMixture.fit_predict:
    same iteration count and converged flag; parameters to 1e-4 (errors compound over iterations
    through components whose covariance has cond ~ 1/reg_covar), labels to 0.2 %."""
    g = load_golden("gmm", case)
    if not ops.gmm_supported(g["z"].shape[1], g["mu0"].shape[0]):
        pytest.skip("GMM kernels not instantiated for this d")
    z = dev(g["z"])
    st = _GmmState(ops, g["w0"], g["mu0"], g["cov0"])
    h = None
    for it in range(100):
        h = st.em(z, tol=float(g["tol"]))
    c = h["ctrl"]
    assert int(c[2]) == int(g["fit_n_iter"]) and bool(c[3]) == bool(g["fit_converged"]) and c[4] == 0
    assert abs(c[0] - float(g["fit_lower_bound"])) < 1e-4 * abs(float(g["fit_lower_bound"]))
    assert rel_err(h["means"], g["fit_means"]) < 1e-4
    assert rel_err(h["weights"], g["fit_weights"]) < 1e-4
    assert _mat_err(h["cov"], g["fit_covariances"]) < 2e-4
    assert (st.labels(z) != g["fit_labels"]).mean() < 2e-3


def test_gmm_hard_assign_matches_onehot_moments(ops):
    """mode=HARD with Sigma=I, pi=1/K gives the one-hot (nearest-mean) statistics sklearn derives
    from k-means labels (_base.py:119-128 -> _gaussian_mixture.py:282-320)."""
    from oracle import gmm as ogmm
    g = load_golden("gmm", "c1")
    X = g["z"].astype(np.float64)
    mu0 = g["mu0"]
    K, d = mu0.shape
    lab = np.argmin(((X[:, None, :] - mu0[None]) ** 2).sum(2), axis=1)
    resp = np.zeros((X.shape[0], K)); resp[np.arange(X.shape[0]), lab] = 1
    with np.errstate(divide="ignore"):
        w_ref, mu_ref, cov_ref, _, _ = ogmm.m_step(X, np.log(resp))
    st = _GmmState(ops, np.full(K, 1.0 / K), mu0, np.tile(np.eye(d), (K, 1, 1)))
    z = dev(g["z"])
    stats = ops.gmm_em_step(z, K, st.params, mode=ops.GMM_HARD)
    ops.gmm_finalize(stats, z.shape[0], st.means, st.weights, st.cov, st.pchol, st.params, st.ctrl, tol=0.0)
    assert rel_err(st.weights.cpu().numpy(), w_ref) < 1e-6
    assert rel_err(st.means.cpu().numpy(), mu_ref) < TOL
    assert _mat_err(st.cov.cpu().numpy(), cov_ref) < TOL


def test_gmm_not_positive_definite_flag(ops):
    """Degenerate data (all points identical, reg_covar = 0) must raise the not-PD flag
    (sklearn raises ValueError, _gaussian_mixture.py:343-367) and freeze the fit."""
    z = torch.ones(512, 9, device="cuda")
    K, d = 4, 9
    st = _GmmState(ops, np.full(K, 0.25), np.ones((K, d)) + 0.01 * np.arange(K)[:, None], np.tile(np.eye(d), (K, 1, 1)))
    stats = ops.gmm_em_step(z, K, st.params, ctrl=st.ctrl)
    ops.gmm_finalize(stats, 512, st.means, st.weights, st.cov, st.pchol, st.params, st.ctrl, reg_covar=0.0, tol=0.0)
    c = st.ctrl.cpu().numpy()
    assert c[4] > 0 and c[5] == 1


@pytest.mark.parametrize("n,d,K,alpha,rd", [(4097, 32, 16, 1.0, 0), (70001, 32, 16, 1.0, 5), (1000, 16, 16, 2.0, 0),
                                            (640, 24, 12, 1.0, 5), (513, 12, 16, 1.0, 0), (2000, 20, 9, 0.5, 0)])
def test_dec_two_launch_step_with_u_handoff(ops, n, d, K, alpha, rd):
    """Tiled shapes (K*d > 160): the assign pass hands u_ij = 1/(1 + d_ij/alpha) to the gradient pass
    (scc_dec_assign_u -> scc_dec_target_kl_grad_u) instead of the gradient pass recomputing the distances.
    Same results as the recomputing chain (to the rounding of one MUFU.RCP) and as the oracle."""
    from oracle import dec as odec
    rng = np.random.default_rng(11 * n + d + K)
    z = dev(rng.normal(size=(n, d)).astype(np.float32) * 1.2 + 0.5)
    mu = dev(rng.normal(size=(K, d)).astype(np.float32) + 0.5)
    scale = 1e-3 / n
    u = torch.empty(n, K, device="cuda")
    q, lab, f = ops.dec_assign_u(z, mu, u, alpha, rd, want_q=True)
    q0, lab0, f0 = ops.dec_assign(z, mu, alpha, rd)
    assert torch.equal(q, q0) and torch.equal(lab, lab0) and torch.equal(f, f0)
    d2 = ((z[:, None, :].double() - mu[None].double()) ** 2).sum(2)
    torch.testing.assert_close(u.double(), 1.0 / (1.0 + d2 / alpha), rtol=2e-6, atol=0)
    st_u, p_u, dz_u = ops.dec_target_kl_grad_u(z, mu, u, f, alpha, rd, scale, want_p=True, want_dz=True)
    st_r, p_r, dz_r = ops.dec_target_kl_grad(z, mu, f, alpha, rd, scale)
    dp = (p_u - p_r).abs().max().item()
    assert dp <= (1.01e-5 if rd else 2e-6)
    assert abs(st_u[0].item() - st_r[0].item()) <= (2e-4 if rd else 2e-6) * abs(st_r[0].item())
    assert (st_u[2:] - st_r[2:]).abs().max() <= (2e-4 if rd else 1e-5) * st_r[2:].abs().max()
    assert (dz_u - dz_r).abs().max() <= (2e-3 if rd else 1e-5) * dz_r.abs().max()
    if not rd:
        ref = odec.dec_step(z.cpu().numpy(), mu.cpu().numpy(), alpha, 1e-3, round_to=None)
        assert abs(st_u[0].item() - ref["loss"]) <= TOL * abs(ref["loss"])
        assert rel_err(st_u[2:].cpu().numpy().reshape(K, d), ref["dmu"]) < TOL
        assert rel_err(dz_u.cpu().numpy(), ref["dz"]) < TOL
        assert rel_err(p_u.cpu().numpy(), ref["p"]) < 2 * TOL
    st_n, none_p, none_dz = ops.dec_target_kl_grad_u(z, mu, u, f, alpha, rd, scale)
    assert none_p is None and none_dz is None and torch.equal(st_n, st_u)


def test_dec_u_handoff_is_for_tiled_shapes_only(ops):
    from spectrogram_cube_clustering_b200 import _lib
    z = torch.randn(1000, 9, device="cuda"); mu = torch.randn(8, 9, device="cuda")
    u = torch.empty(1000, 8, device="cuda")
    _, _, f = ops.dec_assign_u(z, mu, u)                    # writing u is always possible
    with pytest.raises(_lib.SccError):
        ops.dec_target_kl_grad_u(z, mu, u, f)               # register-blocked shape: use dec_step


# ------------------------------------------------------------------------------ float64 precision path
@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_float64_path_matches_reference_to_roundoff(ops, case):
    """scc_dec_*_f64: the reference's dtype and operation order -> its numbers, including the SIDE on which the
    5-decimal roundings land (np.round(q, 5), np.round(p, 5)): the rounded chain matches exactly, not to a quantum."""
    g = load_golden("dec", case)
    n, d = g["z"].shape
    K = g["mu"].shape[0]
    alpha, gamma = float(g["alpha"]), float(g["gamma"])
    z, mu = dev(g["z"], torch.float64), dev(g["mu"], torch.float64)
    q, labels, st = ops.dec_assign_f64(z, mu, alpha, 0)
    assert np.abs(q.cpu().numpy() - g["q"]).max() < 1e-14
    assert np.array_equal(labels.cpu().numpy(), g["labels"]) or case == "tie"
    q5, _, st5 = ops.dec_assign_f64(z, mu, alpha, 5)
    assert np.array_equal(q5.cpu().numpy(), g["q_round"])                   # same side of every rounding boundary
    assert rel_err(st5[:K].cpu().numpy(), g["f"]) < 1e-13
    p5, f5 = ops.dec_target_f64(q5, None, 5)
    dp = np.abs(p5.cpu().numpy() - g["p"])
    assert dp.max() <= 1.0000001e-5 and (dp > 1e-9).mean() < 1e-4           # a last-bit f difference may flip a rare boundary value
    assert rel_err(f5.cpu().numpy(), g["f"]) < 1e-13
    # API mode: the reference's own p -> its loss / dz / dmu
    stats, dz, _ = ops.dec_grad_f64(z, mu, alpha, p=dev(g["p"], torch.float64), scale=gamma / n)
    s = stats.cpu().numpy()
    assert abs(s[0] - float(g["loss"])) < 1e-11 * abs(float(g["loss"]))
    assert rel_err(dz.cpu().numpy(), g["dz"]) < 1e-11 and rel_err(s[2:].reshape(K, d), g["dmu"]) < 1e-11
    # fused mode, rounded chain end to end (p rebuilt from f in float64)
    stats_f, dz_f, p_f = ops.dec_grad_f64(z, mu, alpha, f=st5[:K].contiguous(), round_decimals=5, scale=gamma / n, want_p=True)
    assert torch.equal(p_f, p5)
    sf = stats_f.cpu().numpy()
    flipped = (dp.max(axis=1) > 1e-9)
    assert abs(sf[0] - float(g["loss"])) < (1e-10 if not flipped.any() else 1e-5) * abs(float(g["loss"]))
    assert rel_err(dz_f.cpu().numpy()[~flipped], g["dz"][~flipped]) < 1e-10
    # generic upstream gradient
    stats_g, dz_g, _ = ops.dec_grad_f64(z, mu, alpha, grad_q=dev(g["G"], torch.float64))
    assert rel_err(dz_g.cpu().numpy(), g["dz_generic"]) < 1e-11
    assert rel_err(stats_g[2:].cpu().numpy().reshape(K, d), g["dmu_generic"]) < 1e-11


def test_dec_float64_layer_and_target_distribution_api():
    """model.double() callers: ClusteringLayer keeps float64 end to end; models.target_distribution(numpy float64)
    computes in float64 (weak point of round 1: it computed in float32)."""
    from spectrogram_cube_clustering_b200.networks import ClusteringLayer
    from spectrogram_cube_clustering_b200.models import target_distribution
    g = load_golden("dec", "c1")
    n = g["z"].shape[0]
    layer = ClusteringLayer(8, 9, 1.0, weights=torch.from_numpy(g["mu"])).double().cuda()
    z = torch.from_numpy(g["z"]).double().cuda().requires_grad_(True)
    q = layer(z)
    assert q.dtype == torch.float64 and np.abs(q.detach().cpu().numpy() - g["q"]).max() < 1e-14
    loss = float(g["gamma"]) * torch.nn.KLDivLoss(reduction="sum")(torch.log(q), torch.from_numpy(g["p"]).cuda()) / n
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-12 * abs(float(g["loss"]))
    assert rel_err(z.grad.cpu().numpy(), g["dz"]) < 1e-10 and rel_err(layer.weights.grad.cpu().numpy(), g["dmu"]) < 1e-10
    p = target_distribution(g["q_round"])
    assert p.dtype == np.float64
    d = np.abs(p - g["p"])
    assert d.max() <= 1.0000001e-5 and (d > 1e-9).mean() < 1e-4
    # float32 input still takes the float32 throughput kernels
    q32 = ClusteringLayer(8, 9, 1.0, weights=torch.from_numpy(g["mu"]).float()).cuda()(torch.from_numpy(g["z"]).float().cuda())
    assert q32.dtype == torch.float32 and rel_err(q32.detach().cpu().numpy(), g["q"]) < 1e-5
