"""What bar can ANY float32 implementation of the reference's ROUNDED chain meet?  (CPU, oracle only.)

The reference quantises q and p to 5 decimals (`np.round(q, 5)`, models.py:94; `np.round(p, 5)`, models.py:1322),
so its outputs are a discontinuous function of its inputs.  This test perturbs the INPUT of the float64 oracle by
one float32 ulp (relative 2^-24: the least error a float32 latent set carries) and records how far the float64
reference itself moves:

  * rounded chain: some q entries land on the other side of a rounding boundary (one quantum, 1e-5), p moves by up
    to 2-3 quanta for those rows, and dL/dz of those rows moves by MORE than north_star's 1e-5 relative — while the
    sums over points (f, loss, dL/dmu) stay far inside 1e-5;
  * unrounded chain: everything stays inside 1e-7.

Hence the GPU parity bars (tests/test_gpu_baseline_sizes.py, tests/test_gpu_parity.py): 1e-5 for every output of
the unrounded chain and for the sums of the rounded chain; one quantum for q, three for p, and "exact except on
flipped rows" for dz in the rounded chain.  Callers that need the rounded chain bit for bit use the float64 entry
points (the reference's own `model.double()` precision).
"""
import numpy as np
import pytest

from oracle import dec as odec
from spectrogram_cube_clustering_b200 import synth

QUANTUM = 1e-5


@pytest.mark.parametrize("n", [10_000, 100_000])
def test_float64_reference_under_one_float32_ulp_of_input_noise(n):
    z, mu = synth.latent_points(n, 9, 8)
    z = z.numpy().astype(np.float64)
    mu = mu.numpy().astype(np.float64)
    rng = np.random.default_rng(1)
    zp = z * (1.0 + 2.0 ** -24 * rng.choice([-1.0, 1.0], size=z.shape))

    def rel(a, b):
        return float(np.abs(a - b).max() / np.abs(b).max())

    ref = odec.dec_step_chunked(z, mu, 1.0, 1e-3, 5)
    per = odec.dec_step_chunked(zp, mu, 1.0, 1e-3, 5)
    dq = np.abs(per["q_rounded"] - ref["q_rounded"])
    dp = np.abs(per["p"] - ref["p"])
    flips = float((dq > 1e-7).mean())
    # the float64 reference itself flips entries, by exactly one quantum in q and up to three in p ...
    assert 0.0 < flips < 0.01
    assert abs(dq.max() - QUANTUM) < 1e-9
    assert 2 * QUANTUM * 0.99 <= dp.max() <= 3 * QUANTUM * 1.01
    # ... and dz of those rows moves by more than 1e-5 relative: no float32 kernel can promise 1e-5 there
    ddz = np.abs(per["dz"] - ref["dz"]).max(axis=1) / np.abs(ref["dz"]).max()
    flipped = dq.max(axis=1) > 1e-7
    assert ddz.max() > 1e-5 and ddz[~flipped & (dp.max(axis=1) < 1e-7)].max() < 1e-6
    # the sums over points average the flips out
    assert abs(per["loss"] - ref["loss"]) < 1e-6 * abs(ref["loss"])
    assert rel(per["dmu"], ref["dmu"]) < 2e-6 and rel(per["f"], ref["f"]) < 1e-7
    # unrounded chain: continuous, moves by ~ the input noise
    r0 = odec.dec_step_chunked(z, mu, 1.0, 1e-3, None)
    p0 = odec.dec_step_chunked(zp, mu, 1.0, 1e-3, None)
    for k in ("q", "p", "dz", "dmu"):
        assert rel(p0[k], r0[k]) < 1e-6, k
    assert abs(p0["loss"] - r0["loss"]) < 1e-7 * abs(r0["loss"])
