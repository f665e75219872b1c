"""CPU oracle for the latent-space clustering hot path.  TEST INFRASTRUCTURE ONLY.

This package is a float64 CPU restatement of the reference algorithm for the
path named in BASELINE.json (DEC clustering layer + full-covariance GMM EM).
It exists to *check* the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it; nothing under
``spectrogram_cube_clustering_b200/`` does, and the product path raises when
the CUDA library is missing instead of falling back to this code.

Parity pinning
--------------
The reference ships no tests and no golden vectors, so the
oracle is pinned against *outputs of the reference itself run in the build
container*: ``oracle/make_golden.py`` imports ``/root/reference/Cluster`` and
scikit-learn 1.9.0 (the reference's un-vendored, un-pinned GMM dependency,
``/root/reference/setup.py:33``) and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every oracle function against those
fixtures.
"""
