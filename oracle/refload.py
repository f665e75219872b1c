"""TEST INFRASTRUCTURE — import the UNMODIFIED reference from /root/reference.

Only used in the build container (by oracle/make_golden.py and by the optional
``tests/test_reference_live.py``); /root/reference does not exist on the GPU
box.  The reference's plotting / IO imports (``cmocean, h5py, matplotlib,
obspy, zarr, dask``) are absent from this image and are never touched by the
clustering path, so they are stubbed in ``sys.modules``.
"""
from __future__ import annotations

import os
import sys
import warnings
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("SCC_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "cmocean", "cmocean.cm", "h5py", "matplotlib", "matplotlib.gridspec", "matplotlib.patches",
    "matplotlib.pyplot", "matplotlib.colors", "matplotlib.ticker", "matplotlib.dates",
    "mpl_toolkits", "mpl_toolkits.axes_grid1", "mpl_toolkits.axes_grid1.inset_locator",
    "obspy", "zarr", "dask", "dask.array",
]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Cluster", "networks.py"))


def load():
    """Returns (Cluster.networks, Cluster.models) of the reference."""
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    for name in _STUBS:
        sys.modules.setdefault(name, MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import Cluster.networks as networks
        import Cluster.models as models
    return networks, models
