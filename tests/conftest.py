"""pytest configuration.

Markers: ``gpu`` = needs a real B200 (run by the driver with ``-m gpu`` on the GPU box).  Everything
else runs on CPU in a few minutes.  The oracle (``oracle/``) is the checker in both; the product
path is ``analisi_b200`` -> ``libagofrt.so`` (C ABI) and has no CPU fallback.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("ANALISI_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) and the product library (nvcc cross-compiles without a GPU)."""
    import oracle
    oracle.build()
    from analisi_b200 import build as b
    b.build()


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def live_case(name):
    z = load_golden("live_reference.npz")
    pre = name + "/"
    return {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}


LIVE_CASES = ["tri_wrap", "tri_npt", "ortho_unwrapped", "tri_bigtilt", "ragged"]


@pytest.fixture(scope="session")
def ctx():
    from analisi_b200 import cabi
    c = cabi.Context()
    yield c
    c.close()
