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

# systems of a few thousand atoms through the compiled reference (tests/golden/make_golden.py: TILE_CASES)
TILE_CASES = {
    "tri_tile3": (41, (16, 16, 12), 3, True, 5),
    "ortho_tile_dense": (42, (12, 12, 12), 2, False, 5),
    "ortho_tile_rmin": (43, (12, 12, 12), 1, False, 5),
}


def tile_case(name):
    """Inputs regenerated with synth.small_case (their sha256 is part of the fixture), results from the fixture."""
    import hashlib
    from analisi_b200 import synth
    z = load_golden("live_reference_tile.npz")
    pre = name + "/"
    d = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    seed, cells, ntypes, tri, nframes = TILE_CASES[name]
    pos, box, types = synth.small_case(seed, cells, 1.1, ntypes, tri, nframes, "parity")
    sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pos).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, d["pos_in_sha256"]), "synth.small_case no longer reproduces the fixture's input: regenerate it"
    d["pos_in"], d["box_lammps"], d["types"] = pos, box, types
    return d


@pytest.fixture(scope="session")
def ctx():
    from analisi_b200 import cabi
    c = cabi.Context()
    yield c
    c.close()


# Neighbours / SphericalBase fixtures from the compiled reference (tests/golden/make_golden.py: PAIR_LOOP_CASES)
PAIR_LOOP_CASES = {
    "tri2": (81, (4, 4, 3), 2, True, 1, [(40, 4.0, 4.0), (40, 6.25, 6.25)], 4, 3, [(0.5, 3.0), (0.4, 2.6), (0.4, 2.6), (0.0, 2.0)]),
    "ortho3": (82, (5, 4, 3), 3, False, 2, [(30, 3.0, 3.0), (30, 3.0, 3.0), (30, 5.0, 5.0)], 6, 2, [(0.3, 2.4)] * 9),
    "tri1_l10": (83, (3, 3, 3), 1, True, 0, [(60, 9.0, 9.0)], 10, 2, [(0.0, 3.2)]),
}


def pair_loop_case(name):
    import hashlib
    from analisi_b200 import synth
    z = load_golden("pair_loops.npz")
    pre = name + "/"
    d = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
    seed, cells, ntypes, tri, frame, spec, lmax, nbin, rminmax = PAIR_LOOP_CASES[name]
    pos, box, types = synth.small_case(seed, cells, 1.1, ntypes, tri, 3, "parity")
    sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pos).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, d["pos_in_sha256"]), "synth.small_case no longer reproduces the fixture's input: regenerate it"
    d.update(pos=pos, box=box, types=types, ntypes=ntypes, tri=tri, frame=frame, spec=spec, lmax=lmax, nbin=nbin, rminmax=rminmax)
    return d
