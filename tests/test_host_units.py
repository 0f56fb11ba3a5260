"""CPU: unit checks of the host C++ classes that need no GPU (tests/cpp/host_units.cpp, plain g++): VectorOp,
MediaVar, BlockAverageG block geometry and call order (mock calculation), CalculateMultiThread coercions, box
permutation, the hand-written Householder QR.  MediaVar is compared with the oracle's restatement and with numpy."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle
from analisi_b200 import build as b
from conftest import ROOT


@pytest.fixture(scope="module")
def units(tmp_path_factory):
    b.build_host()
    exe = str(tmp_path_factory.mktemp("cpp") / "host_units")
    srcs = [os.path.join(ROOT, "tests", "cpp", "host_units.cpp")] + [os.path.join(ROOT, "src", "host", s) for s in b.HOST_SOURCES]
    cmd = [b.host_cxx()] + b.HOST_FLAGS + ["-I", os.path.join(ROOT, "include")] + srcs + \
          ["-L", os.path.join(ROOT, "analisi_b200"), "-lagofrt", "-Wl,-rpath," + os.path.join(ROOT, "analisi_b200"), "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout)


def test_vectorop_algebra(units):
    a = np.array([1.0, 2.0, 3.0, 4.0])
    bb = 0.5 * a
    v = (((a + bb) * bb - 0.25) / 2.0) / bb
    assert np.array_equal(units["vectorop"], v)
    assert units["vectorop_size_mismatch_throws"] and units["vectorop_range_throws"] and units["vectorop_copy_len"] == 4


def test_calculate_multithread_coerces_zero_to_one(units):
    """reference lib/include/calculatemultithread.h:41-48"""
    assert units["cmt"] == [1, 1, 1]


def test_box_permutation(units):
    """reference tests/src/test_lammps2020.cpp:50-58 + the internal layout xlo,ylo,zlo,lx/2,ly/2,lz/2"""
    assert units["box_round_trip"]
    assert units["internal_box"] == [1.0, 3.0, 5.0, 0.5, 0.5, 0.5]


def test_triclinic_qr_properties(units):
    assert units["qr_residual"] < 1e-13 and units["q_orthogonality"] < 1e-14 and units["rotate_residual"] < 1e-13
    assert units["qr_signs_ok"] and units["diag_detected"]


def test_block_average_geometry_and_welford(units):
    """reference lib/include/blockaverage.h:107-144: s = (nts - nExtra)/n_b = (47-4)/5 = 8, blocks start at
    iblock*s, in order; MediaVar (calcoliblocchi.h:35-61) = Welford + /((n_b-1) n_b): bit-identical to the oracle's
    restatement, and equal to numpy's mean / variance of the mean to rounding."""
    assert units["ba_block_size"] == 8
    assert units["ba_calls"] == [0, 8, 16, 24, 32]
    blocks = np.array(units["ba_blocks"]).reshape(5, 6)
    mean, var = oracle.mediavar(blocks)
    assert np.array_equal(units["ba_mean"], mean)
    assert np.array_equal(units["ba_var"], var)
    assert np.allclose(mean, blocks.mean(axis=0), rtol=1e-14)
    assert np.allclose(var, blocks.var(axis=0, ddof=1) / 5, rtol=1e-12)
    assert units["ntypes"] == 2 and units["type_ids"] == [0, 0, 1]


# ---- analisi_b200/analysis.py: the python-side helpers that need no extension -----------------------------------------
def test_analysis_lag_split():
    """Analysis.max_l of the reference (pyanalisi/analysis.py:77-89): half of the window by default, the rest are origins;
    a window too short for one origin keeps one."""
    from analisi_b200 import analysis as an
    assert an.lag_split(0, 100) == (50, 50)
    assert an.lag_split(10, 110, 30) == (30, 70)
    assert an.lag_split(0, 10, 10) == (9, 1)
    assert an.lag_split(0, 10, 25) == (9, 1)
    assert an.lag_split(5, 6) == (0, 1)
    assert an.max_l is an.lag_split
    with pytest.raises(RuntimeError):
        an.lag_split(7, 7)


def test_analysis_shell_normalise_and_hist2gofr():
    """g(r) = histogram / volume of the spherical shell of every bin (the only normalisation the reference applies on the
    python side, pyanalisi/analysis.py:91-99)."""
    import math
    from analisi_b200 import analysis as an
    rmin, dr, nbin = 0.5, 0.25, 6
    h = np.arange(2 * 3 * nbin, dtype=np.float64).reshape(2, 3, nbin) + 1.0
    g = an.shell_normalise(h, rmin, dr)
    for k in range(nbin):
        vol = 4.0 * math.pi / 3.0 * ((rmin + (k + 1) * dr) ** 3 - (rmin + k * dr) ** 3)
        assert np.allclose(g[..., k], h[..., k] / vol, rtol=1e-14, atol=0)
    assert np.array_equal(an.hist2gofr(nbin, dr, rmin, h), g)
    with pytest.raises(ValueError):
        an.hist2gofr(nbin + 1, dr, rmin, h)
