"""CPU: the host-side C++ mirror of the reference API (include/analisi, src/host, cli/, python/) builds,
reads LAMMPS binaries and numpy buffers exactly like the reference, and FAILS LOUDLY without a GPU
(there is no CPU compute path behind Gofrt::calculate or the CLI)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from analisi_b200 import build as b
from analisi_b200 import cabi, synth
from conftest import ROOT, load_golden

NO_GPU = cabi.device_count() == 0


@pytest.fixture(scope="module")
def host():
    cli, ext = b.build_host()
    d = os.path.dirname(ext)
    if d not in sys.path:
        sys.path.insert(0, d)
    import pyanalisi
    return cli, pyanalisi


def run_cli(cli, args, cwd=None):
    return subprocess.run([cli] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=cwd, timeout=600)


def test_module_surface(host):
    """names of the reference's python interface for this path (pyanalisi/src/pyanalisi.cpp:65-82, :375-550)"""
    _, pa = host
    for name in ("Gofrt", "Gofrt_lammps", "Trajectory", "Traj", "BoxFormat", "info", "has_mmap"):
        assert hasattr(pa, name), name
    for meth in ("reset", "getNumberOfExtraTimestepsNeeded", "calculate"):
        assert hasattr(pa.Gofrt, meth) and hasattr(pa.Gofrt_lammps, meth)
    for meth in ("write_lammps_binary", "get_positions_copy", "get_velocities_copy", "get_box_copy", "get_type",
                 "get_nloaded_timesteps", "getNtimesteps", "get_current_timestep", "getWrapPbc", "minImage"):
        assert hasattr(pa.Trajectory, meth) and hasattr(pa.Traj, meth), meth
    for meth in ("setWrapPbc", "setAccessWindowSize", "setAccessStart", "get_lammps_id", "get_lammps_type"):
        assert hasattr(pa.Traj, meth), meth
    assert {"Invalid", "CellVectors", "LammpsOrtho", "LammpsTriclinic"} <= set(pa.BoxFormat.__members__)
    assert pa.has_mmap()


@pytest.mark.parametrize("format2020,nchunk,shuffle", [(False, 1, None), (False, 3, 7), (True, 1, None), (True, 4, 9)])
def test_mmap_trajectory_reads_lammps_binaries(host, tmp_path, format2020, nchunk, shuffle):
    """both header flavours, several chunks per frame, atoms in a different order in every frame
    (scatter by id, reference lib/src/trajectory.cpp:629-659); sliding window with overlap reuse (:542-586)"""
    _, pa = host
    pos, box, types = synth.small_case(5, (4, 3, 3), 1.1, 2, True, 9)
    raw = (types * 3 + 2).astype(np.int32)
    vel = np.random.default_rng(1).normal(size=pos.shape)
    ids = np.arange(len(raw)) * 2 + 5
    path = str(tmp_path / "t.bin")
    synth.write_lammps_binary(path, pos, box, raw, vel=vel, ids=ids, format2020=format2020, nchunk=nchunk,
                              shuffle_seed=shuffle, first_step=100, step_stride=10)
    tr = pa.Traj(path)
    tr.setWrapPbc(False)
    assert tr.getNtimesteps() == 9 and tr.get_natoms() == len(raw) and tr.is_triclinic()
    assert tr.setAccessWindowSize(4) == 1
    for start in (3, 5, 0, 5, 1):   # forward overlap, far jump back, forward, backward overlap
        assert tr.setAccessStart(start) == 1
        assert tr.get_current_timestep() == start and tr.get_nloaded_timesteps() == 4
        assert np.array_equal(tr.get_positions_copy(), pos[start:start + 4])
        assert np.array_equal(tr.get_velocities_copy(), vel[start:start + 4])
        assert np.array_equal(tr.get_box_copy(), synth.lammps_rows_to_internal(box[start:start + 4]))
    assert np.array_equal(tr.get_type_ids(), types)
    assert np.array_equal(tr.get_lammps_type(), raw)
    assert np.array_equal(tr.get_lammps_id(), ids)
    assert tr.get_ntypes() == 2


def test_write_lammps_binary_round_trip(host, tmp_path):
    """BaseTrajectory::dump_lammps_bin_traj (reference lib/src/basetrajectory.cpp:4-49) -> Traj"""
    _, pa = host
    pos, box, types = synth.small_case(6, (3, 3, 3), 1.0, 3, False, 5)
    vel = np.random.default_rng(2).normal(size=pos.shape)
    tn = pa.Trajectory(pos, vel, types.astype(np.int32), box, pa.BoxFormat.LammpsOrtho, False, False)
    path = str(tmp_path / "w.bin")
    tn.write_lammps_binary(path, 1, 4)
    tr = pa.Traj(path)
    tr.setWrapPbc(False)
    tr.setAccessWindowSize(3)
    tr.setAccessStart(0)
    assert np.array_equal(tr.get_positions_copy(), pos[1:4])
    assert np.array_equal(tr.get_velocities_copy(), vel[1:4])
    assert np.array_equal(tr.get_box_copy(), tn.get_box_copy()[1:4])
    assert np.array_equal(tr.get_type_ids(), types)


def test_numpy_trajectory_cell_vectors_rotation(host):
    """general cell matrices: hand-written Householder QR vs the reference's Eigen one, fixture made by the
    compiled reference (tests/golden/make_golden.py: cell_vectors_rotation).  Bit-exact."""
    _, pa = host
    z = load_golden("cell_vectors_rotation.npz")
    tr = pa.Trajectory(z["pos"], z["vel"], z["types"], z["cells"], pa.BoxFormat.CellVectors, False, True)
    assert tr.is_triclinic()
    assert np.array_equal(tr.get_box_copy(), z["box_internal"])
    assert np.array_equal(tr.get_rotation_matrix(), z["rotation"])
    assert np.array_equal(tr.get_positions_copy(), z["pos_nowrap"])
    assert np.array_equal(tr.get_type_ids(), z["type_ids"])
    # diagonal cells stay orthorhombic and are used in place
    cells = np.tile(np.diag([4.0, 5.0, 6.0]), (3, 1, 1))
    pos = np.random.default_rng(3).random((3, 10, 3))
    to = pa.Trajectory(pos, np.zeros_like(pos), np.zeros(10, dtype=np.int32), cells, pa.BoxFormat.CellVectors, False, False)
    assert not to.is_triclinic()
    assert np.array_equal(to.get_box_copy(), np.tile([0, 0, 0, 2.0, 2.5, 3.0], (3, 1)))
    assert np.array_equal(to.get_positions_copy(), pos)


def test_numpy_trajectory_rejects_bad_buffers(host):
    """the reference's validation (lib/src/trajectory_numpy.cpp:17-86) -> RuntimeError"""
    _, pa = host
    pos = np.zeros((2, 4, 3))
    vel = np.zeros((2, 4, 3))
    ty = np.zeros(4, dtype=np.int32)
    box = np.tile([0, 1.0, 0, 1.0, 0, 1.0], (2, 1))
    F = pa.BoxFormat
    with pytest.raises(RuntimeError, match="types array should be int"):
        pa.Trajectory(pos, vel, ty.astype(np.int64), box, F.LammpsOrtho, False, False)
    with pytest.raises(RuntimeError, match="Wrong size of the type array"):
        pa.Trajectory(pos, vel, ty[:3], box, F.LammpsOrtho, False, False)
    with pytest.raises(RuntimeError, match=r"must be \(:, 9\)"):
        pa.Trajectory(pos, vel, ty, box, F.LammpsTriclinic, False, False)
    with pytest.raises(RuntimeError, match="must be 3"):
        pa.Trajectory(pos, vel, ty, box, F.CellVectors, False, False)
    with pytest.raises(RuntimeError, match="Shape of positions and velocities"):
        pa.Trajectory(pos, vel[:1], ty, box, F.LammpsOrtho, False, False)
    with pytest.raises(RuntimeError, match="Unsupported stride"):
        pa.Trajectory(np.zeros((2, 4, 6))[:, :, ::2], vel, ty, box, F.LammpsOrtho, False, False)
    with pytest.raises(RuntimeError, match="should be double"):
        pa.Trajectory(pos.astype(np.float32), vel, ty, box, F.LammpsOrtho, False, False)


def test_gofrt_host_arithmetic(host):
    """shape, strides, leff, nExtraTimesteps, column description -- none of it needs the GPU
    (reference lib/src/gofrt.cpp:37-71)"""
    _, pa = host
    pos = np.random.default_rng(4).random((40, 6, 3))
    ty = np.array([5, 5, 2, 2, 9, 9], dtype=np.int32)
    box = np.tile([0, 1.0, 0, 1.0, 0, 1.0], (40, 1))
    tr = pa.Trajectory(pos, np.zeros_like(pos), ty, box, pa.BoxFormat.LammpsOrtho, False, False)
    g = pa.Gofrt(tr, 0.1, 0.5, 7, 4, 2, 3, 1, False)
    g.reset(10)
    a = np.array(g, copy=False)
    assert a.shape == (4, 12, 7) and a.strides == (12 * 7 * 8, 7 * 8, 8)
    assert g.getNumberOfExtraTimestepsNeeded(3) == cabi.gofrt_nextra(40, 3, 4) == 4
    assert g.getNumberOfExtraTimestepsNeeded(19) == 3
    g.reset(2)
    assert np.array(g, copy=False).shape == (2, 12, 7)
    d = g.get_columns_description().split("\n")
    assert d[1] == "#g(0, 0), different atom index: 13" and d[2] == "#g(0, 0), same atom index: 25"
    assert "#g(1, 2), different atom index: 5" in d
    g0 = pa.Gofrt(tr, 0.1, 0.5, 7, 0, 2, 3, 1, False)   # tmax 0: as many lags as averaged steps
    g0.reset(9)
    assert np.array(g0, copy=False).shape == (9, 12, 7)
    assert g0.getNumberOfExtraTimestepsNeeded(3) == 11
    with pytest.raises(RuntimeError, match="trajectory is too short"):
        g.reset(30)
        g.calculate(20)


@pytest.mark.skipif(not NO_GPU, reason="checks the behaviour on a machine without a GPU")
def test_no_cpu_fallback(host, tmp_path):
    """Without a B200 the product refuses to compute: python raises, the CLI exits with code 1."""
    cli, pa = host
    pos, box, types = synth.small_case(7, (3, 3, 3), 1.1, 1, False, 30)
    tr = pa.Trajectory(pos, np.zeros_like(pos), types.astype(np.int32), box, pa.BoxFormat.LammpsOrtho, False, False)
    g = pa.Gofrt(tr, 0.0, 1.5, 10, 2, 1, 1, 1, False)
    g.reset(5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g.calculate(0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pa.Trajectory(pos, np.zeros_like(pos), types.astype(np.int32), box, pa.BoxFormat.LammpsOrtho, True, False)
    path = str(tmp_path / "c.bin")
    synth.write_lammps_binary(path, pos, box, types)
    r = run_cli(cli, ["-i", path, "-g", "10", "-F", "0.0", "1.5", "-B", "2", "-S", "2"], cwd=str(tmp_path))
    assert r.returncode == 1 and "no CPU path" in r.stderr and r.stdout == ""


def test_cli_option_handling(host, tmp_path):
    """exit codes and messages of the option layer (reference analisi/main.cpp:236-262, :552-556, :719-722)"""
    cli, _ = host
    r = run_cli(cli, ["-h"])
    assert r.returncode == 0 and "--gofrt" in r.stdout
    assert run_cli(cli, []).returncode == 1
    r = run_cli(cli, ["-i", "x.bin", "-V"])
    assert r.returncode == 1 and "this build does not provide" in r.stderr
    r = run_cli(cli, ["-i", "x.bin", "--nonsense"])
    assert r.returncode == 1 and "unrecognised option" in r.stderr
    r = run_cli(cli, ["-i", "/nonexistent/file.bin", "-g", "10", "-F", "0", "1"])
    assert r.returncode == 1 and "Error opening the trajectory" in r.stderr
    pos, box, types = synth.small_case(7, (3, 3, 3), 1.1, 1, False, 4)
    path = str(tmp_path / "c.bin")
    synth.write_lammps_binary(path, pos, box, types)
    r = run_cli(cli, ["-i", path, "-g", "10", "-F", "0.5"])
    assert r.returncode == 1 and "distance range with the option -F" in r.stderr
    r = run_cli(cli, ["-i", path, "-g", "10", "-F", "0", "1", "-s", "0"])
    assert r.returncode == 1 and "Allowed options" in r.stdout
    r = run_cli(cli, ["-i", path])
    assert r.returncode == 1 and "Nothing to do" in r.stderr


def test_python_helpers_host_arithmetic():
    """max_l / hist2gofr of the reference's python wrappers (pyanalisi/analysis.py:77-105)"""
    from analisi_b200 import analysis
    assert analysis.max_l(0, 100, 0) == (50, 50)
    assert analysis.max_l(10, 100, 20) == (20, 70)
    assert analysis.max_l(0, 5, 9) == (4, 1)
    with pytest.raises(RuntimeError):
        analysis.max_l(5, 5, 1)
    h = np.ones((1, 2, 4))
    g = analysis.hist2gofr(4, 0.5, 1.0, h)
    r = 1.0 + 0.5 * np.arange(5)
    assert np.allclose(g[0, 0], 1.0 / (4 * np.pi / 3 * (r[1:] ** 3 - r[:-1] ** 3)))


def test_mmap_trajectory_read_ahead(host, tmp_path, capfd):
    """walking the file in equal steps (what BlockAverageG does): from the third window on the frames come
    from the background read-ahead; contents identical, also after a jump that misses the prediction"""
    _, pa = host
    pos, box, types = synth.small_case(8, (4, 3, 3), 1.1, 2, False, 17)
    path = str(tmp_path / "p.bin")
    synth.write_lammps_binary(path, pos, box, types, nchunk=2, shuffle_seed=4)
    tr = pa.Traj(path)
    tr.setLoadVelocities(False)
    tr.setWrapPbc(False)
    tr.setAccessWindowSize(4)
    capfd.readouterr()
    for start in (0, 3, 6, 9, 12, 2, 5, 8):
        assert tr.setAccessStart(start) == 1
        assert np.array_equal(tr.get_positions_copy(), pos[start:start + 4])
        assert np.array_equal(tr.get_box_copy(), synth.lammps_rows_to_internal(box[start:start + 4]))
    err = capfd.readouterr().err
    assert err.count("(read ahead)") == 4   # 6, 9, 12 and 8; 15 would run past the end and is never started
    assert tr.get_velocities_copy().size == 0


def test_householder_qr_200_cells(host):
    """200 general cells (scales 0.5-20, triangular, partly zero, all-negative, nearly degenerate first column):
    internal box, Q and rotated positions bit-identical to the reference's Eigen-based TriclinicLammpsCell
    (fixture tests/golden/qr_many.npz made by the compiled reference)."""
    _, pa = host
    z = load_golden("qr_many.npz")
    pos = z["pos"]
    tr = pa.Trajectory(pos, np.zeros_like(pos), np.zeros(2, dtype=np.int32), z["cells"], pa.BoxFormat.CellVectors, False, True)
    assert np.array_equal(tr.get_box_copy(), z["box_internal"])
    assert np.array_equal(tr.get_rotation_matrix(), z["rotation"])
    assert np.array_equal(tr.get_positions_copy(), z["pos_rotated"])
