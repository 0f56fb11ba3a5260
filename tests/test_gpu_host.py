"""GPU: the reference-facing layers -- the pybind11 module (python/pyanalisi.cpp) and the CLI (cli/main.cpp),
both host C++ over the C ABI -- against the reference's golden outputs, the fixtures made by the
compiled reference, and the oracle.  The tests read like the reference's own (tests/test_gofrt.py,
tests/test_notebook.py, tests/test_cli.sh)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from analisi_b200 import build as b
from analisi_b200 import cabi, synth
from conftest import GOLDEN, LIVE_CASES, PAIR_LOOP_CASES, ROOT, live_case, load_golden, pair_loop_case

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12   # north_star: normalised g(r,t) within 1e-12 relative
REFDATA = os.path.join(ROOT, "tests", "_refdata")   # copies of the reference's test trajectories (git-ignored)


@pytest.fixture(scope="module")
def host():
    cli, ext = b.build_host()
    d = os.path.dirname(ext)
    if d not in sys.path:
        sys.path.insert(0, d)
    import pyanalisi
    return cli, pyanalisi


def close(v, ref):
    scale = np.maximum(np.abs(ref), 1e-300)
    return (np.abs(v - ref) <= REL_TOL * scale).all()


def test_pyanalisi_gofrt_numpy_golden(host, tmp_path, monkeypatch):
    """reference tests/test_gofrt.py verbatim: Gofrt(traj,0.0,3.8,200,10,4,10,False,1); reset(700); calculate(0)
    -- including the positional-argument trap (every=False -> 1, debug=1 -> a gofrt.dump appears)."""
    _, pa = host
    monkeypatch.chdir(tmp_path)
    z = load_golden("gofr_numpy.npz")
    # the fixture keeps the 700 frames the loops touch (origins 0..690, lags 0..9) of the 7958 the reference
    # test loads; the length check of calculate() (gofrt.cpp:81) wants leff + ntimesteps <= frames + 1,
    # so pad with frames that are never read
    pos = np.concatenate([z["pos"], np.repeat(z["pos"][-1:], 10, axis=0)])
    boxl = oracle.internal_to_lammps(np.concatenate([z["box_internal"], np.repeat(z["box_internal"][-1:], 10, axis=0)]))
    tr = pa.Trajectory(pos, np.zeros_like(pos), z["types"].astype(np.int32), boxl, pa.BoxFormat.LammpsOrtho, False, False)
    g = pa.Gofrt(tr, 0.0, 3.8, 200, 10, 4, 10, False, 1)
    g.reset(700)
    g.calculate(0)
    v = np.array(g, copy=True)
    assert v.shape == (10, 12, 200)
    assert np.array_equal(g.counts(), z["counts"])
    nz = z["csv"] != 0
    assert np.abs(v[nz] - z["csv"][nz]).max() <= 1e-11   # the CSV holds the reference's 4-thread float sums
    assert (v[~nz] == 0).all()
    assert os.path.exists(tmp_path / "gofrt.dump")
    rows = open(tmp_path / "gofrt.dump").read().split("\n")
    assert len(rows) == 10 * 200 + 3 and len(rows[0].split()) == 14


def test_pyanalisi_gofrt_lammps_notebook_golden(host, tmp_path):
    """reference tests/test_notebook.py / notebooks/calc_inspector.ipynb: Traj + setWrapPbc(True) +
    Gofrt_lammps(traj,0.5,3.8,100,1,4,10,False,1).  The fixture keeps every 10th frame, so skip is 1 here."""
    _, pa = host
    z = load_golden("gofr_notebook.npz")
    boxl = oracle.internal_to_lammps(z["box_internal"])
    path = str(tmp_path / "nb.bin")
    synth.write_lammps_binary(path, z["pos_unwrapped"], boxl, z["types"])
    tr = pa.Traj(path)
    tr.setWrapPbc(True)
    tr.setAccessWindowSize(100)
    tr.setAccessStart(0)
    assert np.array_equal(tr.get_positions_copy(), z["pos_wrapped"])
    g = pa.Gofrt_lammps(tr, 0.5, 3.8, 100, 1, 4, 1, 1, False)
    g.reset(100)   # the 100 kept frames = the 100 origins 0,10,..,990 of reset(999) with skip 10
    g.calculate(0)
    assert np.array_equal(g.counts(), z["counts"])
    # the reference normalises by int(999/10) = 99 (gofrt.cpp:91), this run by 100
    assert np.abs(np.array(g) * (100.0 / 99.0) - z["csv"]).max() < 1e-11


@pytest.mark.parametrize("name", LIVE_CASES)
def test_pyanalisi_live_reference_cases(host, name):
    """triclinic / NPT / unwrapped / big tilt / ragged loops through pyanalisi.Trajectory + Gofrt: counts
    bit-exact, count*incr within 1e-12 relative of the reference's own vdata (fixtures by the compiled reference)."""
    _, pa = host
    d = live_case(name)
    rmin, rmax, nbin, tmax, skip, every, nts, primo = d["params"]
    fmt = pa.BoxFormat.LammpsTriclinic if d["box_lammps"].shape[1] == 9 else pa.BoxFormat.LammpsOrtho
    pos = np.ascontiguousarray(d["pos_in"])
    tr = pa.Trajectory(pos, np.zeros_like(pos), d["raw_types"].astype(np.int32), d["box_lammps"], fmt, bool(d["wrap"]), False)
    assert np.array_equal(tr.get_positions_copy(), d["pos_ref"])
    assert np.array_equal(tr.get_box_copy(), d["box_internal"])
    assert np.array_equal(tr.get_type_ids(), d["type_ids"])
    g = pa.Gofrt(tr, rmin, rmax, int(nbin), int(tmax), 3, int(skip), int(every), False)
    g.reset(int(nts))
    g.calculate(int(primo))
    assert np.array_equal(g.counts(), d["counts"])
    assert close(np.array(g), d["vdata"])


def test_pyanalisi_cell_vectors_rotation(host):
    """CellVectors input with rotated cells: QR on the host, wrap + g(r,t) on the GPU, vs the compiled reference"""
    _, pa = host
    z = load_golden("cell_vectors_rotation.npz")
    tr = pa.Trajectory(z["pos"], z["vel"], z["types"], z["cells"], pa.BoxFormat.CellVectors, True, True)
    assert np.array_equal(tr.get_positions_copy(), z["pos_wrap"])
    # the rotation matrices live on the GPUs next to positions and cells
    rot = tr.get_rotation_matrix()
    for f in (0, 3, 6):
        assert np.array_equal(tr.get_device_rotation_matrix(f), rot[f])
    rmin, rmax, nbin, tmax, skip, every, nts, primo = z["params"]
    g = pa.Gofrt(tr, rmin, rmax, int(nbin), int(tmax), 2, int(skip), int(every), False)
    g.reset(int(nts))
    g.calculate(int(primo))
    assert np.array_equal(g.counts(), z["counts"])
    assert close(np.array(g), z["vdata"])
    # the minImage probe of the trajectory classes (reference pyanalisi.cpp:356-371)
    bi = z["box_internal"]
    d = oracle.d2_all(z["pos_wrap"][2], z["pos_wrap"][4], bi[2])
    for i, j in ((0, 0), (3, 17), (59, 1)):
        assert np.array_equal(tr.minImage(i, j, 2, 4), d[i, j])


def _rows(text):
    return [l for l in text.split("\n") if l and not l.startswith("#")]


def _g(x):
    return "%g" % x


def _cli_text(mean, var, leff, every, nbin, ncol):
    lines = []
    for t in range(0, leff, every):
        for r in range(nbin):
            row = "%d %d" % (t, r)
            for k in range(ncol):
                row += " %s %s" % (_g(mean[t, k, r]), _g(var[t, k, r]))
            lines.append(row)
    return lines


@pytest.mark.parametrize("tri,ntypes,args", [
    (False, 2, dict(g=40, F=(0.0, 2.5), S=5, s=3, e=1, B=4)),
    (True, 3, dict(g=25, F=(0.4, 2.2), S=6, s=2, e=2, B=3)),
    (True, 1, dict(g=30, F=(0.0, 2.0), S=0, s=1, e=1, B=5)),
])
def test_cli_blocks_vs_oracle(host, tmp_path, tri, ntypes, args):
    """analisi -i f -g nbin -F rmin rmax -S lmax -s skip -e every -B blocks: mmap reader + GPU wrap + Gofrt +
    BlockAverage + MediaVar + printing, against the oracle's restatement of the same chain (the oracle is
    pinned by the reference's CLI goldens in test_oracle_golden.py)."""
    cli, _ = host
    nfr = 46
    pos, box, types = synth.small_case(21, (4, 4, 3), 1.05, ntypes, tri, nfr, "parity", (0.2, -0.1, 0.15), False)
    path = str(tmp_path / "c.bin")
    synth.write_lammps_binary(path, pos, box, types * 2 + 1, nchunk=2, shuffle_seed=3)
    argv = ["-i", path, "-g", str(args["g"]), "-F", str(args["F"][0]), str(args["F"][1]), "-S", str(args["S"]),
            "-s", str(args["s"]), "-e", str(args["e"]), "-B", str(args["B"]), "-N", "3"]
    r = subprocess.run([cli] + argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    # the same chain with the oracle
    bi = synth.lammps_rows_to_internal(box)
    wrapped = oracle.pbc_wrap(pos, bi)
    n_b, lmax = args["B"], args["S"]
    nextra = oracle.nextra(nfr, n_b, lmax)
    s = (nfr - nextra) // n_b
    blocks = []
    for ib in range(n_b):
        lo = ib * s
        blocks.append(oracle.vdata(wrapped[lo:lo + s + nextra], bi[lo:lo + s + nextra], types, args["F"][0], args["F"][1],
                                   args["g"], lmax, s, primo=lo, skip=args["s"], every=args["e"], ntypes=ntypes,
                                   first_frame=lo, total_frames=nfr, ref_nthreads=1))
    mean, var = oracle.mediavar(np.array(blocks))
    leff = oracle.leff(s, lmax)
    want = _cli_text(mean, var, leff, args["e"], args["g"], ntypes * (ntypes + 1))
    got = _rows(r.stdout)
    assert len(got) == len(want)
    assert got == want
    head = [l for l in r.stdout.split("\n") if l.startswith("#")]
    assert head[0].startswith("# The first column is the time difference") and len(head) == 3 + ntypes * (ntypes + 1)


@pytest.mark.parametrize("device_blocks", ["1", "0"], ids=["MediaVarDevice", "MediaVar"])
@pytest.mark.parametrize("name,S", [("pair_corr_no_t", 1), ("pair_corr_t", 10)])
def test_cli_reference_golden_text(host, tmp_path, name, S, device_blocks):
    """reference tests/test_cli.sh:33-34 verbatim: analisi -i lammps2020.bin -g 100 -F 0.0 4.0 -S {1,10} -s 8,
    stdout compared with the reference's golden text (tests/golden/cli_*.txt) the way the script does.
    Needs the 51 MB input, copied to tests/_refdata by build() where the reference tree exists."""
    cli, _ = host
    path = os.path.join(REFDATA, "lammps2020.bin")
    if not os.path.exists(path):
        pytest.skip("tests/_refdata/lammps2020.bin not present")
    # the block averages run on the device by default (MediaVarDevice) and on the host with ANALISI_DEVICE_BLOCKS=0
    r = subprocess.run([cli, "-i", path, "-g", "100", "-F", "0.0", "4.0", "-S", str(S), "-s", "8"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), timeout=900,
                       env=dict(os.environ, ANALISI_DEVICE_BLOCKS=device_blocks))
    assert r.returncode == 0, r.stderr[-2000:]
    gold = open(os.path.join(GOLDEN, "cli_" + name + ".txt")).read()
    assert r.stdout.rstrip("\n") == gold.rstrip("\n")   # `$(...)` strips trailing newlines in the script


def test_cli_c1_bundled_trajectory(host, tmp_path):
    """BASELINE.json configs[0]: analisi -i tests/data/lammps.bin -g 200 -F 0.7 3.5 (56 atoms, 7958 frames,
    20 blocks of 378 steps, 378 lags: 8.96e9 pair evaluations; 1064 s with the reference on 8 cores).
    First and last block are recomputed by the oracle; the text rows of lag 0 and lag 377 must carry the
    block-averaged values of an independent numpy Welford over the python-side block results."""
    cli, pa = host
    path = os.path.join(REFDATA, "lammps.bin")
    if not os.path.exists(path):
        pytest.skip("tests/_refdata/lammps.bin not present")
    r = subprocess.run([cli, "-i", path, "-g", "200", "-F", "0.7", "3.5"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, cwd=str(tmp_path), timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    got = _rows(r.stdout)
    tr = pa.Traj(path)
    tr.setWrapPbc(True)
    nts, n_b, nbin = tr.getNtimesteps(), 20, 200
    g = pa.Gofrt_lammps(tr, 0.7, 3.5, nbin, 0, 2, 1, 1, False)
    nextra = g.getNumberOfExtraTimestepsNeeded(n_b)
    s = (nts - nextra) // n_b
    assert (nts, nextra, s) == (7958, 379, 378)
    assert len(got) == s * nbin
    tr.setAccessWindowSize(s + nextra)
    blocks = []
    for ib in range(n_b):
        tr.setAccessStart(ib * s)
        g.reset(s)
        g.calculate(ib * s)
        blocks.append(np.array(g, copy=True))
        if ib in (0, n_b - 1):   # oracle on a subset of lags of this block (full block: 4.5e8 pair evaluations)
            c = oracle.counts(tr.get_positions_copy(), tr.get_box_copy(), tr.get_type_ids(), 0.7, 3.5, nbin, 3, s,
                              primo=ib * s, skip=1, ntypes=2, first_frame=ib * s, total_frames=nts)
            assert np.array_equal(g.counts()[:3], c)
    mean, var = oracle.mediavar(np.array(blocks))
    want = _cli_text(mean, var, s, 1, nbin, 6)
    assert got == want


@pytest.mark.parametrize("format2020", [False, True])
def test_cli_device_side_ingest_equals_host_reader(host, tmp_path, format2020):
    """The window parsed on the GPUs from the raw dump records (agofrt_traj_upload_records: ids shuffled in every frame,
    three chunks per frame, both header flavours) and the window parsed by the host reader give the same output, with
    blocks one by one and as a batch; a trajectory in which an atom changes type falls back to the host reader, which
    warns like the reference (lib/src/trajectory.cpp:640-646)."""
    cli, _ = host
    pos, box, types = synth.small_case(23, (5, 4, 3), 1.05, 2, True, 50, "parity")
    path = str(tmp_path / "d.bin")
    synth.write_lammps_binary(path, pos, box, types * 2 + 1, nchunk=3, shuffle_seed=11, format2020=format2020,
                              ids=np.arange(pos.shape[1]) * 2 + 7)
    argv = ["-i", path, "-g", "30", "-F", "0.0", "2.4", "-S", "4", "-s", "2", "-B", "5"]

    def run(**env):
        r = subprocess.run([cli] + argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), timeout=600,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-2000:]
        return r
    base = run(ANALISI_DEVICE_PARSE="0", ANALISI_BLOCK_BATCH="0")
    assert "parsed on the GPUs" not in base.stderr
    for env in (dict(ANALISI_DEVICE_PARSE="1", ANALISI_BLOCK_BATCH="0"), dict(ANALISI_DEVICE_PARSE="1", ANALISI_BLOCK_BATCH="1"),
                dict(ANALISI_DEVICE_PARSE="0", ANALISI_BLOCK_BATCH="1"), dict(ANALISI_DEVICE_PARSE="1", ANALISI_PREFETCH="0")):
        r = run(**env)
        assert r.stdout == base.stdout, env
        assert ("parsed on the GPUs" in r.stderr) == (env["ANALISI_DEVICE_PARSE"] == "1")
        if "ANALISI_BLOCK_BATCH" in env:   # (unset: one batch only when the process drives several GPUs)
            assert ("as one batch on the GPUs" in r.stderr) == (env["ANALISI_BLOCK_BATCH"] == "1")
    # an atom changes type in frame 20
    raw = (types * 2 + 1).astype(np.int64)
    path2 = str(tmp_path / "e.bin")
    with open(path2, "wb") as out:
        for f in range(pos.shape[0]):
            t = raw.copy()
            if f >= 20:
                t[5] = 9 - t[5] if t[5] in (1, 3) else t[5]
                t[5] = {1: 3, 3: 1}[int(raw[5])]
            one = str(tmp_path / "one.bin")
            synth.write_lammps_binary(one, pos[f:f + 1], box[f:f + 1], t, format2020=format2020, first_step=f)
            out.write(open(one, "rb").read())
    argv[1] = path2
    a = run(ANALISI_DEVICE_PARSE="0", ANALISI_BLOCK_BATCH="0")
    b = run(ANALISI_DEVICE_PARSE="1", ANALISI_BLOCK_BATCH="0")
    assert a.stdout == b.stdout
    assert "WARNING: atomic type for atom with id 5 is changing" in a.stderr and "WARNING: atomic type for atom with id 5 is changing" in b.stderr


def test_cli_c1_against_the_reference_itself(host, tmp_path):
    """BASELINE.json configs[0] against the UNMODIFIED reference: tests/golden/c1_reference.* holds what the compiled
    reference's own BlockAverage<Gofrt> chain produced for `analisi -i tests/data/lammps.bin -g 200 -F 0.7 3.5`
    (756 s on one core; tests/golden/make_c1_golden.py).  All 378 x 6 x 200 means and variances of the GPU chain --
    CLI text and the python block-average object -- are compared with it."""
    import hashlib
    import json
    cli, pa = host
    path = os.path.join(REFDATA, "lammps.bin")
    if not os.path.exists(path):
        pytest.skip("tests/_refdata/lammps.bin not present")
    gold = np.load(os.path.join(GOLDEN, "c1_reference.npz"))
    meta = json.load(open(os.path.join(GOLDEN, "c1_reference.json")))
    n_b, s, nbin = 20, 378, 200
    S1, S2 = gold["S1"].astype(np.float64), gold["S2"].astype(np.float64)
    ref_mean = S1 / (n_b * s)
    ref_var = (S2 - S1 * S1 / n_b) / (n_b * (n_b - 1) * s ** 2)
    # (the reconstruction reproduces the reference's doubles: checked on the lags stored in full)
    for k, lag in enumerate(gold["pin_lags"]):
        np.testing.assert_allclose(ref_mean[lag], gold["pin_mean"][k], rtol=1e-12, atol=0)
        np.testing.assert_allclose(ref_var[lag], gold["pin_var"][k], rtol=1e-10, atol=1e-18)

    # ---- the python object: full-precision means and variances of the mean
    tr = pa.Traj(path)
    tr.setWrapPbc(True)
    ba = pa.GofrtBlockAverage_lammps(tr, n_b)
    ba.calculate(0.7, 3.5, nbin, 0, 1, 1, 1, False)
    mean, var = ba.mean(), ba.variance()
    assert mean.shape == (s, 6, nbin) == tuple(meta["shape"])
    # the reference adds incr = 1/378 once per counted pair, the GPU path multiplies the count by incr: both are
    # within a few ulp of count/378, the difference must stay at rounding level
    np.testing.assert_allclose(mean, ref_mean, rtol=2e-12, atol=0)
    np.testing.assert_allclose(var, ref_var, rtol=1e-9, atol=1e-16)
    for k, lag in enumerate(gold["pin_lags"]):
        np.testing.assert_allclose(mean[lag], gold["pin_mean"][k], rtol=2e-12, atol=0)
        np.testing.assert_allclose(var[lag], gold["pin_var"][k], rtol=1e-9, atol=1e-16)
    # integer sufficient statistics: sum of the block counts, exactly
    assert np.array_equal(np.round(mean * (n_b * s)), S1)

    # ---- the CLI text: 6 significant digits per value; a value within 1e-12 of a rounding tie may print differently
    r = subprocess.run([cli, "-i", path, "-g", "200", "-F", "0.7", "3.5"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, cwd=str(tmp_path), timeout=1500)
    assert r.returncode == 0, r.stderr[-2000:]
    got = _rows(r.stdout)
    want = _cli_text(ref_mean, ref_var, s, 1, nbin, 6)
    assert len(got) == len(want) == s * nbin
    differing = [(a, b) for a, b in zip(got, want) if a != b]
    assert len(differing) <= 8, differing[:3]
    for a, b in differing:
        np.testing.assert_allclose(np.array(a.split(), dtype=float), np.array(b.split(), dtype=float), rtol=2e-6, atol=0)
    header = [l for l in r.stdout.split("\n") if l.startswith("#")]
    assert header == [l for l in meta["columns_description"].split("\n") if l.startswith("#")]
    if not differing:
        # identical rows: then the whole text is the reference's, byte for byte
        text = meta["columns_description"].rstrip("\n") + "\n" + "".join(
            l + "\n" + ("\n" if (i + 1) % nbin == 0 else "") for i, l in enumerate(got))
        assert hashlib.sha256(text.encode()).hexdigest() == meta["cli_text_sha256"]


@pytest.mark.parametrize("device_blocks", ["1", "0"], ids=["MediaVarDevice", "MediaVar"])
@pytest.mark.parametrize("kind", ["numpy", "lammps"])
def test_block_average_binding_vs_oracle(host, tmp_path, kind, device_blocks, monkeypatch):
    """GofrtBlockAverage (BlockAverageG<TR, Gofrt> from python; TraiettoriaF<Trajectory_numpy> is this
    repository's addition): mean and variance of the mean over blocks, bit-identical to the oracle's MediaVar
    over count*incr blocks (power-of-two incr: the blocks themselves are exact)."""
    _, pa = host
    monkeypatch.setenv("ANALISI_DEVICE_BLOCKS", device_blocks)   # read by BlockAverageG::calculate
    nfr, n_b, lmax, skip, nbin = 41, 4, 3, 2, 24
    pos, box, types = synth.small_case(33, (5, 4, 4), 1.08, 2, True, nfr)
    bi = synth.lammps_rows_to_internal(box)
    if kind == "numpy":
        tr = pa.Trajectory(pos, np.zeros_like(pos), types.astype(np.int32), box, pa.BoxFormat.LammpsTriclinic, True, False)
        ba = pa.GofrtBlockAverage(tr, n_b)
    else:
        path = str(tmp_path / "b.bin")
        synth.write_lammps_binary(path, pos, box, types + 1, format2020=True, nchunk=3, shuffle_seed=5)
        tr = pa.Traj(path)
        tr.setLoadVelocities(False)
        tr.setWrapPbc(True)
        ba = pa.GofrtBlockAverage_lammps(tr, n_b)
    ba.calculate(0.0, 2.4, nbin, lmax, 1, skip, 1, False)
    nextra = oracle.nextra(nfr, n_b, lmax)
    s = (nfr - nextra) // n_b
    assert ba.block_size() == s == 9 and nextra == 3
    wrapped = oracle.pbc_wrap(pos, bi)
    blocks = []
    for ib in range(n_b):
        c = oracle.counts(wrapped, bi, types, 0.0, 2.4, nbin, lmax, s, primo=ib * s, skip=skip, ntypes=2)
        blocks.append(c * cabi.gofrt_incr(s, skip))   # incr = 1/4
    mean, var = oracle.mediavar(np.array(blocks))
    assert np.array_equal(ba.mean(), mean)
    assert np.array_equal(ba.variance(), var)
    # the calculation object is left holding its last block, whichever class consumed the blocks
    assert np.array_equal(np.asarray(ba.last_block()), blocks[-1])
    st = ba.stats()
    assert st["blocks"] == n_b and st["pair_evals"] == n_b * lmax * 5 * len(types) ** 2


def test_python_helpers_compute_gofr(host):
    """compute_gofr over pyanalisi.Trajectory: segments, histogram vs normalised output"""
    _, pa = host
    from analisi_b200 import analysis
    pos, box, types = synth.small_case(61, (4, 4, 4), 1.1, 1, False, 24)
    tr = pa.Trajectory(pos, np.zeros_like(pos), types.astype(np.int32), box, pa.BoxFormat.LammpsOrtho, True, False)
    h = analysis.compute_gofr(tr, 0.0, 2.0, 20, start=0, stop=24, tmax=4, tskip=4, return_histogram=True)
    bi = synth.lammps_rows_to_internal(box)
    ref = oracle.counts(oracle.pbc_wrap(pos, bi), bi, types, 0.0, 2.0, 20, 4, 20, primo=0, skip=4, ntypes=1)
    assert np.array_equal(h, ref * cabi.gofrt_incr(20, 4))
    g = analysis.compute_gofr(tr, 0.0, 2.0, 20, start=0, stop=24, tmax=4, tskip=4)
    assert np.allclose(g, analysis.hist2gofr(20, 0.1, 0.0, h))
    segs = analysis.compute_gofr(tr, 0.0, 2.0, 20, start=0, stop=24, tmax=4, tskip=2, n_segments=2, return_histogram=True)
    assert len(segs) == 2
    ref1 = oracle.counts(oracle.pbc_wrap(pos, bi), bi, types, 0.0, 2.0, 20, 4, 10, primo=10, skip=2, ntypes=1)
    assert np.array_equal(segs[1], ref1 * cabi.gofrt_incr(10, 2))


@pytest.mark.parametrize("nth,tmax", [(1, 1), (3, 1), (3, 4), (1, 4)])
def test_reference_cpp_fixture_gofr(host, nth, tmax):
    """reference tests/src/test_lammps2020.cpp:27-90 (GofrFixture<NTH,TMAX>): Gofrt<double,Trajectory>(traj, 0.9, 2.0,
    20, TMAX, NTH, skip=70) on lammps2020.bin (4000 atoms, wrap OFF, window of 150 frames at 0); reset(75-primo);
    calculate(primo) for primo = 0 and 13.  The reference's golden files for it hold `inf`, so
    the oracle is the checker; the result must not depend on the thread-count argument."""
    _, pa = host
    path = os.path.join(REFDATA, "lammps2020.bin")
    if not os.path.exists(path):
        pytest.skip("tests/_refdata/lammps2020.bin not present")
    tr = pa.Traj(path)
    tr.setWrapPbc(False)
    tr.setAccessWindowSize(150)
    tr.setAccessStart(0)
    g = pa.Gofrt_lammps(tr, 0.9, 2.0, 20, tmax, nth, 70, 1, False)
    pos, box, ids = tr.get_positions_copy(), tr.get_box_copy(), tr.get_type_ids()
    for primo in (0, 13):
        g.reset(75 - primo)
        g.calculate(primo)
        ref = oracle.counts(pos, box, ids, 0.9, 2.0, 20, tmax, 75 - primo, primo=primo, skip=70, ntypes=2,
                            total_frames=tr.getNtimesteps())
        assert np.array(g).shape == (min(tmax, 75 - primo), 6, 20)
        assert np.array_equal(g.counts(), ref)
        assert np.array_equal(np.array(g), ref * cabi.gofrt_incr(75 - primo, 70))   # incr = 1
        assert ref[:, :3].sum() > 0 and ref[:, 3:].sum() == 0   # rmin 0.9: no self pairs at these lags


def test_edge_pairs_through_host_layers(host, ctx, tmp_path):
    """north_star: "any pair within 1 ulp of a bin edge is reported separately" -- Gofrt.setReportEdges / edge_pairs()
    and the CLI's --edge-pairs, against the oracle's count.  Random liquids almost never hold such a pair, so a few
    are built: atoms on the x axis at distances whose SQUARE is exactly a bin threshold (or the double below it)."""
    cli, pa = host
    rmin, rmax, nbin = 0.0, 2.0, 16
    types = np.zeros(2, dtype=np.int32)
    probe = cabi.DeviceTrajectory(ctx, 2, 6, types, 1, 1)
    plan = cabi.Plan(probe, rmin, rmax, nbin)
    thr = plan.thresholds()
    plan.close()
    probe.close()
    xs = []
    for k in range(1, nbin):
        for target in (thr[k], np.nextafter(thr[k], 0.0)):
            x0 = np.sqrt(target)
            for x in (x0, np.nextafter(x0, 0.0), np.nextafter(x0, 4.0)):
                if x * x == target:
                    xs.append(x)
                    break
    assert len(xs) >= 6   # not every threshold is a representable square; a handful is plenty
    pos1 = np.zeros((1 + len(xs), 3))
    pos1[1:, 0] = xs
    pos1 += np.array([0.25, 3.0, 3.0])
    nfr = 5
    pos = np.ascontiguousarray(np.repeat(pos1[None], nfr, axis=0))
    box = np.tile([0.0, 6.0, 0.0, 6.0, 0.0, 6.0], (nfr, 1))
    types = np.zeros(pos.shape[1], dtype=np.int32)
    tr = pa.Trajectory(pos, np.zeros_like(pos), types, box, pa.BoxFormat.LammpsOrtho, False, False)
    gof = pa.Gofrt(tr, rmin, rmax, nbin, 2, 1, 1, 1, False)
    gof.setReportEdges(True)
    gof.reset(3)
    gof.calculate(0)
    bi = synth.lammps_rows_to_internal(box)
    ref, eref = oracle.counts(pos, bi, types, rmin, rmax, nbin, 2, 3, ntypes=1, return_edges=True)
    assert np.array_equal(gof.counts(), ref)
    assert gof.edge_pairs() == eref
    assert eref >= 2 * 2 * 3 * len(xs) // 2   # (0,k) and (k,0), two lags, three origins; some x are "just below" ones
    gof.setReportEdges(False)
    gof.calculate(0)
    assert np.array_equal(gof.counts(), ref) and gof.edge_pairs() == 0
    path = str(tmp_path / "edge.bin")
    synth.write_lammps_binary(path, pos, box, types + 1)
    r = subprocess.run([cli, "-i", path, "-g", str(nbin), "-F", str(rmin), str(rmax), "-S", "1", "-B", "2", "--edge-pairs"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stderr.splitlines() if l.startswith("pairs within 1 ulp of a bin edge")]
    assert len(lines) == 2 and all(int(l.split(":")[1].split()[0]) > 0 for l in lines)


def test_cli_neighbour_reference_golden_text(host, tmp_path):
    """reference tests/test_cli.sh:35 verbatim: analisi -i lammps2020.bin --neighbour 10, stdout against the
    reference's golden text (tests/golden/cli_neighbours.txt); and the same numbers through the python class."""
    cli, pa = host
    path = os.path.join(REFDATA, "lammps2020.bin")
    if not os.path.exists(path):
        pytest.skip("tests/_refdata/lammps2020.bin not present")
    r = subprocess.run([cli, "-i", path, "--neighbour", "10"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    gold = open(os.path.join(GOLDEN, "cli_neighbours.txt")).read()
    assert r.stdout.rstrip("\n") == gold.rstrip("\n")
    tr = pa.Traj(path)
    tr.setLoadVelocities(False)
    tr.setWrapPbc(True)
    s = (tr.getNtimesteps() - 1) // 20
    tr.setAccessWindowSize(s)
    h = pa.NeighbourHistogram_lammps(tr, 10.0, 1, 1)
    h.reset(s)
    for i in range(20):
        tr.setAccessStart(s * i)
        h.calculate(s * i)
    rows = []
    for ty in range(tr.get_ntypes()):
        rows.append('"%d"' % ty)
        rows += ["%d %d" % kv for kv in sorted(h.get_hist(ty).items())]
        rows += ["", ""]
    assert ("\n".join(rows) + "\n").rstrip("\n") == gold.rstrip("\n")


@pytest.mark.parametrize("name,args", [("MSD_normal_full", ["-Q"]), ("MSD_normal", ["-Q", "-s", "10", "-S", "50"]),
                                       ("MSD_cm", ["-q", "-s", "10", "-S", "50"]),
                                       ("MSD_cm_reference", ["-Q", "--mean-square-displacement-self"])])
def test_cli_msd_reference_golden_text(host, tmp_path, name, args):
    """reference tests/test_cli.sh:24-27 verbatim: analisi -i lammps2020.bin -Q | -Q -s 10 -S 50 | -q -s 10 -S 50 |
    -Q --mean-square-displacement-self, stdout against the reference's golden text."""
    cli, _ = host
    path = os.path.join(REFDATA, "lammps2020.bin")
    if not os.path.exists(path):
        pytest.skip("tests/_refdata/lammps2020.bin not present")
    r = subprocess.run([cli, "-i", path] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path),
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    gold = open(os.path.join(GOLDEN, "cli_%s.txt" % name)).read()
    assert r.stdout.rstrip("\n") == gold.rstrip("\n")


def test_pyanalisi_msd(host):
    """pyanalisi.MeanSquareDisplacement(traj, skip, tmax, nthreads, cm_msd, cm_self, debug) as the reference's
    Analysis.compute_msd calls it (pyanalisi/analysis.py:91-98) against the oracle"""
    _, pa = host
    pos, box, types = synth.small_case(96, (6, 6, 5), 1.1, 2, True, 40)
    tr = pa.Trajectory(pos, np.zeros_like(pos), types.astype(np.int32), box, pa.BoxFormat.LammpsTriclinic, False, False)
    msd = pa.MeanSquareDisplacement(tr, 10, 12, 4, True, False, False)
    msd.reset(20)
    msd.calculate(3)
    v = np.array(msd, copy=True)
    assert v.shape == (12, 2, 2)
    ref = oracle.msd(pos, types, 20, 12, primo=3, skip=10, cm_msd=True, ntypes=2)
    assert (np.abs(v[:, 0] - ref[:, 0]) <= 1e-12 * np.maximum(np.abs(ref[:, 0]), 1e-300)).all()
    assert np.array_equal(v[:, 1], ref[:, 1])
    with pytest.raises(RuntimeError, match="trajectory is too short"):
        msd.reset(30)
        msd.calculate(5)
    from analisi_b200 import analysis
    w = analysis.compute_msd(tr, start=2, stop=40, tmax=8, tskip_msd=5)
    ref2 = oracle.msd(pos, types, 30, 8, primo=2, skip=5, cm_msd=True, ntypes=2)
    assert w.shape == (8, 2, 2) and np.allclose(w, ref2, rtol=1e-12, atol=0) and np.array_equal(w[:, 1], ref2[:, 1])


@pytest.mark.parametrize("name", sorted(PAIR_LOOP_CASES))
def test_pyanalisi_neighbours_and_spherical_density(host, name):
    """Neighbours (reference pyanalisi.cpp:193-236: calculate_neigh, get_neigh, get_sann, get_sann_idx) and
    SphericalBase::calc through the host classes, against the fixtures of the compiled reference -- lists, SANN counts,
    spherical-harmonic densities bit for bit."""
    _, pa = host
    d = pair_loop_case(name)
    pos = np.ascontiguousarray(d["pos"])
    fmt = pa.BoxFormat.LammpsTriclinic if d["tri"] else pa.BoxFormat.LammpsOrtho
    tr = pa.Trajectory(pos, np.zeros_like(pos), d["types"].astype(np.int32), d["box"], fmt, True, False)
    nn = pa.Neighbours(tr, [tuple(s) for s in d["spec"]])
    nn.calculate_neigh(d["frame"], True)
    for i in range(0, pos.shape[1], 5):
        for jt in range(d["ntypes"]):
            c = int(d["sorted_counts"][i, jt])
            assert np.array_equal(nn.get_neigh(i, jt), d["sorted_r"][i, jt, :c])
            assert np.array_equal(nn.get_neigh_idx(i, jt), d["sorted_idx"][i, jt, :c])
            sn = int(d["sann_n"][i, jt])
            assert nn.get_sann(i, jt).shape == (sn, 4) and np.array_equal(nn.get_sann_idx(i, jt), d["sorted_idx"][i, jt, :sn])
    with pytest.raises(RuntimeError, match="Too many neighbours in shell"):
        pa.Neighbours(tr, [(1, s[1], s[2]) for s in d["spec"]]).calculate_neigh(d["frame"], False)
    res, cnt = pa.spherical_harmonic_density(tr, d["lmax"], d["nbin"], [tuple(r) for r in d["rminmax"]], d["frame"])
    assert np.array_equal(cnt, d["sh_counter"]) and np.array_equal(res, d["sh"])
