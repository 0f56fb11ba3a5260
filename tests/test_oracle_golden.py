"""CPU: pin the oracle (oracle/gofrt_oracle.c) against the reference's golden vectors and against
outputs of the compiled reference (fixtures made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import oracle
from conftest import LIVE_CASES, REFERENCE, TILE_CASES, live_case, load_golden, tile_case


def test_gofr_numpy_golden():
    """reference tests/test_gofrt.py -> tests/test_gofrt/test_gofr.csv (ortho, unwrapped, 3 types, 10 lags)."""
    z = load_golden("gofr_numpy.npz")
    rmin, rmax, nbin, tmax, skip, nts = z["params"]
    c = oracle.counts(z["pos"], z["box_internal"], z["types"], rmin, rmax, int(nbin), int(tmax), int(nts),
                      primo=0, skip=int(skip), ntypes=3, total_frames=7958)
    assert c.shape == (10, 12, 200)
    assert np.array_equal(c, z["counts"])
    # the float accumulation of the reference with its 4 threads reproduces the CSV to rounding
    v = oracle.vdata(z["pos"], z["box_internal"], z["types"], rmin, rmax, int(nbin), int(tmax), int(nts),
                     primo=0, skip=int(skip), ntypes=3, ref_nthreads=4, total_frames=7958)
    assert np.abs(v - z["csv"]).max() < 1e-11


def test_gofr_notebook_golden():
    """reference tests/test_notebook.py -> tests/test_notebook/test_gofr.csv (mmap trajectory, wrap on)."""
    z = load_golden("gofr_notebook.npz")
    wrapped = oracle.pbc_wrap(z["pos_unwrapped"], z["box_internal"])
    assert np.array_equal(wrapped, z["pos_wrapped"])
    rmin, rmax, nbin, tmax, skip, nts = z["params"]
    nt = int(z["types"].max()) + 1
    c = oracle.counts(wrapped, z["box_internal"], z["types"], rmin, rmax, int(nbin), int(tmax), int(nts),
                      primo=0, skip=int(skip), ntypes=nt)
    assert np.array_equal(c, z["counts"])


def test_min_image_and_pbc_golden():
    """reference tests/src/test_trajectory.cpp:21-59 -> cpp_regression_data/{min_image,pbc_1,pbc_2} (tolerance 1e-10 there)."""
    z = load_golden("min_image_pbc.npz")
    for tag in ("1", "2"):
        w = oracle.pbc_wrap(z["pos_" + tag][None], z["box_" + tag][None])[0]
        g = z["pbc_" + tag]
        assert np.abs(w - g).max() <= 1e-10 * max(1.0, np.abs(g).max())
    assert np.array_equal(oracle.pbc_wrap(z["pos_1"][None], z["box_1"][None])[0], z["pos_wrapped_1"])
    d = oracle.d2_all(z["pos_wrapped_1"], z["pos_wrapped_1"], z["box_1"])
    g = z["min_image_1"]
    assert np.abs(d - g).max() <= 1e-10 * max(1.0, np.abs(g).max())
    lh = z["box_1"][3:6]
    assert (np.abs(d[..., :3]) <= lh + 1e-15).all()


@pytest.mark.parametrize("name", LIVE_CASES)
def test_live_reference_cases(name):
    """triclinic / NPT / unwrapped / ragged loops: fixtures computed by the compiled reference."""
    d = live_case(name)
    rmin, rmax, nbin, tmax, skip, every, nts, primo = d["params"]
    box_internal = d["box_internal"]
    ids, nt = oracle.type_ids(d["raw_types"])
    assert np.array_equal(ids, d["type_ids"])
    from analisi_b200 import synth
    bi = synth.lammps_rows_to_internal(d["box_lammps"])
    assert np.array_equal(bi, box_internal)
    pos = d["pos_in"]
    if bool(d["wrap"]):
        pos = oracle.pbc_wrap(pos, box_internal)
    assert np.array_equal(pos, d["pos_ref"])
    c, edges = oracle.counts(pos, box_internal, ids, rmin, rmax, int(nbin), int(tmax), int(nts), primo=int(primo),
                             skip=int(skip), every=int(every), ntypes=nt, return_edges=True)
    assert np.array_equal(c, d["counts"])
    v = oracle.vdata(pos, box_internal, ids, rmin, rmax, int(nbin), int(tmax), int(nts), primo=int(primo),
                     skip=int(skip), every=int(every), ntypes=nt, ref_nthreads=3)
    assert np.array_equal(v, d["vdata"])  # bitwise: same thread split, same order of additions


def test_oracle_matches_compiled_reference_live():
    """When oracle/_ref is present: a fresh random triclinic case, reference vs restatement."""
    m = oracle.load_ref()
    if m is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from analisi_b200 import synth
    pos, box, types = synth.small_case(77, (5, 4, 3), 1.05, 2, True, 9, "parity", (0.2, 0.1, -0.15), True)
    tr = m.Trajectory(pos, np.zeros_like(pos), types, box, m.BoxFormat.LammpsTriclinic, True, False)
    g = m.Gofrt(tr, 0.1, 2.3, 17, 4, 2, 2, 1, False)
    g.reset(5)
    g.calculate(1)
    v = np.array(g, copy=True)
    bi = synth.lammps_rows_to_internal(box)
    w = oracle.pbc_wrap(pos, bi)
    assert np.array_equal(w, tr.get_positions_copy())
    c = oracle.counts(w, bi, types, 0.1, 2.3, 17, 4, 5, primo=1, skip=2, ntypes=2)
    assert np.array_equal(c, np.rint(v * 2).astype(np.uint64))


def _cli_text(mean, var, leff, nbin, ncol):
    # the CLI's printing loop, reference analisi/main.cpp:564-584
    lines = []
    for t in range(leff):
        for r in range(nbin):
            row = "%d %d" % (t, r)
            for k in range(ncol):
                row += " %s %s" % (_g(mean[t, k, r]), _g(var[t, k, r]))
            lines.append(row)
        lines.append("")
    return lines


def _g(x):
    # iostream default formatting of a double: %g with 6 significant digits
    return "%g" % x


@pytest.mark.parametrize("name,tmax", [("pair_corr_no_t", 1), ("pair_corr_t", 10)])
def test_cli_golden_blocks(name, tmax):
    """reference tests/test_cli.sh: analisi -i lammps2020.bin -g 100 -F 0.0 4.0 -S {1,10} -s 8
    (20 blocks, mean and variance, text).  Needs the reference tree for the 51 MB input."""
    m = oracle.load_ref()
    path = os.path.join(REFERENCE, "tests/data/lammps2020.bin")
    if m is None or not os.path.exists(path):
        pytest.skip("reference tree not available")
    gold = open(os.path.join(REFERENCE, "tests/data/cli", name)).read().split("\n")
    gold_rows = [l for l in gold if l and not l.startswith("#")]
    tr = m.Traj(path)
    tr.setWrapPbc(True)
    nts = tr.get_ntimesteps()
    n_b, nbin, skip = 20, 100, 8
    nextra = oracle.nextra(nts, n_b, tmax)
    s = (nts - nextra) // n_b
    tr.setAccessWindowSize(s + nextra)
    blocks = []
    for ib in range(n_b):
        tr.setAccessStart(ib * s)
        pos = tr.get_positions_copy()
        box = tr.get_box_copy()
        ids = tr.get_type_ids()
        nt = int(tr.get_ntypes())
        v = oracle.vdata(pos, box, ids, 0.0, 4.0, nbin, tmax, s, primo=ib * s, skip=skip, ntypes=nt,
                         first_frame=ib * s, total_frames=nts, ref_nthreads=2)
        blocks.append(v)
    mean, var = oracle.mediavar(np.array(blocks))
    leff = oracle.leff(s, tmax)
    rows = [l for l in _cli_text(mean, var, leff, nbin, nt * (nt + 1)) if l]
    assert len(rows) == len(gold_rows)
    assert rows == gold_rows


def neighbour_text(hist):
    # the CLI's printing loop, reference analisi/main.cpp:636-642
    lines = []
    for k in range(hist.shape[0]):
        lines.append('"%d"' % k)
        for c in np.nonzero(hist[k])[0]:
            lines.append("%d %d" % (c, hist[k][c]))
        lines += ["", ""]
    return "\n".join(lines) + "\n"


def test_neighbour_histogram_cli_golden():
    """reference tests/test_cli.sh:35: analisi -i lammps2020.bin --neighbour 10 (20 blocks of 9 frames, wrap on):
    the oracle's restatement of IstogrammaAtomiRaggio reproduces the reference's golden text."""
    m = oracle.load_ref()
    path = os.path.join(REFERENCE, "tests/data/lammps2020.bin")
    if m is None or not os.path.exists(path):
        pytest.skip("reference tree not available")
    from conftest import GOLDEN
    tr = m.Traj(path)
    tr.setWrapPbc(True)
    nt, n_b = tr.get_ntimesteps(), 20
    s = (nt - 1) // n_b
    tr.setAccessWindowSize(s)
    hist = None
    for i in range(n_b):
        tr.setAccessStart(s * i)
        hist = oracle.neighbour_hist(tr.get_positions_copy(), tr.get_box_copy(), tr.get_type_ids(), 10.0, s * i, s, 1,
                                     ntypes=int(tr.get_ntypes()), first_frame=s * i, hist=hist)
    assert hist.sum() == 2 * 4000 * n_b * s
    gold = open(os.path.join(GOLDEN, "cli_neighbours.txt")).read()
    assert neighbour_text(hist).rstrip("\n") == gold.rstrip("\n")


MSD_CASES = [("MSD_normal_full", 1, 0, True, False), ("MSD_normal", 10, 50, True, False), ("MSD_cm", 10, 50, False, False),
             ("MSD_cm_reference", 1, 0, True, True)]   # name, -s, -S, -Q (centre-of-mass rows), --mean-square-displacement-self


def msd_text(mean, var):
    # the CLI's printing loop, reference analisi/main.cpp:541-546: "mean var " per column, row per lag
    return ["".join("%s %s " % (_g(m), _g(v)) for m, v in zip(mean[i].ravel(), var[i].ravel())) for i in range(mean.shape[0])]


@pytest.mark.parametrize("name,skip,lmax,cm_msd,cm_self", MSD_CASES)
def test_msd_cli_golden(name, skip, lmax, cm_msd, cm_self):
    """reference tests/test_cli.sh:24-27 on lammps2020.bin (unwrapped, 20 blocks): the oracle's restatement of
    MSD<T>::calc_single_th + MediaVar reproduces the reference's golden text for all four flag combinations."""
    m = oracle.load_ref()
    path = os.path.join(REFERENCE, "tests/data/lammps2020.bin")
    if m is None or not os.path.exists(path):
        pytest.skip("reference tree not available")
    from conftest import GOLDEN
    tr = m.Traj(path)
    tr.setWrapPbc(False)
    nts, n_b = tr.get_ntimesteps(), 20
    nextra = oracle.nextra(nts, n_b, lmax)
    s = (nts - nextra) // n_b
    tr.setAccessWindowSize(s + nextra)
    blocks = []
    for ib in range(n_b):
        tr.setAccessStart(ib * s)
        blocks.append(oracle.msd(tr.get_positions_copy(), tr.get_type_ids(), s, lmax, primo=ib * s, skip=skip, cm_msd=cm_msd,
                                 cm_self=cm_self, ntypes=int(tr.get_ntypes()), first_frame=ib * s, total_frames=nts))
    mean, var = oracle.mediavar(np.array(blocks))
    gold = open(os.path.join(GOLDEN, "cli_%s.txt" % name)).read().rstrip("\n").split("\n")
    assert msd_text(mean, var) == gold


@pytest.mark.parametrize("name", sorted(TILE_CASES))
def test_live_reference_tile_cases(name):
    """A few thousand atoms (C3- and C2-shaped, reduced) through the compiled reference: the oracle reproduces its
    wrapped positions and its counts."""
    import hashlib
    d = tile_case(name)
    rmin, rmax, nbin, tmax, skip, every, nts, primo = d["params"]
    from analisi_b200 import synth
    bi = synth.lammps_rows_to_internal(d["box_lammps"])
    assert np.array_equal(bi, d["box_internal"])
    pos = oracle.pbc_wrap(d["pos_in"], bi)
    sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pos).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, d["pos_ref_sha256"])
    ids = d["type_ids"]
    c = oracle.counts(pos, bi, ids, rmin, rmax, int(nbin), int(tmax), int(nts), primo=int(primo), skip=int(skip),
                      every=int(every), ntypes=int(ids.max()) + 1)
    assert np.array_equal(c, d["counts"])


def test_c4_subset_golden_is_self_consistent():
    """tests/golden/c4_subset_counts.*: the stored lag rows reproduce the checksum bench.py compares with (the rows
    between them are empty: every 8th lag), and the self row of lag 0 holds one count per atom and origin."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    meta = json.load(open(os.path.join(GOLDEN, "c4_subset_counts.json")))
    z = np.load(os.path.join(GOLDEN, "c4_subset_counts.npz"))
    full = np.zeros((meta["subset"]["leff"], 2, 500), dtype=np.uint64)
    full[z["lags"]] = z["counts"]
    assert hashlib.sha256(full.astype("<u8").tobytes()).hexdigest() == meta["counts_sha256"]
    assert int(full.sum()) == meta["counts_sum"]
    assert int(full[0, 1, 0]) == 100000 * 8 and int(full[0, 1, 1:].sum()) == 0
    assert list(z["lags"]) == list(range(0, 201, 8))


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_full_n_golden_oracle_equals_the_reference(name):
    """tests/golden/full_n_counts.*: BASELINE.json configs[1] and configs[2] at their FULL atom counts (4096 cubic, 12 288
    triclinic with 3 types) on a few (lag, origin) jobs, counts of the unmodified reference
    (tests/golden/make_full_n_golden.py).  The oracle reproduces them; the GPU test of the same name compares the tile
    kernel with them."""
    import hashlib
    import json
    from analisi_b200 import synth
    from conftest import GOLDEN
    meta = json.load(open(os.path.join(GOLDEN, "full_n_counts.json")))[name]
    gold = np.load(os.path.join(GOLDEN, "full_n_counts.npz"))[name + "/counts"]
    w = synth.WORKLOADS[name]
    assert meta["workload"] == w.name
    pos, box, types = synth.generate(w, nframes=meta["frames"])
    assert hashlib.sha256(np.ascontiguousarray(pos).tobytes()).hexdigest() == meta["pos_in_sha256"], \
        "synth.generate no longer reproduces the fixture's input: regenerate it"
    bi = synth.lammps_rows_to_internal(box)
    pos = oracle.pbc_wrap(pos, bi)
    c = oracle.counts(pos, bi, types, w.rmin, w.rmax, w.nbin, meta["tmax"], meta["ntimesteps"], primo=meta["primo"],
                      skip=meta["skip"], every=meta["every"], ntypes=w.ntypes)
    assert np.array_equal(c, gold)
    assert hashlib.sha256(np.ascontiguousarray(gold).astype("<u8").tobytes()).hexdigest() == meta["counts_sha256"]
