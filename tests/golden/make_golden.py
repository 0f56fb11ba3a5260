"""Generate the committed fixtures under tests/golden/ (run HERE, where /root/reference exists).

    python tests/golden/make_golden.py

Sources of truth, in this order:
  * the reference's own golden files for the g(r,t) path:
      tests/test_gofrt/test_gofr.csv, tests/test_notebook/test_gofr.csv,
      tests/cpp_regression_data/{min_image,pbc_1,pbc_2}
    cut down to the frames they really use so the fixtures stay small;
  * the UNMODIFIED reference compiled by oracle/Makefile (oracle/_ref/analisi_ref*.so) for what no
    golden file pins: triclinic minimum image, every>1, ragged skip, NPT boxes.
The GPU box has no /root/reference: the -m gpu tests read only these .npz files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

REF = os.environ.get("ANALISI_REFERENCE", "/root/reference")

import oracle  # noqa: E402
from analisi_b200 import synth  # noqa: E402


def ref_module():
    oracle.oracle.build_ref(REF)
    m = oracle.load_ref()
    if m is None:
        raise SystemExit("oracle/_ref could not be built/imported")
    return m


def read_csv_column(path):
    return np.loadtxt(path, delimiter=",", skiprows=1, usecols=1)


def gofr_numpy(m):
    """tests/test_gofrt.py: Gofrt(traj,0,3.8,200,10,4,10,False,1); reset(700); calculate(0)
    on positions.npy/cells.npy (CellVectors, wrap=False), 3 types 40/8/8."""
    pos = np.load(REF + "/tests/data/positions.npy")
    cells = np.load(REF + "/tests/data/cells.npy")
    types = np.zeros(pos.shape[1], dtype=np.int32)
    types[-16:-8] = 1
    types[-8:] = 2
    csv = read_csv_column(REF + "/tests/test_gofrt/test_gofr.csv").reshape(10, 12, 200)
    nfr = 700  # origins 0,10,..,690 and lags 0..9 touch frames 0..699
    vel = np.zeros_like(pos[:nfr])
    traj = m.Trajectory(np.ascontiguousarray(pos[:nfr]), vel, types, np.ascontiguousarray(cells[:nfr]),
                        m.BoxFormat.CellVectors, False, False)
    box_internal = traj.get_box_copy()
    assert np.array_equal(traj.get_positions_copy(), pos[:nfr])  # diagonal cell: no rotation
    incr = 1.0 / 70
    counts = np.rint(csv / incr).astype(np.uint64)
    assert np.abs(counts * incr - csv).max() < 1e-9
    np.savez_compressed(os.path.join(HERE, "gofr_numpy.npz"), pos=pos[:nfr], box_internal=box_internal,
                        types=types, csv=csv, counts=counts,
                        params=np.array([0.0, 3.8, 200, 10, 10, 700]))  # rmin rmax nbin tmax skip ntimesteps


def gofr_notebook(m):
    """notebooks/calc_inspector.ipynb via tests/test_notebook.py: Traj('lammps.bin'), wrap on,
    Gofrt_lammps(traj,0.5,3.8,100,1,4,10,False,1); reset(999); calculate(0).  Only lag 0 and the
    origins 0,10,..,990 are touched, so the fixture keeps those 100 frames (skip becomes 1)."""
    csv = read_csv_column(REF + "/tests/test_notebook/test_gofr.csv")
    tr = m.Traj(REF + "/tests/data/lammps.bin")
    tr.setWrapPbc(False)
    tr.setAccessWindowSize(1000)
    tr.setAccessStart(0)
    raw = tr.get_positions_copy()[0:1000:10].copy()
    box = tr.get_box_copy()[0:1000:10].copy()
    types = tr.get_type_ids()
    tw = m.Traj(REF + "/tests/data/lammps.bin")
    tw.setWrapPbc(True)
    tw.setAccessWindowSize(1000)
    tw.setAccessStart(0)
    wrapped = tw.get_positions_copy()[0:1000:10].copy()
    ntypes = int(tw.get_ntypes())
    csv = csv.reshape(1, ntypes * (ntypes + 1), 100)
    incr = 1.0 / 99
    counts = np.rint(csv / incr).astype(np.uint64)
    assert np.abs(counts * incr - csv).max() < 1e-9
    np.savez_compressed(os.path.join(HERE, "gofr_notebook.npz"), pos_unwrapped=raw, pos_wrapped=wrapped,
                        box_internal=box, types=types, csv=csv, counts=counts,
                        params=np.array([0.5, 3.8, 100, 1, 1, 100]))


def min_image_and_pbc(m):
    """tests/src/test_trajectory.cpp:21-59: all 56^2 d2_minImage of lammps.bin frame 0, and the
    wrapped first frames of lammps.bin / lammps2020.bin."""
    out = {}
    for tag, fname, gold in (("1", "lammps.bin", "pbc_1"), ("2", "lammps2020.bin", "pbc_2")):
        tr = m.Traj(REF + "/tests/data/" + fname)
        tr.setWrapPbc(False)
        tr.setAccessWindowSize(1)
        tr.setAccessStart(0)
        pos = tr.get_positions_copy()[0].copy()
        box = tr.get_box_copy()[0].copy()
        g = np.fromfile(REF + "/tests/cpp_regression_data/" + gold, dtype=np.float64).reshape(-1, 3)
        assert g.shape == pos.shape
        out["pos_" + tag] = pos
        out["box_" + tag] = box
        out["pbc_" + tag] = g
        if tag == "1":
            tw = m.Traj(REF + "/tests/data/" + fname)
            tw.setWrapPbc(True)
            tw.setAccessWindowSize(1)
            tw.setAccessStart(0)
            out["pos_wrapped_1"] = tw.get_positions_copy()[0].copy()
            mi = np.fromfile(REF + "/tests/cpp_regression_data/min_image", dtype=np.float64)
            n = pos.shape[0]
            out["min_image_1"] = mi.reshape(n, n, 4)
    np.savez_compressed(os.path.join(HERE, "min_image_pbc.npz"), **out)


def ref_counts(m, pos, box_lammps, types, fmt, wrap, rmin, rmax, nbin, tmax, skip, every, ntimesteps, primo,
               nthreads=3):
    vel = np.zeros_like(pos)
    tr = m.Trajectory(pos, vel, types, box_lammps, fmt, wrap, False)
    g = m.Gofrt(tr, rmin, rmax, nbin, tmax, nthreads, skip, every, False)
    g.reset(ntimesteps)
    g.calculate(primo)
    v = np.array(g, copy=True)
    sk = skip or 1
    incr = 1.0 / (ntimesteps // sk) if ntimesteps // sk > 0 else 1.0
    counts = np.rint(v / incr).astype(np.uint64)
    assert np.abs(counts * incr - v).max() < 1e-9 * max(1.0, v.max())
    return counts, v, tr.get_positions_copy(), tr.get_box_copy(), tr.get_type_ids()


def live_reference_cases(m):
    """What no reference golden pins, taken from the compiled reference itself."""
    cases = {}

    def add(name, seed, cells, ntypes, triclinic, nframes, npt, wrap, gofrt, tilt=(0.15, -0.10, 0.08),
            type_rule="parity", a=1.1):
        pos, box, types = synth.small_case(seed, cells, a, ntypes, triclinic, nframes, type_rule, tilt, npt)
        raw_types = (types * 3 + 2).astype(np.int32)  # non-dense raw ids exercise the compaction
        fmt = m.BoxFormat.LammpsTriclinic if triclinic else m.BoxFormat.LammpsOrtho
        rmin, rmax, nbin, tmax, skip, every, ntimesteps, primo = gofrt
        counts, v, rpos, rbox, rtypes = ref_counts(m, pos, box, raw_types, fmt, wrap, rmin, rmax, nbin, tmax, skip,
                                                   every, ntimesteps, primo)
        cases[name] = dict(pos_in=pos, box_lammps=box, raw_types=raw_types, wrap=wrap, pos_ref=rpos, box_internal=rbox,
                           type_ids=rtypes, counts=counts, vdata=v,
                           params=np.array([rmin, rmax, nbin, tmax, skip, every, ntimesteps, primo], dtype=np.float64))

    # triclinic, wrapped, two types, a few lags
    add("tri_wrap", 11, (6, 5, 4), 2, True, 14, False, True, (0.0, 2.6, 40, 5, 2, 1, 8, 1))
    # triclinic NPT (box changes per frame, lag pairs use the box of the origin frame)
    add("tri_npt", 12, (5, 5, 4), 3, True, 12, True, True, (0.3, 2.4, 25, 4, 3, 2, 7, 0))
    # orthorhombic, UNWRAPPED input drifting out of the box (many images per pair)
    add("ortho_unwrapped", 13, (5, 4, 4), 2, False, 10, False, False, (0.0, 2.0, 30, 3, 1, 1, 6, 2), a=0.9)
    # big tilt: |xy|+|xz| > lx/2, the x loop can need two images even for wrapped input
    add("tri_bigtilt", 14, (4, 4, 4), 1, True, 8, False, True, (0.0, 2.2, 32, 3, 1, 1, 5, 0),
        tilt=(0.45, -0.40, 0.35))
    # ragged: skip does not divide ntimesteps, every does not divide leff, nbin small
    add("ragged", 15, (4, 4, 3), 2, True, 16, False, True, (0.2, 2.0, 7, 6, 4, 4, 9, 1))
    flat = {}
    for name, d in cases.items():
        for k, v in d.items():
            flat[name + "/" + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "live_reference.npz"), **flat)


TILE_CASES = {
    # name: (seed, cells, ntypes, triclinic, nframes, (rmin, rmax, nbin, tmax, skip, every, ntimesteps, primo))
    # C3-shaped, reduced: 3 types, triclinic, 3072 atoms -> the TILE kernel's triclinic path, sparse binning
    "tri_tile3": (41, (16, 16, 12), 3, True, 5, (0.0, 4.0, 64, 2, 2, 1, 4, 0)),
    # C2-shaped, reduced: cubic, 39 % of the pairs in range -> the dense tile kernel (two-floor binning, long rows)
    "ortho_tile_dense": (42, (12, 12, 12), 2, False, 5, (0.0, 6.0, 96, 2, 2, 1, 4, 0)),
    # the same with rmin > 0 a multiple of dr (c0 = -10: guard bins below bin 0)
    "ortho_tile_rmin": (43, (12, 12, 12), 1, False, 5, (0.6, 6.6, 100, 2, 2, 1, 4, 0)),
}


def live_reference_tile_cases(m):
    """Systems of a few thousand atoms through the compiled reference, so that the tile kernel (more than 256 device
    slots) is pinned by the reference itself and not only by the oracle.  The positions are not stored: they are
    synth.small_case(seed, ...) again (numpy Generator streams are stable), with their sha256 kept to say so loudly if
    that ever changes."""
    import hashlib
    out = {}
    for name, (seed, cells, ntypes, tri, nframes, gofrt) in TILE_CASES.items():
        pos, box, types = synth.small_case(seed, cells, 1.1, ntypes, tri, nframes, "parity")
        raw_types = (types * 3 + 2).astype(np.int32)
        fmt = m.BoxFormat.LammpsTriclinic if tri else m.BoxFormat.LammpsOrtho
        rmin, rmax, nbin, tmax, skip, every, nts, primo = gofrt
        counts, v, rpos, rbox, rtypes = ref_counts(m, pos, box, raw_types, fmt, True, rmin, rmax, nbin, tmax, skip, every, nts,
                                                   primo)
        out[name + "/counts"] = counts
        out[name + "/vdata"] = v
        out[name + "/pos_in_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pos).tobytes()).digest(), dtype=np.uint8)
        out[name + "/pos_ref_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(rpos).tobytes()).digest(), dtype=np.uint8)
        out[name + "/box_internal"] = rbox
        out[name + "/type_ids"] = rtypes
        out[name + "/params"] = np.array(gofrt, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "live_reference_tile.npz"), **out)


PAIR_LOOP_CASES = {
    # name: (seed, cells, ntypes, triclinic, frame, neighbour spec [(max, cutoff^2, skin^2)], lmax, nbin, rminmax)
    "tri2": (81, (4, 4, 3), 2, True, 1, [(40, 4.0, 4.0), (40, 6.25, 6.25)], 4, 3, [(0.5, 3.0), (0.4, 2.6), (0.4, 2.6), (0.0, 2.0)]),
    "ortho3": (82, (5, 4, 3), 3, False, 2, [(30, 3.0, 3.0), (30, 3.0, 3.0), (30, 5.0, 5.0)], 6, 2, [(0.3, 2.4)] * 9),
    "tri1_l10": (83, (3, 3, 3), 1, True, 0, [(60, 9.0, 9.0)], 10, 2, [(0.0, 3.2)]),
}


def pair_loop_fixtures(m):
    """The other two pair loops over d2_minImage, from the compiled reference: Neighbours::update_neigh (unsorted and sorted
    lists, SANN counts; lib/src/neighbour.cpp) and SphericalBase::calc (lib/src/sphericalbase.cpp).  Inputs are
    synth.small_case(seed, ...) again (sha256 kept)."""
    import hashlib
    out = {}
    for name, (seed, cells, ntypes, tri, frame, spec, lmax, nbin, rminmax) in PAIR_LOOP_CASES.items():
        pos, box, types = synth.small_case(seed, cells, 1.1, ntypes, tri, 3, "parity")
        fmt = m.BoxFormat.LammpsTriclinic if tri else m.BoxFormat.LammpsOrtho
        tr = m.Trajectory(pos, np.zeros_like(pos), types, box, fmt, True, False)
        out[name + "/pos_in_sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pos).tobytes()).digest(), dtype=np.uint8)
        for sort in (False, True):
            c, idx, r, sann = m.neighbours(tr, spec, frame, sort)
            tag = "/sorted_" if sort else "/unsorted_"
            out[name + tag + "counts"] = c
            out[name + tag + "idx"] = idx
            out[name + tag + "r"] = r
            if sort:
                out[name + "/sann_n"] = sann
        res, cnt = m.sh_density(tr, lmax, nbin, rminmax, frame)
        out[name + "/sh"] = res
        out[name + "/sh_counter"] = cnt
    np.savez_compressed(os.path.join(HERE, "pair_loops.npz"), **out)


def cell_vectors_rotation(m):
    """Trajectory_numpy with BoxFormat.CellVectors and general (rotated) cells: the reference QR-rotates
    cell, positions and velocities into the LAMMPS frame (lib/include/triclinic.h:10-73 with Eigen's
    Householder QR, lib/src/trajectory_numpy.cpp:89-172).  Pins the host mirror's hand-written QR and,
    on the GPU, g(r,t) through that input format."""
    rng = np.random.default_rng(31)
    nfr, n = 7, 60
    cell = np.array([[7.0, 0.0, 0.0], [1.2, 6.5, 0.0], [0.9, -0.7, 6.0]]).T.copy()  # columns = cell vectors
    cells = []
    for f in range(nfr):
        if f in (0, 3, 5):  # runs of identical cells, a new random orientation per run
            q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            cur = q @ (cell * (1.0 + 0.01 * f))
        cells.append(cur)
    cells = np.ascontiguousarray(np.stack(cells))
    frac = rng.random(size=(nfr, n, 3))
    pos = np.ascontiguousarray(np.einsum("fij,fnj->fni", cells, frac))
    vel = rng.normal(size=pos.shape)
    types = (np.arange(n) % 2 * 4 + 1).astype(np.int32)
    out = dict(pos=pos, vel=vel, types=types, cells=cells)
    for wrap in (False, True):
        tr = m.Trajectory(pos, vel, types, cells, m.BoxFormat.CellVectors, wrap, True)
        tag = "wrap" if wrap else "nowrap"
        out["pos_" + tag] = tr.get_positions_copy()
        if not wrap:
            out["box_internal"] = tr.get_box_copy()
            out["rotation"] = tr.get_rotation_matrix()
            out["type_ids"] = tr.get_type_ids()
        else:
            g = m.Gofrt(tr, 0.0, 3.0, 30, 3, 2, 1, 1, False)
            g.reset(4)
            g.calculate(1)
            v = np.array(g, copy=True)
            out["vdata"] = v
            out["counts"] = np.rint(v * 4).astype(np.uint64)
            out["params"] = np.array([0.0, 3.0, 30, 3, 1, 1, 4, 1], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "cell_vectors_rotation.npz"), **out)


def qr_many(m):
    """200 general cell matrices (random scales, triangular, partly zero, negative, nearly degenerate first
    column) through the reference's Trajectory_numpy: internal box rows, rotation matrices and one rotated
    position per frame.  Pins the hand-written 3x3 Householder QR of include/analisi/triclinic.h bit for bit."""
    rng = np.random.default_rng(77)
    cells = []
    for k in range(200):
        M = rng.normal(size=(3, 3)) * rng.choice([0.5, 3.0, 20.0])
        if k % 10 == 1:
            M = np.tril(M)
        if k % 10 == 2:
            M = np.triu(M)
        if k % 10 == 3:
            M[1, 0] = 0
            M[2, 0] = 0
        if k % 10 == 4:
            M[2, 1] = 0
        if k % 10 == 5:
            M = -np.abs(M)
        if k % 10 == 6:
            M = np.diag(np.abs(rng.normal(size=3)) + 1) @ np.array([[1, 0.3, 0.2], [0, 1, 0.1], [0, 0, 1.0]])
        if k % 10 == 7:
            M[:, 0] = [1e-9, 2.0, 0]
        cells.append(M)
    cells = np.ascontiguousarray(np.stack(cells))
    pos = np.ascontiguousarray(rng.normal(size=(len(cells), 2, 3)))
    tr = m.Trajectory(pos, np.zeros_like(pos), np.zeros(2, dtype=np.int32), cells, m.BoxFormat.CellVectors, False, True)
    np.savez_compressed(os.path.join(HERE, "qr_many.npz"), cells=cells, pos=pos, box_internal=tr.get_box_copy(),
                        rotation=tr.get_rotation_matrix(), pos_rotated=tr.get_positions_copy())


def cli_golden_text():
    """The reference's CLI golden output for the g(r,t) branch (tests/test_cli.sh:33-34), copied as is:
    analisi -i tests/data/lammps2020.bin -g 100 -F 0.0 4.0 -S {1,10} -s 8  (20 blocks, mean and variance)."""
    import shutil
    # ... and the neighbour-count histogram of the next scope row (tests/test_cli.sh:35: --neighbour 10)
    # ... and the mean square displacement (tests/test_cli.sh:24-27: -Q; -Q -s 10 -S 50; -q -s 10 -S 50; -Q --mean-square-displacement-self)
    for name in ("pair_corr_no_t", "pair_corr_t", "neighbours", "MSD_normal_full", "MSD_normal", "MSD_cm", "MSD_cm_reference"):
        shutil.copyfile(os.path.join(REF, "tests/data/cli", name), os.path.join(HERE, "cli_" + name + ".txt"))


def main():
    m = ref_module()
    gofr_numpy(m)
    gofr_notebook(m)
    min_image_and_pbc(m)
    live_reference_cases(m)
    live_reference_tile_cases(m)
    pair_loop_fixtures(m)
    cell_vectors_rotation(m)
    qr_many(m)
    cli_golden_text()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz") or f.endswith(".txt"):
            print("%-24s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()
