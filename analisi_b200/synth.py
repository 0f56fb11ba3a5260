"""Synthetic liquid trajectories of the shapes BASELINE.json names.

Recipe (same for every config): jittered lattice at liquid density for frame 0 (lattice site +
N(0, 0.15 a)), then an independent Gaussian random walk N(0, 0.03 a) per atom per frame, so the
self part of g(r,t) spreads with the lag.  Positions are generated UNWRAPPED and are meant to be fed
with wrap=True; velocities are zero; everything is float64; ``numpy.random.default_rng(seed)``.

Boxes are returned in the LAMMPS row format the reference's python API takes
(``BoxFormat.LammpsOrtho``: xlo,xhi,ylo,yhi,zlo,zhi; ``LammpsTriclinic``: + xy,xz,yz).
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class Workload:
    """One named configuration: trajectory shape + the Gofrt arguments used on it."""
    name: str
    seed: int
    cells: tuple          # lattice sites along a, b, c
    a: object             # lattice spacing (float, or one per axis)
    ntypes: int
    type_rule: str        # "blocks" (by atom-index blocks) | "parity" (index % ntypes)
    triclinic: bool
    nframes: int
    rmin: float
    rmax: float
    nbin: int
    tmax: int             # Gofrt ctor "tmax" (number of lags, 0..tmax-1)
    skip: int
    every: int = 1
    nblocks: int = 1
    tilt: tuple = (0.15, -0.10, 0.08)   # xy/lx, xz/lx, yz/ly

    @property
    def natoms(self):
        return self.cells[0] * self.cells[1] * self.cells[2]

    @property
    def spacing(self):
        return tuple(self.a) if isinstance(self.a, (tuple, list)) else (self.a, self.a, self.a)


_A_LJ = (1.0 / 0.8442) ** (1.0 / 3.0)   # simple-cubic spacing at rho* = 0.8442

# C2..C5.  C1 is the bundled tests/data/lammps.bin (not synthetic).
WORKLOADS = {
    "C2": Workload("C2 4096-atom LJ liquid, cubic, 2000 frames, 200 bins, lags 0-100", 2002,
                   (16, 16, 16), 16.9276 / 16, 1, "blocks", False, 2000, 0.0, 8.0, 200, 101, 1),
    "C3": Workload("C3 12288-atom 3-type water-like, triclinic, 1000 frames, 300 bins", 2003,
                   (32, 24, 16), (49.71 / 32, 49.71 / 24, 49.71 / 16), 3, "blocks", True, 1000, 0.0, 12.0, 300, 10, 15),
    "C4": Workload("C4 100k-atom triclinic liquid, 1000 frames, 500 bins, lags 0-200", 2004,
                   (50, 50, 40), _A_LJ, 1, "blocks", True, 1000, 0.0, 10.0, 500, 201, 12),
    "C5": Workload("C5 1M-atom 2-type melt, cubic, 200 frames, 8 blocks (MediaBlocchi)", 2005,
                   (100, 100, 100), _A_LJ, 2, "parity", False, 200, 0.0, 8.0, 200, 1, 6, 1, 8),
}


# The default bench step (bench.py) for workloads whose full block is too long for one step: the same atoms,
# frames, cell and bins, a regular SUBSET of the block's (lag, origin) jobs.  C4: every 8th lag x every 96th
# origin = 26 x 8 = 208 jobs of 1e10 pair evaluations; int(768/96) = 8 keeps incr a power of two.  The counts of
# exactly this subset from the unmodified reference are committed (tests/golden/c4_subset_counts.*).
BENCH_SUBSET = {"C4": {"every": 8, "skip": 96, "ntimesteps": 768}}


def bench_subset(name):
    """(workload with the subset's skip / every, ntimesteps) of the default bench step of ``name``."""
    import dataclasses
    s = BENCH_SUBSET[name]
    w = WORKLOADS[name]
    return dataclasses.replace(w, skip=s["skip"], every=s["every"], name=w.name + " [subset: every %dth lag x every %dth origin]"
                               % (s["every"], s["skip"])), s["ntimesteps"]


def lattice_types(w: Workload):
    n = w.natoms
    idx = np.arange(n)
    if w.type_rule == "parity":
        return (idx % w.ntypes).astype(np.int32)
    per = -(-n // w.ntypes)
    return (idx // per).astype(np.int32)


def lammps_box_row(w: Workload):
    """One LAMMPS box row (6 or 9 doubles) for the workload's cell."""
    ax, ay, az = w.spacing
    lx, ly, lz = (w.cells[0] * ax, w.cells[1] * ay, w.cells[2] * az)
    if not w.triclinic:
        return np.array([0.0, lx, 0.0, ly, 0.0, lz])
    return np.array([0.0, lx, 0.0, ly, 0.0, lz, w.tilt[0] * lx, w.tilt[1] * lx, w.tilt[2] * ly])


def generate(w: Workload, nframes=None, first_frame=0, dtype=np.float64):
    """Return (pos[F,N,3] unwrapped, box_lammps[F,6|9], types[N] int32) for frames
    first_frame .. first_frame+nframes-1 of the workload's random walk.

    The walk of frame f depends only on (seed, f), so any window can be generated without the
    frames before it being held in memory: frame f = frame0 + sum_{g<=f} step_g, where the partial
    sums are accumulated chunk by chunk.
    """
    if nframes is None:
        nframes = w.nframes
    nx, ny, nz = w.cells
    n = w.natoms
    row = lammps_box_row(w)
    lx, ly, lz = row[1], row[3], row[5]
    # fractional lattice sites, x fastest
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    frac = np.stack([(ix.T.ravel() + 0.5) / nx, (iy.T.ravel() + 0.5) / ny, (iz.T.ravel() + 0.5) / nz], axis=1)
    if w.triclinic:
        xy, xz, yz = row[6], row[7], row[8]
        cell = np.array([[lx, 0.0, 0.0], [xy, ly, 0.0], [xz, yz, lz]])  # rows = a, b, c
    else:
        cell = np.diag([lx, ly, lz])
    site = frac @ cell
    rng0 = np.random.default_rng([w.seed, 0])
    amin = min(w.spacing)
    cur = site + rng0.normal(0.0, 0.15 * amin, size=(n, 3))
    pos = np.empty((nframes, n, 3), dtype=dtype)
    sigma = 0.03 * amin
    for f in range(first_frame + nframes):
        if f > 0:
            cur += np.random.default_rng([w.seed, f]).normal(0.0, sigma, size=(n, 3))
        if f >= first_frame:
            pos[f - first_frame] = cur
    box = np.tile(row, (nframes, 1))
    return pos, box, lattice_types(w)


def small_case(seed, natoms_cells=(6, 5, 4), a=1.1, ntypes=2, triclinic=True, nframes=12, type_rule="parity",
               tilt=(0.15, -0.10, 0.08), npt=False):
    """A small workload of the same recipe for parity tests; ``npt`` makes the box breathe per frame."""
    w = Workload("small", seed, tuple(natoms_cells), a, ntypes, type_rule, triclinic, nframes,
                 0.0, 1.0, 10, 1, 1, tilt=tuple(tilt))
    pos, box, types = generate(w)
    if npt:
        rng = np.random.default_rng([seed, 999])
        s = 1.0 + 0.02 * rng.standard_normal(nframes)
        box = box * s[:, None]
        pos = pos * s[:, None, None]
    return pos, box, types


def lammps_rows_to_internal(box_lammps):
    """[xlo,xhi,ylo,yhi,zlo,zhi(,xy,xz,yz)] -> internal [xlo,ylo,zlo,lx/2,ly/2,lz/2(,xy,xz,yz)]
    with the reference's arithmetic (basetrajectory.h:94-105: (hi-lo)/2)."""
    b = np.array(box_lammps, dtype=np.float64)
    out = b.copy()
    out[..., 0] = b[..., 0]
    out[..., 1] = b[..., 2]
    out[..., 2] = b[..., 4]
    out[..., 3] = (b[..., 1] - b[..., 0]) / 2
    out[..., 4] = (b[..., 3] - b[..., 2]) / 2
    out[..., 5] = (b[..., 5] - b[..., 4]) / 2
    return out


def write_lammps_binary(path, pos, box_lammps, raw_types, vel=None, ids=None, format2020=False, nchunk=1,
                        shuffle_seed=None, first_step=0, step_stride=1):
    """Write frames as a LAMMPS binary dump "id type xu yu zu vx vy vz" (8 doubles per atom).

    ``box_lammps`` rows are [xlo,xhi,ylo,yhi,zlo,zhi(,xy,xz,yz)]; 9 columns make the file triclinic.
    ``format2020`` writes the header flavour with the magic string / revision 2 / unit style / column
    names (the layout reference lib/include/lammps_struct.h:127-189 reads), otherwise the older one
    (:31-101).  ``nchunk`` splits the atoms of every frame over several chunks and ``shuffle_seed``
    permutes their order per frame (after frame 0), as a parallel LAMMPS run does; readers must
    scatter by atom id."""
    import struct
    pos = np.asarray(pos, dtype=np.float64)
    nfr, n, _ = pos.shape
    box = np.asarray(box_lammps, dtype=np.float64).reshape(nfr, -1)
    tri = box.shape[1] == 9
    vel = np.zeros_like(pos) if vel is None else np.asarray(vel, dtype=np.float64)
    ids = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    raw_types = np.asarray(raw_types)
    rng = np.random.default_rng(shuffle_seed) if shuffle_seed is not None else None
    with open(path, "wb") as f:
        for t in range(nfr):
            step = first_step + t * step_stride
            if format2020:
                magic = b"DUMPCUSTOM"
                f.write(struct.pack("<q", -len(magic)) + magic + struct.pack("<ii", 1, 2))
            f.write(struct.pack("<qq", step, n))
            f.write(struct.pack("<i6i", 1 if tri else 0, 0, 0, 0, 0, 0, 0))
            f.write(box[t, :6].tobytes())
            if tri:
                f.write(box[t, 6:9].tobytes())
            f.write(struct.pack("<i", 8))
            if format2020:
                units, cols = b"lj", b"id type xu yu zu vx vy vz"
                f.write(struct.pack("<i", len(units)) + units)
                f.write(struct.pack("<c", b"\x00"))
                f.write(struct.pack("<i", len(cols)) + cols)
            f.write(struct.pack("<i", nchunk))
            order = np.arange(n)
            if rng is not None and t > 0:
                order = rng.permutation(n)
            rows = np.empty((n, 8), dtype=np.float64)
            rows[:, 0] = ids[order]
            rows[:, 1] = raw_types[order]
            rows[:, 2:5] = pos[t, order]
            rows[:, 5:8] = vel[t, order]
            cuts = np.linspace(0, n, nchunk + 1).astype(int)
            for c in range(nchunk):
                part = rows[cuts[c]:cuts[c + 1]]
                f.write(struct.pack("<i", part.size))
                f.write(part.tobytes())
    return path


def frames(w: Workload, nframes=None):
    """Generator over (frame index, pos[N,3]) of the workload's random walk: the same frames as
    ``generate`` (same seeds), one at a time, so a trajectory larger than memory can be streamed to a file."""
    if nframes is None:
        nframes = w.nframes
    pos0, _, _ = generate(w, nframes=1)
    cur = pos0[0].copy()
    sigma = 0.03 * min(w.spacing)
    for f in range(nframes):
        if f > 0:
            cur += np.random.default_rng([w.seed, f]).normal(0.0, sigma, size=cur.shape)
        yield f, cur


def write_workload_lammps(path, w: Workload, nframes=None):
    """Stream the workload to a LAMMPS binary dump (pre-2020 header, one chunk per frame, ids 1..N, raw
    types 1..ntypes); returns the number of bytes written.  C5: 200 frames x 1M atoms x 64 B = 12.8 GB."""
    import struct
    row = lammps_box_row(w)
    tri = row.size == 9
    types = lattice_types(w)
    n = w.natoms
    rows = np.zeros((n, 8), dtype=np.float64)
    rows[:, 0] = np.arange(1, n + 1)
    rows[:, 1] = types + 1
    total = 0
    with open(path, "wb") as f:
        for t, pos in frames(w, nframes):
            head = struct.pack("<qq", t, n) + struct.pack("<i6i", 1 if tri else 0, 0, 0, 0, 0, 0, 0) + row[:6].tobytes()
            if tri:
                head += row[6:9].tobytes()
            head += struct.pack("<ii", 8, 1) + struct.pack("<i", n * 8)
            rows[:, 2:5] = pos
            f.write(head)
            f.write(rows.tobytes())
            total += len(head) + rows.nbytes
    return total
