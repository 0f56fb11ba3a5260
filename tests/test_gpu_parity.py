"""GPU: the CUDA path (through the C ABI of libagofrt.so) against the oracle and the committed
golden fixtures.  Integer bin counts must be bit-exact; count*incr must equal the reference's
float sums within 1e-12 relative (stated in each test)."""
import numpy as np
import pytest

import oracle
from analisi_b200 import cabi, synth
from conftest import LIVE_CASES, PAIR_LOOP_CASES, TILE_CASES, live_case, load_golden, pair_loop_case, tile_case

pytestmark = pytest.mark.gpu

REL_TOL = 1e-12   # north_star: normalised g(r,t) within 1e-12 relative


def gpu_counts(ctx, pos, box_internal, ids, ntypes, rmin, rmax, nbin, tmax, nts, primo=0, skip=1, every=1,
               options=0, edges=False, first_frame=0):
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], box_internal.shape[1], ids, ntypes, max(1, pos.shape[0]))
    tr.upload(first_frame, np.ascontiguousarray(pos), box_internal)
    plan = cabi.Plan(tr, rmin, rmax, nbin)
    leff = cabi.gofrt_leff(nts, tmax)
    out = plan.block(primo, nts, leff, skip, every, options=options, edges=edges)
    plan.close()
    tr.close()
    return out


# systems of up to 256 device slots take the small-system kernel unless told otherwise (up to 512 with OPT_SMALL):
# the small fixtures run through both kernels
BOTH_KERNELS = pytest.mark.parametrize("no_small", [0, cabi.OPT_NO_SMALL], ids=["small-kernel", "tile-kernel"])


def ran_small(st):
    return bool(st["kernel_modes"] & cabi.MODE_BIT_SMALL)


@BOTH_KERNELS
def test_gofr_numpy_golden(ctx, no_small):
    """reference tests/test_gofrt.py: ortho, UNWRAPPED input (many images per pair), 3 types, 10 lags."""
    z = load_golden("gofr_numpy.npz")
    rmin, rmax, nbin, tmax, skip, nts = z["params"]
    c, st = gpu_counts(ctx, z["pos"], z["box_internal"], z["types"], 3, rmin, rmax, int(nbin), int(tmax), int(nts),
                       skip=int(skip), options=no_small)
    assert np.array_equal(c, z["counts"])
    assert ran_small(st) == (no_small == 0)
    incr = cabi.gofrt_incr(int(nts), int(skip))
    v = c * incr
    nz = z["csv"] != 0
    assert np.abs(v[nz] - z["csv"][nz]).max() <= 1e-11  # the CSV itself carries the reference's 4-thread float sums
    assert (v[~nz] == 0).all()
    assert st["pair_evals_total"] == 10 * 70 * 56 * 56


@BOTH_KERNELS
def test_gofr_notebook_golden(ctx, no_small):
    """reference tests/test_notebook.py: mmap trajectory, wrap on; the wrap itself runs on the GPU."""
    z = load_golden("gofr_notebook.npz")
    pos = z["pos_unwrapped"].copy()
    ctx.pbc_wrap(pos, z["box_internal"])
    assert np.array_equal(pos, z["pos_wrapped"])
    rmin, rmax, nbin, tmax, skip, nts = z["params"]
    nt = int(z["types"].max()) + 1
    c, st = gpu_counts(ctx, pos, z["box_internal"], z["types"], nt, rmin, rmax, int(nbin), int(tmax), int(nts),
                       skip=int(skip), options=no_small)
    assert np.array_equal(c, z["counts"])
    assert st["jobs_fast"] == st["jobs"]  # wrapped orthorhombic input: single-pass minimum image proven
    assert st["kernel_modes"] & 0xff in (1 << 3, 1 << 4, 1 << 5)   # ... binned by a safe-zone kernel
    assert ran_small(st) == (no_small == 0)


def test_min_image_and_pbc_golden(ctx):
    """reference tests/src/test_trajectory.cpp:21-59 (tolerance 1e-10 there; bit-exact vs the oracle here)."""
    z = load_golden("min_image_pbc.npz")
    for tag in ("1", "2"):
        p = z["pos_" + tag][None].copy()
        ctx.pbc_wrap(p, z["box_" + tag][None])
        g = z["pbc_" + tag]
        assert np.abs(p[0] - g).max() <= 1e-10 * max(1.0, np.abs(g).max())
        assert np.array_equal(p, oracle.pbc_wrap(z["pos_" + tag][None], z["box_" + tag][None]))
    pw = z["pos_wrapped_1"]
    n = pw.shape[0]
    tr = cabi.DeviceTrajectory(ctx, n, 6, np.zeros(n, np.int32), 1, 1)
    tr.upload(0, pw[None].copy(), z["box_1"][None])
    d = tr.d2_all(0, 0)
    assert np.array_equal(tr.download_frame(0), pw)
    tr.close()
    g = z["min_image_1"]
    assert np.abs(d - g).max() <= 1e-10 * max(1.0, np.abs(g).max())
    assert np.array_equal(d, oracle.d2_all(pw, pw, z["box_1"]))


@pytest.mark.parametrize("name", LIVE_CASES)
@pytest.mark.parametrize("options", [0, cabi.OPT_FORCE_GENERAL, cabi.OPT_AGGREGATE, cabi.OPT_NO_SAFE,
                                     cabi.OPT_DENSE, cabi.OPT_SPARSE, cabi.OPT_NO_UBOX, cabi.OPT_NO_UBOX | cabi.OPT_NO_SAFE])
@BOTH_KERNELS
def test_live_reference_cases(ctx, name, options, no_small):
    """Fixtures computed by the compiled reference: triclinic, NPT, unwrapped, big tilt, ragged loops."""
    d = live_case(name)
    rmin, rmax, nbin, tmax, skip, every, nts, primo = d["params"]
    ids, nt = d["type_ids"], int(d["type_ids"].max()) + 1
    pos = d["pos_in"].copy()
    if bool(d["wrap"]):
        ctx.pbc_wrap(pos, d["box_internal"])
    assert np.array_equal(pos, d["pos_ref"])
    c, st = gpu_counts(ctx, pos, d["box_internal"], ids, nt, rmin, rmax, int(nbin), int(tmax), int(nts),
                       primo=int(primo), skip=int(skip), every=int(every), options=options | no_small)
    assert np.array_equal(c, d["counts"])
    assert ran_small(st) == (no_small == 0)   # every live fixture has at most 120 atoms
    incr = cabi.gofrt_incr(int(nts), int(skip))
    v, ref = c * incr, d["vdata"]
    nz = ref != 0
    assert (np.abs(v[nz] - ref[nz]) <= REL_TOL * np.abs(ref[nz])).all()
    assert (v[~nz] == 0).all()
    if options == cabi.OPT_FORCE_GENERAL:
        assert st["jobs_fast"] == 0
    if options == cabi.OPT_NO_SAFE:
        assert st["kernel_modes"] & (1 << 3) == 0


@pytest.mark.parametrize("name", sorted(TILE_CASES))
@pytest.mark.parametrize("options", [0, cabi.OPT_FORCE_GENERAL, cabi.OPT_NO_SAFE, cabi.OPT_SAFE2, cabi.OPT_SAFE2 | cabi.OPT_SKEW,
                                     cabi.OPT_SAFE2 | cabi.OPT_DENSE | cabi.OPT_NO_UBOX, cabi.OPT_DENSE, cabi.OPT_SPARSE,
                                     cabi.OPT_NO_UBOX, cabi.OPT_AGGREGATE])
def test_live_reference_tile_cases(ctx, name, options):
    """Fixtures of a few thousand atoms computed by the compiled reference (C3-shaped: 3 types, triclinic; C2-shaped:
    cubic, dense): the TILE kernel in every binning mode, bit-exact against the reference's own counts."""
    import hashlib
    d = tile_case(name)
    rmin, rmax, nbin, tmax, skip, every, nts, primo = d["params"]
    ids, nt = d["type_ids"], int(d["type_ids"].max()) + 1
    pos = d["pos_in"].copy()
    ctx.pbc_wrap(pos, d["box_internal"])
    sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(pos).tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, d["pos_ref_sha256"])   # the GPU wrap gives the reference's wrapped positions
    c, st = gpu_counts(ctx, pos, d["box_internal"], ids, nt, rmin, rmax, int(nbin), int(tmax), int(nts),
                       primo=int(primo), skip=int(skip), every=int(every), options=options)
    assert np.array_equal(c, d["counts"])
    assert not ran_small(st)
    if options & cabi.OPT_SAFE2 and (name.startswith("ortho_tile") or options & cabi.OPT_DENSE):
        assert st["kernel_modes"] & (1 << 5), "the dense two-floor kernel (MODE_SAFE2) was asked for here"
    if not options & cabi.OPT_SAFE2:
        assert st["kernel_modes"] & (1 << 5) == 0
    incr = cabi.gofrt_incr(int(nts), int(skip))
    v, ref = c * incr, d["vdata"]
    nz = ref != 0
    assert (np.abs(v[nz] - ref[nz]) <= REL_TOL * np.abs(ref[nz])).all()


@pytest.mark.parametrize("name", LIVE_CASES)
@BOTH_KERNELS
def test_edge_pairs_match_oracle(ctx, name, no_small):
    """Pairs whose d2 is a bin threshold or the double just below one are reported separately."""
    d = live_case(name)
    rmin, rmax, nbin, tmax, skip, every, nts, primo = d["params"]
    ids, nt = d["type_ids"], int(d["type_ids"].max()) + 1
    pos = d["pos_ref"]
    c, st, e = gpu_counts(ctx, pos, d["box_internal"], ids, nt, rmin, rmax, int(nbin), int(tmax), int(nts),
                          primo=int(primo), skip=int(skip), every=int(every), edges=True, options=no_small)
    co, eo = oracle.counts(pos, d["box_internal"], ids, rmin, rmax, int(nbin), int(tmax), int(nts), primo=int(primo),
                           skip=int(skip), every=int(every), ntypes=nt, return_edges=True)
    assert np.array_equal(c, co)
    assert e == eo


def test_thresholds_reproduce_reference_binning(ctx):
    """The table the kernel bins with is exact: at and just below every threshold the reference
    expression (oracle's gofrt_oracle_bin) changes value exactly there."""
    import ctypes as C
    lib = oracle.oracle._load()
    lib.gofrt_oracle_bin.argtypes = [C.c_double, C.c_double, C.c_double, C.c_uint]
    lib.gofrt_oracle_bin.restype = C.c_int
    tr = cabi.DeviceTrajectory(ctx, 1, 6, np.zeros(1, np.int32), 1, 1)
    for rmin, rmax, nbin in ((0.0, 3.8, 200), (0.7, 3.5, 200), (0.5, 3.8, 100), (0.0, 10.0, 500), (0.3, 2.4, 25)):
        plan = cabi.Plan(tr, rmin, rmax, nbin)
        T = plan.thresholds()
        dr = (rmax - rmin) / nbin
        for k in range(nbin + 1):
            assert lib.gofrt_oracle_bin(T[k], rmin, dr, nbin) >= k
            if T[k] > 0:
                assert lib.gofrt_oracle_bin(np.nextafter(T[k], -np.inf), rmin, dr, nbin) < k
        plan.close()
    tr.close()


@pytest.mark.parametrize("triclinic", [False, True])
def test_multi_tile_random(ctx, triclinic):
    """N large enough for several i tiles, j tiles and j chunks; 2 types of unequal size (padding)."""
    pos, box, types = synth.small_case(31 + triclinic, (12, 11, 10), 1.07, 2, triclinic, 5, "blocks")
    types = (np.arange(pos.shape[1]) % 5 == 0).astype(np.int32)
    bi = synth.lammps_rows_to_internal(box)
    ctx.pbc_wrap(pos, bi)
    args = (0.0, 4.5, 90, 3, 3)
    c, st = gpu_counts(ctx, pos, bi, types, 2, *args, skip=2)
    co = oracle.counts(pos, bi, types, *args, skip=2, ntypes=2)
    assert np.array_equal(c, co)
    assert st["jobs_fast"] == st["jobs"]
    c2, st2 = gpu_counts(ctx, pos, bi, types, 2, *args, skip=2, options=cabi.OPT_FORCE_GENERAL | cabi.OPT_AGGREGATE)
    assert np.array_equal(c2, co)
    c3, st3 = gpu_counts(ctx, pos, bi, types, 2, *args, skip=2, options=cabi.OPT_NO_SAFE)
    assert np.array_equal(c3, co) and st3["kernel_modes"] == 1   # 1320 atoms: never the small-system kernel
    for opt, bit in ((cabi.OPT_DENSE, 4), (cabi.OPT_DENSE | cabi.OPT_SAFE2, 5), (cabi.OPT_SPARSE, 3)):
        c4, st4 = gpu_counts(ctx, pos, bi, types, 2, *args, skip=2, options=opt)
        assert np.array_equal(c4, co) and st4["kernel_modes"] == 1 << bit
    # every ordered pair with d2 in range lands somewhere: lag 0, self slot, bin 0 holds exactly N per origin
    n0, n1 = int((types == 0).sum()), int((types == 1).sum())
    assert c[0, 3 + 2, 0] == 2 * n0 and c[0, 3 + 1, 0] == 2 * n1


def test_window_offset_and_errors(ctx):
    """A window that does not start at frame 0 (Trajectory::set_access_at), and the error codes."""
    pos, box, types = synth.small_case(5, (4, 4, 4), 1.1, 2, False, 12)
    bi = synth.lammps_rows_to_internal(box)
    ctx.pbc_wrap(pos, bi)
    ref = oracle.counts(pos, bi, types, 0.0, 2.0, 20, 3, 4, primo=5, skip=1, ntypes=2)
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], 6, types, 2, 8)
    tr.upload(4, np.ascontiguousarray(pos[4:12]), bi[4:12])
    plan = cabi.Plan(tr, 0.0, 2.0, 20)
    c, _ = plan.block(5, 4, 3)
    assert np.array_equal(c, ref)
    with pytest.raises(cabi.AgofrtError) as e:
        plan.block(3, 4, 3)
    assert e.value.code == cabi.ERR_WINDOW
    with pytest.raises(cabi.AgofrtError) as e:
        plan.block(8, 4, 3)
    assert e.value.code == cabi.ERR_WINDOW
    bad = pos[4:12].copy()
    bad[2, 7, 1] = np.inf
    tr.upload(4, bad, bi[4:12])
    with pytest.raises(cabi.AgofrtError) as e:
        plan.block(5, 4, 3)
    assert e.value.code == cabi.ERR_NONFINITE
    # NaN coordinates are simply never in range (the reference skips them too)
    nanpos = pos[4:12].copy()
    nanpos[:, 3, :] = np.nan
    tr.upload(4, nanpos, bi[4:12])
    c, _ = plan.block(5, 4, 3)
    refn = oracle.counts(np.concatenate([pos[:4], nanpos]), bi, types, 0.0, 2.0, 20, 3, 4, primo=5, skip=1, ntypes=2)
    assert np.array_equal(c, refn)
    # empty block
    c, st = plan.block(5, 0, 0)
    assert c.shape == (0, 6, 20) and st["jobs"] == 0
    plan.close()
    tr.close()


def test_shards_sum_to_whole(ctx):
    """Work-unit shards (what each of `world` GPUs computes before the all-reduce) add up exactly."""
    pos, box, types = synth.small_case(9, (9, 8, 8), 1.05, 1, True, 6)
    bi = synth.lammps_rows_to_internal(box)
    ctx.pbc_wrap(pos, bi)
    whole, st = gpu_counts(ctx, pos, bi, types, 1, 0.0, 3.0, 50, 3, 3)
    total = np.zeros_like(whole)
    world = 3
    for r in range(world):
        c2 = cabi.Context()
        c2.set_shard(r, world)
        part, st2 = gpu_counts(c2, pos, bi, types, 1, 0.0, 3.0, 50, 3, 3)
        total += part
        assert st2["world"] == world
        c2.close()
    assert np.array_equal(total, whole)
    assert np.array_equal(whole, oracle.counts(pos, bi, types, 0.0, 3.0, 50, 3, 3, ntypes=1))


def test_size_independent_properties_large(ctx):
    """At a size the oracle cannot finish: total of the lag-0 'distinct' rows + self rows equals the
    number of ordered pairs within rmax counted independently per shard, and the lag-0 histogram is
    symmetric under swapping the two frames' roles (d2(i,j) == d2(j,i))."""
    w = synth.WORKLOADS["C2"]
    pos, box, types = synth.generate(w, nframes=3)
    bi = synth.lammps_rows_to_internal(box)
    ctx.pbc_wrap(pos, bi)
    c, st = gpu_counts(ctx, pos, bi, types, 1, w.rmin, w.rmax, w.nbin, 2, 2)
    n = w.natoms
    assert st["pair_evals_total"] == 2 * 2 * n * n
    assert c[0, 1, 0] == 2 * n               # self pairs at lag 0: d2 == 0 -> bin 0, once per origin
    assert c[0, 1, 1:].sum() == 0
    assert c[1, 1].sum() == 2 * n            # every atom's own displacement over one frame is < rmax
    # same frames, general kernel and aggregated atomics: identical integers
    c2, _ = gpu_counts(ctx, pos, bi, types, 1, w.rmin, w.rmax, w.nbin, 2, 2,
                       options=cabi.OPT_FORCE_GENERAL | cabi.OPT_AGGREGATE)
    assert np.array_equal(c, c2)
    # lag 0 distinct counts are even: (i,j) and (j,i) fall in the same bin
    assert (c[0, 0] % 2 == 0).all()


@pytest.mark.parametrize("rmin,rmax,nbin,expect_safe", [
    (1.5, 2.5, 500, True),      # rmin/dr = 750 guard bins below bin 0 in every shared-memory row
    (-0.3, 2.0, 40, False),     # negative rmin: pairs below rmin^2 are skipped by the reference but have a bin >= 0;
                                # the device validation of the plan sees it and gives the safe-zone mode up
    (2.0, 2.2, 1000, False),    # rmin/dr = 10000 > the guard budget: the plan falls back to the threshold kernel
    (0.0, 2.6, 3, False),       # 9 counters in all: the warp-aggregated threshold kernel is chosen instead
    (0.0, 2.6, 30, True),       # the plain case
])
@BOTH_KERNELS
def test_guarded_rows_and_fallback(ctx, rmin, rmax, nbin, expect_safe, no_small):
    """The unconditional safe-zone binning writes into guard bins for everything outside the histogram; the
    counts must not depend on how many guard bins a row has, nor on the kernel that was chosen."""
    pos, box, types = synth.small_case(71, (7, 6, 6), 1.07, 2, True, 6)
    bi = synth.lammps_rows_to_internal(box)
    ctx.pbc_wrap(pos, bi)
    ref, eref = oracle.counts(pos, bi, types, rmin, rmax, nbin, 3, 3, ntypes=2, return_edges=True)
    c, st = gpu_counts(ctx, pos, bi, types, 2, rmin, rmax, nbin, 3, 3, options=no_small)
    assert ran_small(st) == (no_small == 0)   # 252 atoms: four warps per job
    assert np.array_equal(c, ref)
    safe_ran = bool(st["kernel_modes"] & ((1 << 3) | (1 << 4) | (1 << 5)))
    assert safe_ran == expect_safe
    c2, st2 = gpu_counts(ctx, pos, bi, types, 2, rmin, rmax, nbin, 3, 3, options=cabi.OPT_NO_SAFE | no_small)
    assert np.array_equal(c2, ref) and not (st2["kernel_modes"] & ((1 << 3) | (1 << 4) | (1 << 5)))
    c3, st3, e3 = gpu_counts(ctx, pos, bi, types, 2, rmin, rmax, nbin, 3, 3, edges=True, options=no_small)
    assert np.array_equal(c3, ref) and e3 == eref
    for opt in (cabi.OPT_DENSE, cabi.OPT_SPARSE):
        c4, _ = gpu_counts(ctx, pos, bi, types, 2, rmin, rmax, nbin, 3, 3, options=opt | no_small)
        assert np.array_equal(c4, ref)


@BOTH_KERNELS
def test_ghost_slots_and_nan_atoms(ctx, no_small):
    """Type groups are padded to 8 slots with NaN ghosts; NaN coordinates of real atoms are never in range
    (as in the reference: every comparison with NaN is false).  Both go through the min.f32 clamp of the
    safe-zone guess or the exact path, never into a counted bin."""
    pos, box, types = synth.small_case(72, (5, 5, 3), 1.1, 3, False, 5)   # 75 atoms, 25 per type: 7 ghosts each
    bi = synth.lammps_rows_to_internal(box)
    ctx.pbc_wrap(pos, bi)
    ref = oracle.counts(pos, bi, types, 0.0, 2.5, 50, 2, 3, ntypes=3)
    c, st = gpu_counts(ctx, pos, bi, types, 3, 0.0, 2.5, 50, 2, 3, options=no_small)
    assert np.array_equal(c, ref)
    bad = pos.copy()
    bad[1, 7] = np.nan
    bad[2, 30, 1] = np.nan
    refn = oracle.counts(bad, bi, types, 0.0, 2.5, 50, 2, 3, ntypes=3)
    cn, stn = gpu_counts(ctx, bad, bi, types, 3, 0.0, 2.5, 50, 2, 3, options=no_small)
    assert np.array_equal(cn, refn)
    assert refn.sum() < ref.sum()


def test_independent_count_at_100k_atoms(ctx):
    """At C4's atom count (100 000, the oracle's N^2 loop would take minutes): the cumulative histogram of one
    frame against an INDEPENDENT algorithm -- scipy's periodic k-d tree counting pairs within each bin edge.
    Different arithmetic (per-dimension min(|d|, L-|d|), tree pruning), same physics: the counts may only
    differ by pairs sitting on a bin edge to the last bits."""
    import dataclasses
    from scipy.spatial import cKDTree
    w = dataclasses.replace(synth.WORKLOADS["C4"], triclinic=False)
    pos, box, types = synth.generate(w, nframes=1)
    L = np.array([box[0, 1], box[0, 3], box[0, 5]])
    p = np.ascontiguousarray(np.mod(pos, L))
    bi = synth.lammps_rows_to_internal(box)
    rmax, nbin = 5.0, 50
    c, st = gpu_counts(ctx, p, bi, types, 1, 0.0, rmax, nbin, 1, 1)
    assert st["pair_evals_total"] == w.natoms ** 2 and st["jobs_fast"] == 1
    assert c[0, 1, 0] == w.natoms and c[0, 1, 1:].sum() == 0       # self part: every atom at distance 0 from itself
    assert (c[0, 0] % 2 == 0).all()                               # (i,j) and (j,i) fall in the same bin
    tree = cKDTree(p[0], boxsize=L)
    edges = np.linspace(0.0, rmax, nbin + 1)[1:]
    cum_tree = tree.count_neighbors(tree, edges, cumulative=True)  # ordered pairs with d <= edge, i == j included
    cum_gpu = np.cumsum(c[0, 0] + c[0, 1]).astype(np.int64)
    # the reference rounds the double quotient (d - rmin)/dr to FLOAT before flooring (gofrt.cpp:115), so a pair within
    # 2^-24 (relative) below an edge belongs to the bin above it: a few pairs per edge out of millions, one-sided
    diff = cum_tree - cum_gpu
    assert diff.min() >= 0 and diff.max() <= 40, (diff.min(), diff.max())
    assert cum_gpu[-1] > 4e7


def test_c4_size_bit_exact_vs_oracle(ctx):
    """The headline target at its real atom count: 100 000 atoms, triclinic cell, 500 bins, one (lag 1, origin) job =
    1e10 pair evaluations -- bin counts AND the edge-pair count bit-exact against the oracle (about 10 s of the
    box's host threads), with the default kernel, the general minimum-image kernel and the edge-counting one."""
    w = synth.WORKLOADS["C4"]
    pos, box, types = synth.generate(w, nframes=2)
    bi = synth.lammps_rows_to_internal(box)
    pos = np.ascontiguousarray(pos)
    ctx.pbc_wrap(pos, bi)
    assert np.array_equal(pos, oracle.pbc_wrap(synth.generate(w, nframes=2)[0], bi))   # the wrap itself, 6e5 coordinates
    args = (w.rmin, w.rmax, w.nbin, 2, 2)   # reset(2), two lags, skip 2: the one origin 0 at lags 0 and 1
    ref, eref = oracle.counts(pos, bi, types, *args, skip=2, ntypes=1, total_frames=w.nframes, return_edges=True)
    assert ref.shape == (2, 2, 500) and ref[1].sum() > 1e8
    c, st = gpu_counts(ctx, pos, bi, types, 1, *args, skip=2)
    assert np.array_equal(c, ref)
    assert st["pair_evals_total"] == 2 * 10 ** 10 and st["jobs_fast"] == 2
    c2, st2, e2 = gpu_counts(ctx, pos, bi, types, 1, *args, skip=2, edges=True)
    assert np.array_equal(c2, ref) and e2 == eref
    c3, st3 = gpu_counts(ctx, pos, bi, types, 1, *args, skip=2, options=cabi.OPT_FORCE_GENERAL | cabi.OPT_NO_SAFE)
    assert np.array_equal(c3, ref) and st3["jobs_fast"] == 0


def test_upload_wrap_retarget_and_concurrent_upload(ctx):
    """Double-buffered windows through the C ABI: agofrt_traj_upload_wrap (wrap on the device, wrapped frames handed
    back) == oracle wrap + plain upload; agofrt_plan_retarget moves a plan between two windows; and the one
    concurrency the header allows -- a second host thread uploading window B while agofrt_block runs on A."""
    import threading
    pos, box, types = synth.small_case(81, (12, 10, 10), 1.06, 2, True, 12)
    bi = synth.lammps_rows_to_internal(box)
    raw_a, raw_b = np.ascontiguousarray(pos[:6]), np.ascontiguousarray(pos[6:])
    wa, wb = oracle.pbc_wrap(raw_a, bi[:6]), oracle.pbc_wrap(raw_b, bi[6:])
    A = cabi.DeviceTrajectory(ctx, pos.shape[1], 9, types, 2, 6)
    B = cabi.DeviceTrajectory(ctx, pos.shape[1], 9, types, 2, 6)
    buf_a = raw_a.copy()
    A.upload_wrap(0, buf_a, bi[:6])
    assert np.array_equal(buf_a, wa)
    assert np.array_equal(A.download_frame(3), wa[3])
    plan = cabi.Plan(A, 0.0, 3.0, 60)
    ref_a = oracle.counts(wa, bi[:6], types, 0.0, 3.0, 60, 3, 4, ntypes=2)
    ref_b = oracle.counts(wb, bi[6:], types, 0.0, 3.0, 60, 3, 4, primo=6, ntypes=2, first_frame=6, total_frames=12)
    out = {}
    buf_b = raw_b.copy()

    def compute():
        for _ in range(20):   # keep the device busy on A while B is uploaded
            out["a"] = plan.block(0, 4, 3)[0]

    def upload():
        for _ in range(5):
            buf_b[...] = raw_b
            B.upload_wrap(6, buf_b, bi[6:])

    ta, tb = threading.Thread(target=compute), threading.Thread(target=upload)
    ta.start()
    tb.start()
    ta.join()
    tb.join()
    assert np.array_equal(out["a"], ref_a)
    assert np.array_equal(buf_b, wb)
    plan.retarget(B)
    cb, st = plan.block(6, 4, 3)
    assert np.array_equal(cb, ref_b)
    plan.retarget(A)
    assert np.array_equal(plan.block(0, 4, 3)[0], ref_a)
    other = cabi.DeviceTrajectory(ctx, pos.shape[1], 9, np.zeros(pos.shape[1], dtype=np.int32), 1, 6)
    with pytest.raises(cabi.AgofrtError):
        plan.retarget(other)   # another number of types
    plan.close()
    for t in (A, B, other):
        t.close()


@pytest.mark.parametrize("triclinic,ntypes,wrap", [(False, 1, True), (True, 3, True), (True, 2, False)])
def test_neighbour_histogram_vs_oracle(ctx, triclinic, ntypes, wrap):
    """next scope row: IstogrammaAtomiRaggio::calculate through agofrt_neighbour_hist --
    orthorhombic / triclinic, several types (ghost slots), wrapped (single-pass kernel) and unwrapped input (general
    minimum image), skip, and accumulation over calls, bit-exact against the oracle."""
    pos, box, types = synth.small_case(91, (7, 6, 5), 1.08, ntypes, triclinic, 9)
    bi = synth.lammps_rows_to_internal(box)
    pos = np.ascontiguousarray(pos)
    if wrap:
        ctx.pbc_wrap(pos, bi)
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], bi.shape[1], types, ntypes, 9)
    tr.upload(0, pos, bi)
    r = 2.3
    h, st = tr.neighbour_hist(r, 1, 7, 3)            # frames 1, 4, 7
    ref = oracle.neighbour_hist(pos, bi, types, r, 1, 7, 3, ntypes=ntypes)
    assert np.array_equal(h, ref)
    assert h.sum() == 3 * pos.shape[1] * ntypes and st["jobs"] == 3
    assert st["jobs_fast"] == (3 if wrap else 0) or not wrap
    h, _ = tr.neighbour_hist(r, 0, 2, 1, hist=h)      # accumulates, like the reference's maps
    ref = oracle.neighbour_hist(pos, bi, types, r, 0, 2, 1, ntypes=ntypes, hist=ref)
    assert np.array_equal(h, ref)
    with pytest.raises(cabi.AgofrtError):
        tr.neighbour_hist(r, 5, 9, 1)                 # runs past the uploaded window
    tr.close()


@pytest.mark.parametrize("cm_msd,cm_self,lmax,skip", [(False, False, 0, 1), (True, False, 5, 3), (True, True, 4, 2), (False, True, 6, 1)])
def test_msd_vs_oracle(ctx, cm_msd, cm_self, lmax, skip):
    """scope row 8f-3: MSD<T>::calculate through agofrt_msd against the oracle -- the atom rows within 1e-12 relative
    (a sum / count on the device, a running mean in the reference), the centre-of-mass rows bit for bit."""
    pos, box, types = synth.small_case(95, (9, 8, 7), 1.1, 3, False, 32)   # 504 atoms in 3 types: several tiles per type
    bi = synth.lammps_rows_to_internal(box)
    pos = np.ascontiguousarray(pos)            # unwrapped, as the MSD wants them
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], 6, types, 3, pos.shape[0])
    tr.upload(0, pos, bi)
    cm = oracle.cm_positions(pos, types, 3)
    tr.set_cm(0, cm)
    nts, primo = 14, 3
    v, st = tr.msd(primo, nts, lmax, skip, cm_msd, cm_self)
    ref = oracle.msd(pos, types, nts, lmax, primo=primo, skip=skip, cm_msd=cm_msd, cm_self=cm_self, ntypes=3, cm=cm)
    assert v.shape == ref.shape
    assert np.array_equal(v[0, 0], np.zeros(3)) or cm_self is False or np.abs(v[0, 0]).max() < 1e-20
    scale = np.maximum(np.abs(ref[:, 0]), 1e-300)
    assert (np.abs(v[:, 0] - ref[:, 0]) <= 1e-12 * scale).all()
    if cm_msd:
        assert np.array_equal(v[:, 1], ref[:, 1])
    assert st["jobs"] == v.shape[0] * ((nts + skip - 1) // skip)
    with pytest.raises(cabi.AgofrtError):
        tr.msd(20, 14, 0, 1)   # needs frames beyond the window
    tr.close()


@pytest.mark.parametrize("cm_msd,nts,lmax,primo", [(False, 37, 21, 2), (True, 16, 16, 0), (False, 5, 33, 1), (True, 40, 0, 0)])
def test_msd_register_ring_vs_oracle(ctx, cm_msd, nts, lmax, primo):
    """MSD with every frame an origin (skip = 1, no centre of mass subtracted): the kernel that keeps the later frames
    of an atom in a register ring.  Origin counts that are not multiples of the ring, a ragged last chunk of lags,
    fewer origins than the ring is long."""
    pos, box, types = synth.small_case(96, (7, 6, 5), 1.1, 2, True, 82)   # 210 atoms, 82 frames
    pos = np.ascontiguousarray(pos)
    bi = synth.lammps_rows_to_internal(box)
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], 9, types, 2, pos.shape[0])
    tr.upload(0, pos, bi)
    cm = oracle.cm_positions(pos, types, 2)
    tr.set_cm(0, cm)
    v, st = tr.msd(primo, nts, lmax, 1, cm_msd, False)
    ref = oracle.msd(pos, types, nts, lmax, primo=primo, skip=1, cm_msd=cm_msd, cm_self=False, ntypes=2, cm=cm)
    assert v.shape == ref.shape
    scale = np.maximum(np.abs(ref[:, 0]), 1e-300)
    assert (np.abs(v[:, 0] - ref[:, 0]) <= 1e-12 * scale).all()
    if cm_msd:
        assert np.array_equal(v[:, 1], ref[:, 1])
    tr.close()


# ---- the small-system kernel (up to 512 device slots) -------------------------------------------------------------
def _small_system(seed, natoms, ntypes, triclinic, nframes, npt=False):
    """`natoms` atoms cut out of a jittered lattice (the box shrunk to match), random types of unequal share."""
    rng = np.random.default_rng([seed, natoms, ntypes])
    cells = (6, 5, 5) if natoms <= 150 else (8, 8, 8)
    pos, box, _ = synth.small_case(seed, cells, 1.1, 1, triclinic, nframes, npt=npt)
    pick = np.sort(rng.choice(pos.shape[1], natoms, replace=False))
    pos = np.ascontiguousarray(pos[:, pick])
    types = np.minimum((rng.random(natoms) ** 2 * ntypes).astype(np.int32), ntypes - 1)
    types[:ntypes] = np.arange(ntypes)   # every type present
    bi = synth.lammps_rows_to_internal(box)
    return pos, bi, types.astype(np.int32)


@pytest.mark.parametrize("natoms,ntypes,triclinic", [
    (5, 1, False), (33, 2, True), (56, 2, False), (64, 1, True),          # one round of 64 slots per job
    (65, 3, False), (97, 2, True), (128, 1, False), (110, 3, True),       # two
    (150, 2, False), (256, 1, True), (300, 2, True), (380, 3, False),     # three, four, five, six
    (440, 1, False), (490, 2, True), (512, 1, False)])                    # seven, eight
def test_small_system_kernel(ctx, natoms, ntypes, triclinic):
    """The small-system kernel (contiguous job ranges per CTA, jobs split over the warps) on 5 .. 512 atoms: same counts
    as the oracle and as the tile kernel, ghost slots in every type group, ragged lag / origin loops."""
    nframes = 40
    pos, bi, types = _small_system(400 + natoms, natoms, ntypes, triclinic, nframes)
    ctx.pbc_wrap(pos, bi)
    args = (0.3, 2.9, 64, 9, 30)   # rmin, rmax, nbin, tmax, ntimesteps: 9 lags x 30 origins
    co = oracle.counts(pos, bi, types, *args, ntypes=ntypes)
    c, st = gpu_counts(ctx, pos, bi, types, ntypes, *args)
    assert np.array_equal(c, co)
    assert st["jobs"] == 9 * 30 and st["pair_evals"] == 9 * 30 * natoms * natoms
    # the default: up to 256 device slots (four warps per job); beyond, up to 512, on request
    assert ran_small(st) == (cabi.device_slots(types) <= cabi.SMALL_DEFAULT_SLOTS)
    if natoms > 150:   # the larger cases: both kernels, both minimum-image paths and the edge count only
        sm = cabi.OPT_SMALL
        c1, st1 = gpu_counts(ctx, pos, bi, types, ntypes, *args, options=sm)
        assert np.array_equal(c1, co) and ran_small(st1) and st1["pair_evals"] == 9 * 30 * natoms * natoms
        c2, st2 = gpu_counts(ctx, pos, bi, types, ntypes, *args, options=cabi.OPT_NO_SMALL)
        assert np.array_equal(c2, co) and not ran_small(st2)
        c4, st4 = gpu_counts(ctx, pos, bi, types, ntypes, *args, options=sm | cabi.OPT_FORCE_GENERAL | cabi.OPT_NO_UBOX)
        assert np.array_equal(c4, co) and ran_small(st4)
        c5, st5, e5 = gpu_counts(ctx, pos, bi, types, ntypes, *args, edges=True, options=sm)
        _, eo = oracle.counts(pos, bi, types, *args, ntypes=ntypes, return_edges=True)
        assert np.array_equal(c5, co) and e5 == eo and ran_small(st5)
        return
    c2, st2 = gpu_counts(ctx, pos, bi, types, ntypes, *args, options=cabi.OPT_NO_SMALL)
    assert np.array_equal(c2, co) and not ran_small(st2)
    # ragged loops (skip does not divide ntimesteps, every > 1) and a block that does not start at frame 0
    kw = dict(primo=3, skip=4, every=2)
    co3 = oracle.counts(pos, bi, types, 0.3, 2.9, 64, 7, 27, ntypes=ntypes, **kw)
    c3, st3 = gpu_counts(ctx, pos, bi, types, ntypes, 0.3, 2.9, 64, 7, 27, **kw)
    assert np.array_equal(c3, co3) and ran_small(st3)
    # every binning mode and both minimum-image paths of the small-system kernel
    for opt in (cabi.OPT_FORCE_GENERAL, cabi.OPT_AGGREGATE, cabi.OPT_NO_SAFE, cabi.OPT_DENSE, cabi.OPT_SPARSE,
                cabi.OPT_NO_UBOX, cabi.OPT_NO_UBOX | cabi.OPT_FORCE_GENERAL | cabi.OPT_NO_SAFE):
        c4, st4 = gpu_counts(ctx, pos, bi, types, ntypes, *args, options=opt)
        assert np.array_equal(c4, co), opt
        assert ran_small(st4)
    c5, st5, e5 = gpu_counts(ctx, pos, bi, types, ntypes, *args, edges=True)
    _, eo = oracle.counts(pos, bi, types, *args, ntypes=ntypes, return_edges=True)
    assert np.array_equal(c5, co) and e5 == eo and ran_small(st5)


@pytest.mark.parametrize("natoms,triclinic,wrap", [(56, False, False), (90, True, False), (70, True, True),
                                                   (200, True, False), (330, False, True)])
def test_small_system_kernel_npt_and_unwrapped(ctx, natoms, triclinic, wrap):
    """A box that changes every frame (the box of the ORIGIN frame serves both atoms) and unwrapped input (the
    literal minimum-image loops) through the small-system kernel."""
    pos, bi, types = _small_system(500 + natoms, natoms, 2, triclinic, 24, npt=True)
    if wrap:
        ctx.pbc_wrap(pos, bi)
    else:
        # every atom whole cells away from where it was, a different number for each atom
        shift = np.random.default_rng(natoms).integers(-2, 3, size=(1, natoms, 3))
        pos = pos + shift * bi[:, None, 3:6] * 2
    args = (0.0, 3.1, 50, 6, 18)
    co = oracle.counts(pos, bi, types, *args, ntypes=2)
    c, st = gpu_counts(ctx, pos, bi, types, 2, *args, options=cabi.OPT_SMALL)   # (330 atoms: beyond the default range)
    assert np.array_equal(c, co) and ran_small(st)
    if not wrap:
        assert st["jobs_fast"] < st["jobs"]


def test_small_system_many_units_per_lag(ctx):
    """Enough origins that every lag is cut into several runs, each merged by a different CTA."""
    pos, bi, types = _small_system(77, 56, 2, False, 1500)
    ctx.pbc_wrap(pos, bi)
    args = (0.7, 3.5, 200, 3, 1400)   # 3 lags x 1400 origins
    co = oracle.counts(pos, bi, types, *args, ntypes=2)
    c, st = gpu_counts(ctx, pos, bi, types, 2, *args)
    assert np.array_equal(c, co) and ran_small(st)
    assert int(c.sum()) == int(co.sum())


@pytest.mark.parametrize("natoms,ntypes,triclinic", [(5, 1, False), (20, 2, True), (37, 3, False), (56, 2, False),
                                                     (72, 2, True), (100, 3, False), (130, 1, True), (200, 2, True)])
@pytest.mark.parametrize("batch,hists", [(8, 1), (8, 3), (3, 2), (2, 8)])
def test_small_system_packed_batches(ctx, monkeypatch, natoms, ntypes, triclinic, batch, hists):
    """The shapes a launch with many jobs takes, forced on a few jobs: `batch` jobs per warp with their i slots packed
    over its lanes (two consecutive slots per thread, up to three jobs within one round), `hists` lag histograms per
    CTA (batches that straddle lags), lag 0 with rmin = 0 (the self pairs are left out of the main pass and counted
    alone), implicit and explicit job lists, every binning mode, both minimum-image paths."""
    monkeypatch.setenv("AGOFRT_SMALL_BATCH", str(batch))
    monkeypatch.setenv("AGOFRT_SMALL_HISTS", str(hists))
    nframes = 40
    pos, bi, types = _small_system(900 + natoms, natoms, ntypes, triclinic, nframes)
    ctx.pbc_wrap(pos, bi)
    for args, kw in (((0.0, 2.9, 64, 11, 25), {}),                                   # 11 lags x 25 origins, self pairs in bin 0
                     ((0.5, 3.0, 40, 9, 27), dict(primo=2, skip=2, every=3))):      # ragged loops, rmin / dr an integer
        co = oracle.counts(pos, bi, types, *args, ntypes=ntypes, **kw)
        for opt in (0, cabi.OPT_EXPLICIT_JOBS, cabi.OPT_FORCE_GENERAL, cabi.OPT_AGGREGATE, cabi.OPT_NO_SAFE,
                    cabi.OPT_DENSE, cabi.OPT_SPARSE | cabi.OPT_NO_UBOX):
            c, st = gpu_counts(ctx, pos, bi, types, ntypes, *args, options=opt | cabi.OPT_SMALL, **kw)
            assert np.array_equal(c, co), (args, opt)
            assert ran_small(st)
        c5, st5, e5 = gpu_counts(ctx, pos, bi, types, ntypes, *args, edges=True, options=cabi.OPT_SMALL, **kw)
        _, eo = oracle.counts(pos, bi, types, *args, ntypes=ntypes, return_edges=True, **kw)
        assert np.array_equal(c5, co) and e5 == eo and ran_small(st5)


@pytest.mark.parametrize("batch,hists", [(8, 2), (4, 4)])
def test_small_system_packed_batches_npt_unwrapped(ctx, monkeypatch, batch, hists):
    """Explicit job lists of both kinds (single-pass and general minimum image) in one block, box changing every
    frame, through packed batches."""
    monkeypatch.setenv("AGOFRT_SMALL_BATCH", str(batch))
    monkeypatch.setenv("AGOFRT_SMALL_HISTS", str(hists))
    natoms = 88
    pos, bi, types = _small_system(1234, natoms, 2, True, 30, npt=True)
    shift = np.random.default_rng(5).integers(-2, 3, size=(1, natoms, 3))
    shift[:, ::3] = 0
    pos = pos + shift * bi[:, None, 3:6] * 2
    args = (0.0, 3.1, 50, 8, 20)
    co = oracle.counts(pos, bi, types, *args, ntypes=2)
    c, st = gpu_counts(ctx, pos, bi, types, 2, *args)
    assert np.array_equal(c, co) and ran_small(st)
    assert 0 < st["jobs_fast"] < st["jobs"] or st["jobs_fast"] in (0, st["jobs"])


# ---- block averages on the device (MediaVar) ------------------------------------------------------------------------
def test_block_average_on_device(ctx):
    """agofrt_blockavg_*: mean and variance of the mean over blocks, bit-identical to the oracle's MediaVar
    (reference lib/include/calcoliblocchi.h:21-65) fed with count*incr of the same blocks."""
    pos, bi, types = _small_system(91, 80, 2, True, 70)
    ctx.pbc_wrap(pos, bi)
    rmin, rmax, nbin, tmax = 0.2, 3.0, 48, 4
    n_b, s, skip = 6, 10, 3
    incr = cabi.gofrt_incr(s, skip)
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], bi.shape[1], types, 2, pos.shape[0])
    tr.upload(0, np.ascontiguousarray(pos), bi)
    plan = cabi.Plan(tr, rmin, rmax, nbin)
    leff = cabi.gofrt_leff(s, tmax)
    acc = cabi.BlockAverage(ctx)
    with pytest.raises(cabi.AgofrtError):
        acc.push(plan, incr)                      # before begin
    acc.begin(leff * 6 * nbin)
    with pytest.raises(cabi.AgofrtError):
        acc.push(plan, incr)                      # no block on the device yet
    blocks = []
    for b in range(n_b):
        none, st = plan.block(b * s, s, leff, skip, 1, options=cabi.OPT_ON_DEVICE)
        assert none is None
        acc.push(plan, incr)
        got = plan.last_counts(leff)
        ref = oracle.counts(pos, bi, types, rmin, rmax, nbin, tmax, s, primo=b * s, skip=skip, ntypes=2)
        assert np.array_equal(got, ref)
        blocks.append(ref * incr)
    mean, var = acc.end(n_b)
    omean, ovar = oracle.mediavar(np.array(blocks))
    assert np.array_equal(mean, omean.ravel())    # bit-identical
    assert np.array_equal(var, ovar.ravel())
    assert (var > 0).any()
    # a block of another size is refused, as VectorOp refuses operands of different sizes
    acc.begin(leff * 6 * nbin)
    plan.block(0, s, leff - 1, skip, 1, options=cabi.OPT_ON_DEVICE)
    with pytest.raises(cabi.AgofrtError):
        acc.push(plan, incr)
    # the accumulator starts over at begin(): one block -> mean = the block, variance 0/0
    acc.begin(leff * 6 * nbin)
    plan.block(0, s, leff, skip, 1, options=cabi.OPT_ON_DEVICE)
    acc.push(plan, incr)
    mean1, var1 = acc.end(1)
    assert np.array_equal(mean1, blocks[0].ravel()) and np.isnan(var1).all()
    acc.close()
    plan.close()
    tr.close()


@pytest.mark.parametrize("natoms,tri", [(80, True), (56, False), (300, True)])
def test_batch_of_whole_blocks(ctx, natoms, tri):
    """agofrt_blocks: every block of a batch equals the oracle's block, the device Welford over the batch
    (one launch) is bit-identical to the oracle's MediaVar, and the last block is what agofrt_block would have left."""
    pos, bi, types = _small_system(93, natoms, 2, tri, 70)
    ctx.pbc_wrap(pos, bi)
    rmin, rmax, nbin, tmax = 0.0, 3.0, 40, 5
    n_b, s, skip = 6, 10, 3
    incr = cabi.gofrt_incr(s, skip)
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], bi.shape[1], types, 2, pos.shape[0])
    tr.upload(0, np.ascontiguousarray(pos), bi)
    plan = cabi.Plan(tr, rmin, rmax, nbin)
    leff = cabi.gofrt_leff(s, tmax)
    for options in (0, cabi.OPT_FORCE_GENERAL):
        st = plan.blocks(0, s, n_b, s, leff, skip, 1, options=options)
        assert st["jobs"] == n_b * leff * 4 and st["launches"] >= n_b
        assert st["jobs_fast"] == (0 if options else st["jobs"])
        blocks = []
        for b in range(n_b):
            ref = oracle.counts(pos, bi, types, rmin, rmax, nbin, tmax, s, primo=b * s, skip=skip, ntypes=2)
            assert np.array_equal(plan.block_counts(b, leff), ref)
            blocks.append(ref * incr)
        assert np.array_equal(plan.last_counts(leff), plan.block_counts(n_b - 1, leff))
        acc = cabi.BlockAverage(ctx)
        acc.begin(leff * 6 * nbin)
        acc.push_blocks(plan, incr)
        mean, var = acc.end(n_b)
        omean, ovar = oracle.mediavar(np.array(blocks))
        assert np.array_equal(mean, omean.ravel()) and np.array_equal(var, ovar.ravel())
        acc.close()
    with pytest.raises(cabi.AgofrtError):
        plan.block_counts(n_b, leff)
    # a batch whose blocks have no regular job list is refused (the caller runs them one by one)
    far = np.ascontiguousarray(pos).copy()
    far[::2, 0, 0] += 5.0 * 2 * bi[0, 3]   # atom 0 far outside the cell in every other frame: no range-wide proof
    tr.upload(0, far, bi)
    with pytest.raises(cabi.AgofrtError):
        plan.blocks(0, s, n_b, s, leff, skip, 1)
    plan.close()
    tr.close()


@pytest.mark.parametrize("tri", [False, True])
def test_window_from_raw_dump_records(ctx, tri):
    """agofrt_traj_upload_records: the frame loop of Trajectory::set_access_at on the device -- records in a different
    order in every frame, split over several chunks, ids that are not 0..N-1 -- gives the window the host reader gives."""
    pos, box, types = synth.small_case(97, (6, 5, 4), 1.08, 2, tri, 9)
    bi = synth.lammps_rows_to_internal(box)
    n = pos.shape[1]
    rng = np.random.default_rng(5)
    ids = rng.permutation(3 * n)[:n].astype(np.int32) + 1        # sparse-ish, shuffled, starting above 0
    raw_type = (types * 2 + 3).astype(np.int32)
    frames = []
    for f in range(pos.shape[0]):
        order = rng.permutation(n)
        rec = np.zeros((n, 8))
        rec[:, 0] = ids[order]
        rec[:, 1] = raw_type[order]
        rec[:, 2:5] = pos[f, order]
        rec[:, 5:8] = rng.normal(size=(n, 3))                    # velocities: ignored
        cuts = [0, n // 3, n // 3 + 7, n]
        frames.append([rec[cuts[k]:cuts[k + 1]] for k in range(3)])
    tr = cabi.DeviceTrajectory(ctx, n, bi.shape[1], types, 2, pos.shape[0])
    with pytest.raises(cabi.AgofrtError):
        tr.upload_records(0, frames, bi)                         # no id table yet
    tr.set_ids(ids, raw_type)
    back = np.empty_like(pos)
    tr.upload_records(0, frames, bi, wrap=True, out=back)
    wrapped = oracle.pbc_wrap(pos, bi)
    assert np.array_equal(back, wrapped) and np.array_equal(tr.download(0, pos.shape[0]), wrapped)
    plan = cabi.Plan(tr, 0.0, 2.8, 48)
    c, st = plan.block(1, 5, 3, 2, 1)
    assert np.array_equal(c, oracle.counts(wrapped, bi, types, 0.0, 2.8, 48, 3, 5, primo=1, skip=2, ntypes=2))
    # without the wrap the parsed frames come back as they are in the file
    tr.upload_records(0, frames, bi, out=back)
    assert np.array_equal(back, pos)
    # an id that is not in the table, and a changed type
    bad = [[c.copy() for c in fr] for fr in frames]
    bad[4][1][2, 0] = 10 * n + 5
    with pytest.raises(cabi.AgofrtError) as e:
        tr.upload_records(0, bad, bi)
    assert e.value.code == cabi.ERR_ARG
    bad = [[c.copy() for c in fr] for fr in frames]
    bad[7][0][0, 1] += 1
    with pytest.raises(cabi.AgofrtError) as e:
        tr.upload_records(0, bad, bi)
    assert e.value.code == cabi.ERR_RETYPED
    plan.close()
    tr.close()


def test_c4_subset_against_the_reference_itself(ctx):
    """The north-star shape at full size -- 100 000 atoms, triclinic, 500 bins -- on the (lag, origin) subset that is
    bench.py's default step (26 lags x 8 origins = 2.08e12 pair evaluations): bit-exact against the counts of the
    UNMODIFIED reference (69 minutes on 7 host threads; tests/golden/make_c4_subset_golden.py)."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    meta = json.load(open(os.path.join(GOLDEN, "c4_subset_counts.json")))
    gold = np.load(os.path.join(GOLDEN, "c4_subset_counts.npz"))
    w, nts = synth.bench_subset("C4")
    assert meta["workload"] == w.name and meta["subset"] == {"ntimesteps": nts, "skip": w.skip, "every": w.every, "leff": 201}
    leff = 201
    nframes = (nts - 1) // w.skip * w.skip + (leff - 1) // w.every * w.every + 1
    pos, box, types = synth.generate(w, nframes=nframes)
    bi = synth.lammps_rows_to_internal(box)
    tr = cabi.DeviceTrajectory(ctx, w.natoms, 9, types, 1, nframes)
    tr.upload_ex(0, pos, bi, wrap=True)
    plan = cabi.Plan(tr, w.rmin, w.rmax, w.nbin)
    c, st = plan.block(0, nts, leff, w.skip, w.every)
    plan.close()
    tr.close()
    assert st["jobs"] == meta["jobs"] == 208 and st["jobs_fast"] == 208
    assert np.array_equal(c[gold["lags"]], gold["counts"])
    assert hashlib.sha256(np.ascontiguousarray(c).astype("<u8").tobytes()).hexdigest() == meta["counts_sha256"]
    assert int(c.sum()) == meta["counts_sum"]


@pytest.mark.parametrize("name", ["C2", "C3"])
@pytest.mark.parametrize("options", [0, cabi.OPT_FORCE_GENERAL, cabi.OPT_NO_SAFE], ids=["default", "general", "thresholds"])
def test_full_n_against_the_reference_itself(ctx, name, options):
    """BASELINE.json configs[1] (C2: 4096 atoms, cubic, dense) and configs[2] (C3: 12 288 atoms, 3 types, triclinic) at
    their full atom counts, on the atoms / cell / bins bench.py uses for them, a few (lag, origin) jobs: the tile kernel
    against the counts of the UNMODIFIED reference (tests/golden/make_full_n_golden.py), wrap on the device."""
    import hashlib
    import json
    import os
    from conftest import GOLDEN
    meta = json.load(open(os.path.join(GOLDEN, "full_n_counts.json")))[name]
    gold = np.load(os.path.join(GOLDEN, "full_n_counts.npz"))[name + "/counts"]
    w = synth.WORKLOADS[name]
    pos, box, types = synth.generate(w, nframes=meta["frames"])
    assert hashlib.sha256(np.ascontiguousarray(pos).tobytes()).hexdigest() == meta["pos_in_sha256"]
    bi = synth.lammps_rows_to_internal(box)
    tr = cabi.DeviceTrajectory(ctx, w.natoms, bi.shape[1], types, w.ntypes, meta["frames"])
    tr.upload_ex(0, pos, bi, wrap=True)
    plan = cabi.Plan(tr, w.rmin, w.rmax, w.nbin)
    leff = cabi.gofrt_leff(meta["ntimesteps"], meta["tmax"])
    c, st = plan.block(meta["primo"], meta["ntimesteps"], leff, meta["skip"], meta["every"], options=options)
    plan.close()
    tr.close()
    assert not ran_small(st)
    assert np.array_equal(c, gold)
    assert int(c.sum()) == meta["counts_sum"]


# ---- the other pair loops over d2_minImage: neighbour lists and spherical-harmonic densities --------------------------
@pytest.mark.parametrize("name", sorted(PAIR_LOOP_CASES))
def test_neighbour_lists_vs_the_reference(ctx, name):
    """agofrt_neighbours against Neighbours::update_neigh of the compiled reference: counts, partner indices in the
    reference's order (ascending index, or ascending distance), distances and minimum-image vectors bit for bit."""
    d = pair_loop_case(name)
    bi = synth.lammps_rows_to_internal(d["box"])
    pos = np.ascontiguousarray(d["pos"]).copy()
    ctx.pbc_wrap(pos, bi)
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], bi.shape[1], d["types"], d["ntypes"], pos.shape[0])
    tr.upload(0, pos, bi)
    spec = [(s[0], s[1]) for s in d["spec"]]
    for sort, tag in ((False, "unsorted_"), (True, "sorted_")):
        counts, idx, r = tr.neighbours(d["frame"], spec, sort=sort)
        assert np.array_equal(counts, d[tag + "counts"])
        assert np.array_equal(idx, d[tag + "idx"])
        assert np.array_equal(r, d[tag + "r"])
    assert counts.sum() > 0
    # a list that overflows is the reference's exception
    with pytest.raises(cabi.AgofrtError) as e:
        tr.neighbours(d["frame"], [(2, s[1]) for s in d["spec"]])
    assert e.value.code == cabi.ERR_TOO_LARGE and "Too many neighbours in shell" in str(e.value)
    tr.close()


@pytest.mark.parametrize("name", sorted(PAIR_LOOP_CASES))
def test_spherical_harmonic_density_vs_the_reference(ctx, name):
    """agofrt_sh_density against SphericalBase<l,double,Trajectory_numpy>::calc of the compiled reference (l = 4, 6, 10):
    the per-atom, per-type, per-bin sums of real spherical harmonics and the neighbour counters, bit for bit."""
    d = pair_loop_case(name)
    bi = synth.lammps_rows_to_internal(d["box"])
    pos = np.ascontiguousarray(d["pos"]).copy()
    ctx.pbc_wrap(pos, bi)
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], bi.shape[1], d["types"], d["ntypes"], pos.shape[0])
    tr.upload(0, pos, bi)
    res, cnt = tr.sh_density(d["frame"], d["lmax"], d["nbin"], d["rminmax"])
    assert np.array_equal(cnt, d["sh_counter"]) and cnt.sum() > 0
    assert res.shape == d["sh"].shape
    bad = res != d["sh"]
    assert not bad.any(), (int(bad.sum()), float(np.abs(res - d["sh"]).max()))
    tr.close()
