"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front end of ``libgofrt_oracle.so``.

See ``gofrt_oracle.h`` for the reference file:line each function restates.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgofrt_oracle.so")
_lib = None


class OracleError(RuntimeError):
    pass


class _Traj(C.Structure):
    _fields_ = [
        ("natoms", C.c_size_t),
        ("nframes", C.c_size_t),
        ("first_frame", C.c_size_t),
        ("total_frames", C.c_size_t),
        ("ntypes", C.c_int),
        ("triclinic", C.c_int),
        ("pos", C.c_void_p),
        ("box", C.c_void_p),
        ("type_id", C.c_void_p),
    ]


class _Params(C.Structure):
    _fields_ = [
        ("rmin", C.c_double),
        ("rmax", C.c_double),
        ("nbin", C.c_uint),
        ("lmax", C.c_uint),
        ("skip", C.c_uint),
        ("every", C.c_uint),
    ]


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    src = os.path.join(_HERE, "gofrt_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "gofrt_oracle.h"))
    )
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _LIB_PATH


def build_ref(reference="/root/reference"):
    """Compile the unmodified reference into oracle/_ref (only where the tree exists)."""
    if not os.path.isdir(reference):
        return None
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref", "REF=" + reference])
    return os.path.join(_HERE, "_ref")


def load_ref():
    """Import oracle/_ref/analisi_ref*.so (the compiled reference) or return None."""
    d = os.path.join(_HERE, "_ref")
    if not os.path.isdir(d) or not any(f.startswith("analisi_ref") and f.endswith(".so") for f in os.listdir(d)):
        return None
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        return importlib.import_module("analisi_ref")
    except ImportError:
        return None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = C.CDLL(_LIB_PATH)
    dp = C.POINTER(C.c_double)
    lib.gofrt_oracle_neighbour_hist.argtypes = [C.c_void_p, C.c_double, C.c_size_t, C.c_uint, C.c_uint,
                                                C.POINTER(C.c_uint64), C.c_uint]
    lib.gofrt_oracle_cm.argtypes = [dp, C.POINTER(C.c_int), C.c_size_t, C.c_int, dp]
    lib.gofrt_oracle_cm.restype = None
    lib.gofrt_oracle_msd.argtypes = [C.c_void_p, dp, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, C.c_int, C.c_int, dp]
    lib.gofrt_oracle_lammps_to_internal.argtypes = [dp]
    lib.gofrt_oracle_internal_to_lammps.argtypes = [dp]
    lib.gofrt_oracle_min_image.argtypes = [dp, dp, dp, C.c_int]
    lib.gofrt_oracle_d2.argtypes = [dp, dp, dp, dp, C.c_int, dp]
    lib.gofrt_oracle_d2.restype = C.c_double
    lib.gofrt_oracle_pbc_wrap.argtypes = [dp, C.c_size_t, dp, C.c_int]
    lib.gofrt_oracle_type_ids.argtypes = [C.POINTER(C.c_int), C.c_size_t, C.POINTER(C.c_int)]
    lib.gofrt_oracle_type_ids.restype = C.c_int
    lib.gofrt_oracle_itype.argtypes = [C.c_uint] * 3
    lib.gofrt_oracle_itype.restype = C.c_uint
    lib.gofrt_oracle_nextra.argtypes = [C.c_size_t, C.c_uint, C.c_uint]
    lib.gofrt_oracle_nextra.restype = C.c_uint
    lib.gofrt_oracle_leff.argtypes = [C.c_uint, C.c_uint]
    lib.gofrt_oracle_leff.restype = C.c_uint
    lib.gofrt_oracle_counts.argtypes = [
        C.POINTER(_Traj), C.POINTER(_Params), C.c_size_t, C.c_uint,
        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_uint,
    ]
    lib.gofrt_oracle_counts.restype = C.c_int
    lib.gofrt_oracle_vdata.argtypes = [
        C.POINTER(_Traj), C.POINTER(_Params), C.c_size_t, C.c_uint, dp, C.c_uint,
    ]
    lib.gofrt_oracle_vdata.restype = C.c_int
    lib.gofrt_oracle_mediavar.argtypes = [dp, C.c_uint, C.c_size_t, dp, dp]
    _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _traj(pos, box, type_id, ntypes, first_frame, total_frames):
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    box = np.ascontiguousarray(box, dtype=np.float64)
    type_id = np.ascontiguousarray(type_id, dtype=np.int32)
    assert pos.ndim == 3 and pos.shape[2] == 3
    assert box.ndim == 2 and box.shape[0] == pos.shape[0] and box.shape[1] in (6, 9)
    assert type_id.shape == (pos.shape[1],)
    if ntypes is None:
        ntypes = int(type_id.max()) + 1
    t = _Traj(
        pos.shape[1], pos.shape[0], first_frame,
        pos.shape[0] + first_frame if total_frames is None else total_frames,
        ntypes, 1 if box.shape[1] == 9 else 0,
        pos.ctypes.data, box.ctypes.data, type_id.ctypes.data,
    )
    return t, (pos, box, type_id), ntypes


def _check(rc):
    if rc == -1:
        # the reference's std::runtime_error text, gofrt.cpp:82
        raise OracleError("trajectory is too short for this kind of calculation. Select a different starting "
                          "timestep or lower the size of the average or the lenght of the time lag")
    if rc != 0:
        raise OracleError("bad argument (rc=%d)" % rc)


def counts(pos, box_internal, type_id, rmin, rmax, nbin, lmax, ntimesteps, primo=0, skip=1, every=1,
           ntypes=None, first_frame=0, total_frames=None, nthreads=None, return_edges=False):
    """uint64 [leff][ntypes*(ntypes+1)][nbin] of ``reset(ntimesteps); calculate(primo)``."""
    lib = _load()
    t, keep, ntypes = _traj(pos, box_internal, type_id, ntypes, first_frame, total_frames)
    p = _Params(rmin, rmax, nbin, lmax, skip, every)
    le = lib.gofrt_oracle_leff(ntimesteps, lmax)
    out = np.zeros((le, ntypes * (ntypes + 1), nbin), dtype=np.uint64)
    edges = C.c_uint64(0)
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    rc = lib.gofrt_oracle_counts(C.byref(t), C.byref(p), primo, ntimesteps,
                                 out.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(edges), nthreads)
    _check(rc)
    del keep
    return (out, int(edges.value)) if return_edges else out


def vdata(pos, box_internal, type_id, rmin, rmax, nbin, lmax, ntimesteps, primo=0, skip=1, every=1,
          ntypes=None, first_frame=0, total_frames=None, ref_nthreads=1):
    """float64 result exactly as the reference accumulates it with ``ref_nthreads`` threads."""
    lib = _load()
    t, keep, ntypes = _traj(pos, box_internal, type_id, ntypes, first_frame, total_frames)
    p = _Params(rmin, rmax, nbin, lmax, skip, every)
    le = lib.gofrt_oracle_leff(ntimesteps, lmax)
    out = np.zeros((le, ntypes * (ntypes + 1), nbin), dtype=np.float64)
    rc = lib.gofrt_oracle_vdata(C.byref(t), C.byref(p), primo, ntimesteps, _dp(out), ref_nthreads)
    _check(rc)
    del keep
    return out


def mediavar(blocks):
    """MediaVar over blocks[n_b, ...] -> (mean, variance of the mean)."""
    lib = _load()
    b = np.ascontiguousarray(blocks, dtype=np.float64)
    n_b = b.shape[0]
    flat = b.reshape(n_b, -1)
    mean = np.zeros(flat.shape[1])
    var = np.zeros(flat.shape[1])
    lib.gofrt_oracle_mediavar(_dp(flat), n_b, flat.shape[1], _dp(mean), _dp(var))
    return mean.reshape(b.shape[1:]), var.reshape(b.shape[1:])


def neighbour_hist(pos, box_internal, type_id, r, tstart, ntimesteps, skip=1, ntypes=None, first_frame=0, nthreads=None,
                   hist=None):
    """IstogrammaAtomiRaggio::calculate (lib/src/istogrammaatomiraggio.cpp:31-85): hist[type][count] over frames
    tstart, tstart+skip, ... < tstart+ntimesteps; returns (and adds to) a uint64 array [ntypes][natoms+1]."""
    lib = _load()
    t, keep, ntypes = _traj(pos, box_internal, type_id, ntypes, first_frame, None)
    if hist is None:
        hist = np.zeros((ntypes, pos.shape[1] + 1), dtype=np.uint64)
    if nthreads is None:
        nthreads = os.cpu_count() or 1
    rc = lib.gofrt_oracle_neighbour_hist(C.byref(t), float(r), int(tstart), int(ntimesteps), int(skip),
                                         hist.ctypes.data_as(C.POINTER(C.c_uint64)), int(nthreads))
    _check(rc)
    del keep
    return hist


def cm_positions(pos, type_id, ntypes):
    """Per-type centres of mass [F][ntypes][3], running mean in atom order (trajectory_numpy.cpp:201-223)."""
    lib = _load()
    pos = np.ascontiguousarray(pos, dtype=np.float64)
    tid = np.ascontiguousarray(type_id, dtype=np.int32)
    out = np.zeros((pos.shape[0], ntypes, 3), dtype=np.float64)
    for f in range(pos.shape[0]):
        lib.gofrt_oracle_cm(_dp(pos[f]), tid.ctypes.data_as(C.POINTER(C.c_int)), pos.shape[1], ntypes, _dp(out[f]))
    return out


def msd(pos, type_id, ntimesteps, lmax=0, primo=0, skip=1, cm_msd=False, cm_self=False, ntypes=None, cm=None,
        first_frame=0, total_frames=None):
    """MSD<T>::calc_single_th over all lags (lib/src/msd.cpp:63-125): returns vdata [leff][f_cm][ntypes]."""
    lib = _load()
    box = np.tile([0.0, 0.0, 0.0, 1.0, 1.0, 1.0], (pos.shape[0], 1))   # not used by the MSD
    t, keep, ntypes = _traj(pos, box, type_id, ntypes, first_frame, total_frames)
    if cm is None and (cm_msd or cm_self):
        cm = cm_positions(pos, type_id, ntypes)
    le = ntimesteps if (ntimesteps < lmax or lmax == 0) else lmax
    out = np.zeros((le, 2 if cm_msd else 1, ntypes), dtype=np.float64)
    cmp_ = _dp(np.ascontiguousarray(cm)) if cm is not None else None
    rc = lib.gofrt_oracle_msd(C.byref(t), cmp_, int(primo), int(ntimesteps), int(lmax), int(skip), int(bool(cm_msd)),
                              int(bool(cm_self)), _dp(out))
    _check(rc)
    del keep
    return out


def min_image(delta, box_row):
    lib = _load()
    d = np.array(delta, dtype=np.float64)
    b = np.ascontiguousarray(box_row, dtype=np.float64)
    tri = 1 if b.shape[0] == 9 else 0
    tilt = b[6:] if tri else np.zeros(3)
    lh = np.ascontiguousarray(b[3:6])
    tilt = np.ascontiguousarray(tilt)
    lib.gofrt_oracle_min_image(_dp(d), _dp(lh), _dp(tilt), tri)
    return d


def d2_all(pos_i, pos_j, box_row):
    """(N,N,4) array of (dx,dy,dz,d2) for every ordered pair, box of frame i."""
    lib = _load()
    pi = np.ascontiguousarray(pos_i, dtype=np.float64)
    pj = np.ascontiguousarray(pos_j, dtype=np.float64)
    b = np.ascontiguousarray(box_row, dtype=np.float64)
    tri = 1 if b.shape[0] == 9 else 0
    lh = np.ascontiguousarray(b[3:6])
    tilt = np.ascontiguousarray(b[6:9]) if tri else np.zeros(3)
    n = pi.shape[0]
    out = np.zeros((n, n, 4))
    x = np.zeros(3)
    for i in range(n):
        for j in range(n):
            d2 = lib.gofrt_oracle_d2(_dp(pi[i]), _dp(pj[j]), _dp(lh), _dp(tilt), tri, _dp(x))
            out[i, j, :3] = x
            out[i, j, 3] = d2
    return out


def pbc_wrap(pos, box_internal):
    """Wrapped copy of pos[F,N,3] (each frame with its own box row)."""
    lib = _load()
    p = np.array(pos, dtype=np.float64, order="C")
    b = np.ascontiguousarray(box_internal, dtype=np.float64)
    tri = 1 if b.shape[1] == 9 else 0
    for f in range(p.shape[0]):
        row = np.zeros(9)
        row[: b.shape[1]] = b[f]
        lib.gofrt_oracle_pbc_wrap(_dp(p[f]), p.shape[1], _dp(row), tri)
    return p


def type_ids(raw):
    lib = _load()
    r = np.ascontiguousarray(raw, dtype=np.int32)
    out = np.zeros_like(r)
    nt = lib.gofrt_oracle_type_ids(r.ctypes.data_as(C.POINTER(C.c_int)), r.shape[0],
                                   out.ctypes.data_as(C.POINTER(C.c_int)))
    return out, nt


def itype(ntypes, t1, t2):
    return int(_load().gofrt_oracle_itype(ntypes, t1, t2))


def nextra(total_frames, n_b, lmax):
    return int(_load().gofrt_oracle_nextra(total_frames, n_b, lmax))


def leff(ntimesteps, lmax):
    return int(_load().gofrt_oracle_leff(ntimesteps, lmax))


def lammps_to_internal(rows):
    lib = _load()
    b = np.array(rows, dtype=np.float64, order="C")
    flat = b.reshape(-1, b.shape[-1])
    for r in flat:
        tmp = np.ascontiguousarray(r[:6])
        lib.gofrt_oracle_lammps_to_internal(_dp(tmp))
        r[:6] = tmp
    return b


def internal_to_lammps(rows):
    lib = _load()
    b = np.array(rows, dtype=np.float64, order="C")
    flat = b.reshape(-1, b.shape[-1])
    for r in flat:
        tmp = np.ascontiguousarray(r[:6])
        lib.gofrt_oracle_internal_to_lammps(_dp(tmp))
        r[:6] = tmp
    return b
