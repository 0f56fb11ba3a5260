"""Python-side conveniences around the ``pyanalisi`` extension of this repository (``python/pyanalisi*.so``).

They play the role of the reference's pure-python wrappers for this path -- ``Analysis.compute_gofr`` /
``compute_msd`` (pyanalisi/analysis.py:91-144) and ``analyze_gofr`` (pyanalisi/common.py:186-204) -- with the same
argument meaning and the same results, written for this repository:

* a window ``[start, stop)`` of frames is split into lags and averaged origins by :func:`lag_split`;
* the histogram that ``Gofrt`` returns is turned into g(r) by dividing every bin by the volume of its spherical
  shell (:func:`shell_normalise`) -- the only normalisation the reference applies on the python side (no density, no
  atom count);
* ``n_segments > 1`` gives one result per consecutive segment of origins.

Nothing here computes pairs: ``Gofrt.calculate`` runs on the GPU.  The reference's own callers pass
``(..., nthreads, tskip, False, 1)`` positionally, which lands on ``every=False`` and ``debug=1`` and makes every
calculate() append to ./gofrt.dump; the helpers below pass ``every=1, debug=False``.
"""
import math

import numpy as np


def _ext():
    import pyanalisi
    return pyanalisi


def wrapper_name(traj, name="Gofrt"):
    """The extension class that goes with this trajectory object: ``<name>`` for ``pyanalisi.Trajectory`` (numpy
    arrays), ``<name>_lammps`` for ``pyanalisi.Traj`` (LAMMPS binary through mmap)."""
    pa = _ext()
    for cls, suffix in ((pa.Trajectory, ""), (pa.Traj, "_lammps")):
        if isinstance(traj, cls):
            return getattr(pa, name + suffix)
    raise RuntimeError("Wrapper for trajectory class not implemented")


def lag_split(start, stop, tmax=0):
    """(number of lags, number of averaged origins) for the frames ``[start, stop)``.

    ``tmax <= 0`` asks for half of the window.  The origins are what is left after the lags; a window too short
    for even one origin keeps a single origin and as many lags as then fit."""
    span = stop - start
    if span <= 0:
        raise RuntimeError("start index must be less than the end index")
    lags = tmax if tmax > 0 else span // 2
    origins = span - lags
    if origins < 1:
        lags, origins = span - 1, 1
    return lags, origins


max_l = lag_split   # the reference's name (Analysis.max_l)


def shell_normalise(hist, rmin, dr):
    """Divide bin k (last axis) by the volume of the shell ``rmin + k dr <= r < rmin + (k+1) dr``."""
    hist = np.asarray(hist, dtype=np.float64)
    edges = rmin + dr * np.arange(hist.shape[-1] + 1)
    return hist / (4.0 * math.pi / 3.0 * np.diff(edges ** 3))


def hist2gofr(gr_N, gr_dr, gr_0, gofr):
    """The reference's signature (Analysis.hist2gofr): number of bins, bin width, first edge, histogram."""
    gofr = np.asarray(gofr)
    if gofr.shape[-1] != gr_N:
        raise ValueError("the histogram has %d bins, not %d" % (gofr.shape[-1], gr_N))
    return shell_normalise(gofr, gr_0, gr_dr)


def compute_gofr(traj, startr, endr, nbin, start=0, stop=None, tmax=1, tskip=10, n_segments=1, nthreads=1,
                 return_histogram=False):
    """g(r) -- or the van Hove function g(r,t) for ``tmax > 1`` -- of a ``pyanalisi.Trajectory`` / ``pyanalisi.Traj``:
    ``tmax`` time lags, one origin every ``tskip`` frames, ``nbin`` bins between ``startr`` and ``endr``.
    One array ``(lags, ntypes*(ntypes+1), nbin)``, or a list of ``n_segments`` of them."""
    if n_segments < 1:
        raise IndexError("n_segments must be > 0 (%r)" % (n_segments,))
    if stop is None:
        stop = traj.get_nloaded_timesteps()
    lags, origins = lag_split(start, stop, tmax)
    calc = wrapper_name(traj)(traj, startr, endr, nbin, lags, nthreads, tskip, 1, False)
    dr = (endr - startr) / nbin

    def one(first):
        calc.calculate(first)
        h = np.array(calc, copy=True)
        return h if return_histogram else shell_normalise(h, startr, dr)

    if n_segments == 1:
        calc.reset(origins)
        return one(start)
    per_segment = max(1, origins // n_segments)
    calc.reset(per_segment)
    firsts = [start + k * per_segment for k in range(n_segments) if (k + 1) * per_segment <= origins]
    return [one(f) for f in firsts]


def analyze_gofr(traj, start, stop, startr, endr, nbin, tmax=1, nthreads=1, tskip=10, n_segments=1):
    """The raw histogram instead of g(r) (the reference's common.analyze_gofr)."""
    return compute_gofr(traj, startr, endr, nbin, start=start, stop=stop, tmax=tmax, tskip=tskip, n_segments=n_segments,
                        nthreads=nthreads, return_histogram=True)


def compute_msd(traj, start=0, stop=-1, tmax=0, tskip_msd=10, center_of_mass_MSD=True, center_of_mass_frame=False, nthreads=1):
    """Mean square displacement per type (and of the per-type centres of mass) of an UNWRAPPED trajectory:
    array ``(lags, 2 if center_of_mass_MSD else 1, ntypes)``."""
    if stop < 0:
        stop = traj.get_nloaded_timesteps()
    lags, origins = lag_split(start, stop, tmax)
    calc = wrapper_name(traj, "MeanSquareDisplacement")(traj, tskip_msd, lags, nthreads, center_of_mass_MSD,
                                                         center_of_mass_frame, False)
    calc.reset(origins)
    calc.calculate(start)
    return np.array(calc, copy=True)
