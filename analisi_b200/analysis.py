"""Python-side helpers of the g(r,t) path, over the ``pyanalisi`` extension of this repository
(``python/pyanalisi*.so``): the counterparts of the reference's pure-python wrappers

    Analysis.max_l / Analysis.hist2gofr / Analysis.compute_gofr   (pyanalisi/analysis.py:77-144)
    Analysis.compute_msd                                           (pyanalisi/analysis.py:91-98)
    analyze_gofr                                                   (pyanalisi/common.py:186-204)

with the same argument meaning.  Nothing here computes pairs: ``Gofrt.calculate`` runs on the GPU.
The reference's callers pass ``(..., nthreads, tskip, False, 1)`` positionally, which lands on
``(skip, every=False -> 1, debug=1 -> True)`` and makes every calculate() append to ./gofrt.dump
(SURVEY.md section 8b); the helpers below pass the arguments by their real meaning (debug off).
"""
import numpy as np


def _ext():
    import pyanalisi
    return pyanalisi


def wrapper_name(traj, name="Gofrt"):
    """Class for this trajectory kind: ``Gofrt`` for numpy trajectories, ``Gofrt_lammps`` for mmap ones
    (the reference's Analysis.pyanalisi_wrapper / common.pyanalisi_wrapper)."""
    pa = _ext()
    if isinstance(traj, pa.Trajectory):
        return getattr(pa, name)
    if isinstance(traj, pa.Traj):
        return getattr(pa, name + "_lammps")
    raise RuntimeError("Wrapper for trajectory class not implemented")


def max_l(start, stop, tmax=0):
    """Split [start, stop) into the number of lags and the number of averaged origins (analysis.py:77-88)."""
    if tmax <= 0:
        tmax = (stop - start) // 2
    n_ave = stop - start - tmax
    if start >= stop:
        raise RuntimeError("start index must be less than the end index")
    if n_ave <= 0:
        tmax = stop - start - 1
        n_ave = 1
    return tmax, n_ave


def hist2gofr(gr_N, gr_dr, gr_0, gofr):
    """Histogram -> g(r): division by the shell volumes 4 pi/3 (r+^3 - r-^3), the only normalisation the
    reference applies on the python side (analysis.py:100-105; no density or N factor)."""
    rs_m = np.arange(gr_N) * gr_dr + gr_0
    rs_p = (np.arange(gr_N) + 1) * gr_dr + gr_0
    vols = 4 * np.pi / 3 * (rs_p ** 3 - rs_m ** 3)
    return gofr / vols


def compute_gofr(traj, startr, endr, nbin, start=0, stop=None, tmax=1, tskip=10, n_segments=1, nthreads=1,
                 return_histogram=False):
    """g(r) / van Hove g(r,t) of a ``pyanalisi.Trajectory`` or ``pyanalisi.Traj`` (analysis.py:108-144):
    ``tmax`` time lags, origins every ``tskip`` frames; ``n_segments`` > 1 returns one result per segment."""
    if stop is None:
        stop = traj.get_nloaded_timesteps()
    tmax, n_ave = max_l(start, stop, tmax)
    gofr = wrapper_name(traj)(traj, startr, endr, nbin, tmax, nthreads, tskip, 1, False)
    conv = (lambda h: h) if return_histogram else (lambda h: hist2gofr(nbin, (endr - startr) / nbin, startr, h))
    if n_segments == 1:
        gofr.reset(n_ave)
        gofr.calculate(start)
        return conv(np.array(gofr, copy=True))
    if n_segments > 1:
        res = []
        segment_size = max(1, n_ave // n_segments)
        gofr.reset(segment_size)
        for i in range(0, min(segment_size * n_segments, n_ave), segment_size):
            gofr.calculate(start + i)
            res.append(conv(np.array(gofr, copy=True)))
        return res
    raise IndexError("n_segments must be > 0 (%r)" % (n_segments,))


def analyze_gofr(traj, start, stop, startr, endr, nbin, tmax=1, nthreads=1, tskip=10, n_segments=1):
    """The raw-histogram variant (common.py:186-204)."""
    return compute_gofr(traj, startr, endr, nbin, start=start, stop=stop, tmax=tmax, tskip=tskip, n_segments=n_segments,
                        nthreads=nthreads, return_histogram=True)


def compute_msd(traj, start=0, stop=-1, tmax=0, tskip_msd=10, center_of_mass_MSD=True, center_of_mass_frame=False, nthreads=1):
    """Mean square displacement per type (and of the per-type centres of mass) of an UNWRAPPED trajectory
    (analysis.py:91-98): array (tmax, 2 if center_of_mass_MSD else 1, ntypes)."""
    if stop < 0:
        stop = traj.get_nloaded_timesteps()
    tmax, n_ave = max_l(start, stop, tmax)
    msd = wrapper_name(traj, "MeanSquareDisplacement")(traj, tskip_msd, tmax, nthreads, center_of_mass_MSD, center_of_mass_frame, False)
    msd.reset(n_ave)
    msd.calculate(start)
    return np.array(msd, copy=True)
