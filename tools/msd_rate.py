#!/usr/bin/env python
"""Throughput of the MSD kernels (agofrt_msd) against the HBM roofline: algorithmic bytes = 48 B per (atom, lag,
origin) displacement (two float64 positions read), from the library's CUDA-event timing.  python tools/msd_rate.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from analisi_b200 import cabi, synth  # noqa: E402

try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = None
ctx = cabi.Context([0])
fp64 = ctx.fp64_peak(0.5)   # DFMA/s measured here: 9 FP64 operations per displacement is the other bound
# (label, workload, frames in the window, averaged steps, lags); "C4 full" is the north-star window: 2.3 GB, far larger than L2
CASES = (("C2", "C2", 400, 300, 100), ("C4", "C4", 60, 40, 20), ("C4full", "C4", 957, 757, 200))
only = sys.argv[1:]
for label, name, nframes, nts, lmax in CASES:
    if only and label not in only:
        continue
    w = synth.WORKLOADS[name]
    pos, box, types = synth.generate(w, nframes=nframes)
    bi = synth.lammps_rows_to_internal(box)
    pos = np.ascontiguousarray(pos)
    tr = cabi.DeviceTrajectory(ctx, w.natoms, bi.shape[1], types, w.ntypes, nframes)
    tr.upload(0, pos, bi)
    tr.msd(0, nts, lmax, 1)
    v, st = tr.msd(0, nts, lmax, 1)
    nbytes = st["pair_evals_total"] * 48.0
    gbs = nbytes / (st["kernel_ms"] * 1e-3) / 1e9
    print(json.dumps({"case": label, "natoms": w.natoms, "frames": nframes, "window_bytes": nframes * w.natoms * 24, "lags": lmax,
                      "origins": nts, "displacements": st["pair_evals_total"], "kernel_ms": st["kernel_ms"],
                      "algorithmic_gbs": gbs, "hbm_peak_gbs": peak}), file=sys.stderr)
    print("%s: %d atoms, window %d frames (%.0f MB), %d lags x %d origins: %.3e displacements in %.2f ms = %.0f GB/s algorithmic%s, "
          "%.1f %% of the FP64 rate (9 operations per displacement, %.3g DFMA/s); msd(lag %d) = %.17g"
          % (label, w.natoms, nframes, nframes * w.natoms * 24 / 1e6, lmax, nts, st["pair_evals_total"], st["kernel_ms"], gbs,
             " = %.0f %% of the measured HBM peak %.0f GB/s" % (100 * gbs / peak, peak) if peak else "",
             100 * st["pair_evals_total"] * 9 / (st["kernel_ms"] * 1e-3) / fp64, fp64, lmax - 1, v[-1, 0, 0]))
    tr.close()
ctx.close()
