#!/usr/bin/env python
"""Throughput of the neighbour-count kernel (agofrt_neighbour_hist) on the synthetic workload shapes: pair
evaluations per second from the library's own CUDA-event timing.  python tools/neighbour_rate.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from analisi_b200 import cabi, synth  # noqa: E402

ctx = cabi.Context([0])
peak = ctx.fp64_peak(0.5)
for name, nframes, r in (("C2", 64, 3.0), ("C4", 2, 3.0), ("C3", 8, 4.0)):
    w = synth.WORKLOADS[name]
    pos, box, types = synth.generate(w, nframes=nframes)
    bi = synth.lammps_rows_to_internal(box)
    pos = np.ascontiguousarray(pos)
    ctx.pbc_wrap(pos, bi)
    tr = cabi.DeviceTrajectory(ctx, w.natoms, bi.shape[1], types, w.ntypes, nframes)
    tr.upload(0, pos, bi)
    tr.neighbour_hist(r, 0, nframes, 1)
    h, st = tr.neighbour_hist(r, 0, nframes, 1)
    rate = st["pair_evals_total"] / (st["kernel_ms"] * 1e-3)
    ops = 20 if w.triclinic else 17   # the pair kernel's 19 / 16 + the r2 compare, which is FP64 here
    print("%s: %d atoms x %d frames, %.3e pair evals in %.2f ms = %.3e /s = %.1f %% of the %d-op FP64 roofline (mean neighbours %.1f)"
          % (name, w.natoms, nframes, st["pair_evals_total"], st["kernel_ms"], rate, 100 * rate * ops / peak, ops,
             float((h * np.arange(h.shape[1])).sum() / max(1, h.sum()) * w.ntypes)))
    tr.close()
ctx.close()
