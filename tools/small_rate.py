#!/usr/bin/env python
"""Throughput of the two pair kernels on SMALL systems (56 .. 512 atoms, the sizes of ab-initio cells): the
small-system kernel (one job per group of ceil(slots/64) warps, the default up to 256 device slots) against the tile
kernel (AGOFRT_OPT_NO_SMALL), pair evaluations per second from the library's own CUDA-event timing, one JSON
line per size.  python tools/small_rate.py [sizes...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from analisi_b200 import cabi, synth  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [56, 128, 192, 256, 384, 512]
LAGS, ORIGINS = 64, 128
ctx = cabi.Context([0])
peak = ctx.fp64_peak(0.3)
for n in sizes:
    pos, box, _ = synth.small_case(1000 + n, (8, 8, 8), 1.1, 1, False, LAGS + ORIGINS)
    pick = np.sort(np.random.default_rng(n).choice(512, n, replace=False))   # uniform density
    pos = np.ascontiguousarray(pos[:, pick])
    box = box * (n / 512.0) ** (1.0 / 3.0)        # same density
    pos = pos * (n / 512.0) ** (1.0 / 3.0)
    types = (np.arange(n) % 2).astype(np.int32)
    bi = synth.lammps_rows_to_internal(box)
    ctx.pbc_wrap(pos, bi)
    tr = cabi.DeviceTrajectory(ctx, n, 6, types, 2, pos.shape[0])
    tr.upload(0, pos, bi)
    plan = cabi.Plan(tr, 0.0, 0.45 * float(2 * bi[0, 3]), 200)
    out = {"natoms": n, "jobs": LAGS * ORIGINS, "pair_evals": LAGS * ORIGINS * n * n}
    ref = None
    for name, opt in (("small_kernel", cabi.OPT_SMALL), ("tile_kernel", cabi.OPT_NO_SMALL)):
        best = None
        for k in range(4):
            c, st = plan.block(0, ORIGINS, LAGS, 1, 1, options=opt)
            if k:
                best = st["kernel_ms"] if best is None else min(best, st["kernel_ms"])
        assert bool(st["kernel_modes"] & cabi.MODE_BIT_SMALL) == (opt == cabi.OPT_SMALL)
        if ref is None:
            ref = c
        assert np.array_equal(c, ref), "the two kernels disagree"
        rate = st["pair_evals_total"] / (best * 1e-3)
        out[name] = {"kernel_ms": best, "pair_evals_per_s": rate, "frac_of_16op_fp64_roofline": rate * 16 / peak}
    out["speedup"] = out["tile_kernel"]["kernel_ms"] / out["small_kernel"]["kernel_ms"]
    print(json.dumps(out), flush=True)
    plan.close()
    tr.close()
ctx.close()
