"""Golden counts of the DEFAULT bench workload (C4 subset), produced by the UNMODIFIED reference.

    python tests/golden/make_c4_subset_golden.py [threads]        (authoring container only: needs oracle/_ref)

BASELINE.json configs[3] (100 000-atom triclinic liquid, 500 bins, lags 0-200) is 1.29e14 pair evaluations per
block -- days of CPU.  bench.py's default step is a SUBSET of that block on the same atoms, frames, cell and
bins: every 8th lag (0, 8, ..., 200: 26 lags) x every 96th origin (0, 96, ..., 672: 8 origins) = 208
(lag, origin) jobs of 1e10 pair evaluations (synth.BENCH_SUBSET["C4"]).  This script runs exactly that through
the compiled reference (analisi_ref.Trajectory(wrap=True) + Gofrt.reset(768).calculate(0), about 1.5 h on 7
threads), turns the float result back into integer counts (incr = 1/8 is a power of two, so vdata/incr is
exact), and stores the non-empty lag rows plus the sha256 of the full [201][2][500] uint64 array that bench.py
prints as ``counts_sha256`` at every GPU count.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from analisi_b200 import synth  # noqa: E402
import oracle  # noqa: E402


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    ref = oracle.load_ref()
    assert ref is not None, "oracle/_ref is not built"
    w, nts = synth.bench_subset("C4")
    leff = min(nts, w.tmax)
    # the loops touch frames 0 .. 872, but the reference's length check (lib/src/gofrt.cpp:81-83) wants
    # leff + ntimesteps - 1 frames in the trajectory
    nframes = nts + leff - 1
    t0 = time.time()
    pos, box, types = synth.generate(w, nframes=nframes)
    print("generated %d frames in %.0f s" % (nframes, time.time() - t0), flush=True)
    tr = ref.Trajectory(pos, np.zeros_like(pos), types, box, ref.BoxFormat.LammpsTriclinic, True, False)
    g = ref.Gofrt(tr, w.rmin, w.rmax, w.nbin, w.tmax, threads, w.skip, w.every, False)
    g.reset(nts)
    t0 = time.time()
    g.calculate(0)
    dt = time.time() - t0
    v = np.array(g, copy=True)
    incr = 1.0 / (nts // w.skip)
    c = v / incr
    assert np.array_equal(c, np.round(c)), "non-integer counts"
    counts = c.astype(np.uint64)
    assert counts.shape == (leff, 2, w.nbin)
    sha = hashlib.sha256(np.ascontiguousarray(counts).astype("<u8").tobytes()).hexdigest()
    lags = np.arange(0, leff, w.every)
    njobs = len(lags) * ((nts + w.skip - 1) // w.skip)
    np.savez_compressed(os.path.join(HERE, "c4_subset_counts.npz"), lags=lags, counts=counts[lags])
    meta = {"workload": w.name, "subset": {"ntimesteps": nts, "skip": w.skip, "every": w.every, "leff": leff},
            "frames": nframes, "jobs": int(njobs), "pair_evals": float(njobs) * w.natoms ** 2,
            "counts_sha256": sha, "counts_sum": int(counts.sum()),
            "source": "unmodified reference (oracle/_ref analisi_ref.Gofrt), %d threads, %.0f s" % (threads, dt)}
    json.dump(meta, open(os.path.join(HERE, "c4_subset_counts.json"), "w"), indent=1)
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
