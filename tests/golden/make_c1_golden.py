"""Golden of BASELINE.json configs[0] produced by the UNMODIFIED reference (authoring container only).

    python tests/golden/make_c1_golden.py        (about 13 minutes on one core)

`analisi -i tests/data/lammps.bin -g 200 -F 0.7 3.5` = BlockAverage<Gofrt> over 20 blocks of 378 steps x 378 lags
(reference analisi/main.cpp:552-585).  oracle/_ref's `block_average_gofrt` runs exactly that chain of the compiled
reference (Trajectory -> set_pbc_wrap(true) -> BlockAverage<Gofrt<double,Trajectory>,...>::calculate) and returns its
mean and variance-of-the-mean arrays [378][6][200].

Stored (tests/golden/c1_reference.npz + .json):
  * S1 = sum over blocks of the integer counts, S2 = sum of their squares, per element -- recovered from the
    reference's own mean and variance (mean = S1 / (20*378), var = (S2 - S1^2/20) / (380 * 378^2)): they come out
    integers to 6e-9 / 2.4e-7, which is itself a check, and they reproduce the reference's doubles to 6e-14 (mean)
    and 1.2e-12 (variance) relative.  Two small integer arrays instead of 7 MB of doubles;
  * the reference's doubles themselves for lags 0, 1, 2, 100, 377 (full precision pins);
  * sha256 of the raw mean / variance arrays and of the CLI text the reference's printing loop
    (analisi/main.cpp:564-584) makes of them.
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
import oracle  # noqa: E402

N_B, S, NBIN, RMIN, RMAX = 20, 378, 200, 0.7, 3.5
PIN_LAGS = [0, 1, 2, 100, 377]


def cli_text(mean, var, desc):
    """The reference's printing loop (analisi/main.cpp:564-584): header, then `t r mean var ...` rows, a blank
    line after every lag; C++ ostream default formatting = %g with 6 significant digits."""
    out = [desc.rstrip("\n")]
    leff, ncol, nbin = mean.shape
    for t in range(leff):
        for r in range(nbin):
            row = ["%d" % t, "%d" % r]
            for c in range(ncol):
                row.append("%g" % mean[t, c, r])
                row.append("%g" % var[t, c, r])
            out.append(" ".join(row))
        out.append("")
    return "\n".join(out) + "\n"


def main():
    cache = os.environ.get("C1_REF_CACHE")   # directory with mean.npy / var.npy / desc.txt of an earlier run
    if cache and os.path.exists(os.path.join(cache, "mean.npy")):
        mean, var = np.load(os.path.join(cache, "mean.npy")), np.load(os.path.join(cache, "var.npy"))
        desc = open(os.path.join(cache, "desc.txt")).read()
        dt = float("nan")
    else:
        ref = oracle.load_ref()
        assert ref is not None, "oracle/_ref is not built"
        t0 = time.time()
        mean, var, desc = ref.block_average_gofrt("/root/reference/tests/data/lammps.bin", N_B, RMIN, RMAX, NBIN, 0, 1, 1, 1,
                                                  False, True)
        mean, var = np.asarray(mean), np.asarray(var)
        dt = time.time() - t0
    assert mean.shape == (S, 6, NBIN)
    s1 = mean * (N_B * S)
    S1 = np.round(s1)
    s2 = var * (N_B * (N_B - 1)) * S ** 2 + S1 * S1 / N_B
    S2 = np.round(s2)
    res1, res2 = float(np.abs(s1 - S1).max()), float(np.abs(s2 - S2).max())
    assert res1 < 1e-7 and res2 < 1e-5, (res1, res2)
    np.savez_compressed(os.path.join(HERE, "c1_reference.npz"), S1=S1.astype(np.uint32), S2=S2.astype(np.uint64),
                        pin_lags=np.array(PIN_LAGS), pin_mean=mean[PIN_LAGS], pin_var=var[PIN_LAGS])
    text = cli_text(mean, var, desc)
    meta = {"command": "analisi -i tests/data/lammps.bin -g 200 -F 0.7 3.5  (20 blocks x 378 steps x 378 lags, 56 atoms)",
            "source": "oracle/_ref analisi_ref.block_average_gofrt (unmodified reference), 1 thread, %.0f s" % dt,
            "shape": list(mean.shape), "mean_sha256": hashlib.sha256(np.ascontiguousarray(mean).tobytes()).hexdigest(),
            "var_sha256": hashlib.sha256(np.ascontiguousarray(var).tobytes()).hexdigest(),
            "cli_text_sha256": hashlib.sha256(text.encode()).hexdigest(), "cli_text_bytes": len(text),
            "integer_residual_S1": res1, "integer_residual_S2": res2, "columns_description": desc}
    json.dump(meta, open(os.path.join(HERE, "c1_reference.json"), "w"), indent=1)
    print(json.dumps({k: meta[k] for k in ("source", "mean_sha256", "cli_text_sha256", "integer_residual_S1", "integer_residual_S2")}))


if __name__ == "__main__":
    main()
