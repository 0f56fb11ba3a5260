#!/usr/bin/env python
"""bench.py -- g(r,t) pair-distance evaluations per second on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4] [--impl native|reference]

Default step: the north-star workload C4 (100 000 atoms, triclinic cell), a regular subset of its (lag, origin) jobs
(synth.BENCH_SUBSET: every 8th lag, every 96th origin -- 2.08e12 pair evaluations, 2.3 s on one B200).  Its counts are
compared with the counts of the UNMODIFIED reference on the same input (tests/golden/c4_subset_counts.json, sha256).
A "step" is one Gofrt block: ``reset(ntimesteps); calculate(primo)``, i.e. ceil(leff/every) * ceil(ntimesteps/skip)
(lag, origin) jobs of N^2 pair evaluations each.

``value``   pair evaluations / s with the trajectory window already resident in HBM, through the C ABI of
            libagofrt.so, timed with CUDA events on the library's own stream (agofrt_stats.total_ms), max over ranks.
``e2e``     the same through the reference-facing python module of this repository (python/pyanalisi*.so), every step:
            ``Trajectory(pos, vel, types, box, fmt, wrap=True)`` (box conversion, type compaction, upload from page-locked
            numpy arrays -- with N ranks every rank uploads 1/N of the frames and the shares are exchanged GPU to GPU --,
            wrap on the GPUs), ``Gofrt(...)``, ``reset``, ``calculate``, ``np.array(g)``; wall clock between barriers.
            ``e2e.from_pageable_arrays`` is one more step from ordinary (pageable) numpy arrays.
``roofline`` FP64-pipe bound: 16 (orthorhombic) / 19 (triclinic) FP64 operations per pair
            evaluation against the DFMA issue rate measured in this same run; also the fraction counted on the FP64
            instructions the kernel really issues (14 / 17), and the DRAM traffic of one launch, measured by re-running
            the step once under ncu at the end of the run (nothing of that child is timed).
``cpu_baseline`` the reference's own CPU implementation (oracle/_ref, compiled from the unmodified sources) -- or the
            oracle port when that is absent -- on a bounded sample of the same workload on this box's host cores.
``counts_sha256`` / ``counts_check`` the checksum of the step's integer counts (identical for every N), the reference's
            checksum, and the lag-0 self-pair property.

N > 1: one process per GPU (torchrun); every rank holds the whole window, the work units of the block are sharded over
the ranks, one NCCL all-reduce of the integer histograms per step ("strong" scaling: the total work per step is fixed).
No PyTorch in the data path: torch is only used for the multi-process rendezvous, the barrier and the NCCL-id broadcast.
"""
import argparse
import dataclasses
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from analisi_b200 import synth  # noqa: E402

METRIC = "g(r,t) pair-distance evals/sec"
UNIT = "pair_evals/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------
def block_spec(name, quick=False, full=False):
    """(workload, ntimesteps, primo) of the step: SURVEY.md section 8(d)."""
    w = synth.WORKLOADS[name]
    if name in synth.BENCH_SUBSET and not quick and not full:
        # the default step of a workload whose full block takes minutes: a regular subset of its (lag, origin) jobs
        w, nts = synth.bench_subset(name)
        return w, nts, 0
    if name == "C2":
        # throughput run: reset(1899); calculate(0), skip 1, lags 0..100 -> 101*1899 jobs of 4096^2
        nts = 1899 if not quick else 64
        return w, nts, 0
    if name == "C3":
        if quick:
            w = dataclasses.replace(w, tmax=4, skip=15)
            return w, 120, 0
        return w, 960, 0
    if name == "C4":
        if quick:
            w = dataclasses.replace(w, tmax=4, skip=12)
            return w, 48, 0
        return w, 768, 0
    raise SystemExit("workload %s is not a single-block bench workload" % name)


def run_c5(args, rank, world, name="C5"):
    """BASELINE.json configs[4]: 1M-atom 2-type melt, 200 frames, block-averaged with variance (MediaBlocchi,
    8 blocks) -- the whole reference-facing chain in one process: LAMMPS binary file -> mmap Trajectory (window
    reads, wrap on the GPU) -> BlockAverageG<Gofrt> -> mean and variance, on the first ``--gpus`` devices of
    the box (work units of every block sharded over them, one NCCL all-reduce per block).  The host C++ classes
    own the devices, so under torchrun rank 0 drives all N GPUs and the other ranks exit."""
    if rank != 0:
        return 0
    import tempfile
    bundled = None
    if name == "C1":
        # BASELINE.json configs[0]: analisi -i tests/data/lammps.bin -g 200 -F 0.7 3.5 (56 atoms, 2 types, 7958
        # frames, -B 20 -S 0 -s 1: 20 blocks of 378 steps x 378 lags = 8.96e9 pair evaluations).  The file is the
        # reference's own test trajectory, copied to tests/_refdata by build() where the reference tree exists.
        bundled = os.path.join(ROOT, "tests", "_refdata", "lammps.bin")
        if not os.path.exists(bundled):
            raise SystemExit("tests/_refdata/lammps.bin is not here (built where /root/reference exists)")
        w = synth.Workload("C1 analisi -i tests/data/lammps.bin -g 200 -F 0.7 3.5 (bundled 56-atom trajectory, 20 blocks)",
                           0, (56, 1, 1), 1.0, 2, "blocks", False, 7958, 0.7, 3.5, 200, 0, 1, 1, 20)
    else:
        w = synth.WORKLOADS["C5"]
        if args.quick:
            w = dataclasses.replace(w, cells=(40, 40, 40), name=w.name + " [quick: 64k atoms]")
    os.environ["ANALISI_DEVICES"] = ",".join(str(i) for i in range(max(1, args.gpus)))
    from analisi_b200 import build as b
    _, ext = b.build_host()
    sys.path.insert(0, os.path.dirname(ext))
    import pyanalisi as pa
    tmpdir = os.environ.get("AGOFRT_TMP", tempfile.gettempdir())
    path = os.path.join(tmpdir, "agofrt_c5_%d_%d.bin" % (w.natoms, os.getpid()))
    t0 = time.time()
    if bundled:
        path = bundled
    else:
        nbytes = synth.write_workload_lammps(path, w)
        log("[bench] wrote %s: %.2f GB in %.1f s" % (path, nbytes / 1e9, time.time() - t0))
    try:
        def one_pass():
            tr = pa.Traj(path)
            tr.setLoadVelocities(False)
            tr.setWrapPbc(True)
            ba = pa.GofrtBlockAverage_lammps(tr, w.nblocks)
            t0 = time.time()
            ba.calculate(w.rmin, w.rmax, w.nbin, w.tmax, 1, w.skip, w.every, False)
            return ba, time.time() - t0
        sampler = ClockSampler(0)
        walls, devs, kers, st = [], [], [], None
        mean0 = None
        for k in range(args.warmup + args.steps):
            if k == args.warmup:
                sampler.start()
                tt0 = time.time()
            ba, wall = one_pass()
            st = ba.stats()
            log("[bench] pass %d: wall %.2f s, device %.2f s, kernels %.2f s, %d blocks on %d GPU(s)"
                % (k, wall, st["device_ms"] / 1e3, st["kernel_ms"] / 1e3, st["blocks"], st["ndev"]))
            if k >= args.warmup:
                walls.append(wall)
                devs.append(st["device_ms"] / 1e3)
                kers.append(st["kernel_ms"] / 1e3)
            mean, var = ba.mean(), ba.variance()
            if mean0 is not None and not (np.array_equal(mean, mean0)):
                raise SystemExit("non-deterministic block averages between passes")
            mean0 = mean
        clocks = sampler.stop(tt0, time.time())
        # size-independent checks: every atom meets itself at distance 0 once per origin -> self row, bin 0 =
        # atoms of the type, in every block (variance 0); nothing else in the self rows at lag 0
        nt, P = w.ntypes, w.ntypes * (w.ntypes + 1) // 2
        per_type = np.bincount(synth.lattice_types(w), minlength=nt)
        for a in range(nt if not bundled else 0):   # (C1 starts at rmin = 0.7: no self pairs at lag 0)
            slot = P - (a + 1) * (a + 2) // 2 + a + P
            assert mean[0, slot, 0] == per_type[a] and var[0, slot, 0] == 0.0, (a, mean[0, slot, 0])
            assert mean[0, slot, 1:].sum() == 0.0
        pairs = float(st["pair_evals"])
        ops = 16
        peak = None
        try:
            from analisi_b200 import cabi
            c = cabi.Context([0])
            peak = c.fp64_peak(1.0)
            c.close()
        except Exception as e:
            log("[bench] fp64 peak not measured: %r" % (e,))
        ngpu = int(st["ndev"])
        value = pairs / float(np.mean(devs))
        e2e = pairs / float(np.mean(walls))
        kernel_rate = pairs / float(np.mean(kers)) / ngpu
        s_blk = ba.block_size()
        # block averages on the device (MediaVarDevice) unless ANALISI_DEVICE_BLOCKS=0: one more kernel per block,
        # and only mean, variance and the last block come back instead of every block's counts
        dev_blocks = os.environ.get("ANALISI_DEVICE_BLOCKS", "1") != "0"
        a_ = w.nframes // (w.nblocks + 1) + 1
        win = s_blk + (a_ if (a_ < w.tmax or w.tmax == 0) else w.tmax)   # frames of one block window (Gofrt::nExtraTimesteps)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ngpu, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(devs)) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic" if not bundled else "reference test trajectory",
            "config": {"workload": w.name, "natoms": w.natoms, "frames": w.nframes, "ntypes": w.ntypes, "triclinic": False,
                       "rmin": w.rmin, "rmax": w.rmax, "nbin": w.nbin, "lags": w.tmax, "skip": w.skip, "blocks": w.nblocks,
                       "block_size": int(s_blk), "pair_evals_per_step": pairs,
                       "parallelism": "one process, %d GPU(s): work units of each block sharded, 1 NCCL all-reduce/block; "
                                      "MediaVar %s in block order" % (ngpu, "on the device" if dev_blocks else "on the host"),
                       "step": "file -> %d block windows -> mean and variance (the analisi -g ... -B %d chain)" % (w.nblocks, w.nblocks),
                       "l2": "block window: %.1f MB%s" % (win * w.natoms * 24 / 1e6, " (larger than L2)" if win * w.natoms * 24 > 126e6 else " (fits in L2; 56 atoms: one (lag, origin) job per warp, small-system kernel)")},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(w.nblocks * win * w.natoms * 24 * 2),
                    "d2h_bytes_per_step": int(w.nblocks * win * w.natoms * 24 + (3 if dev_blocks else w.nblocks) * mean.size * 8),
                    "ms_per_step": float(np.mean(walls)) * 1e3,
                    "includes": "mmap read + id scatter (host threads), wrap round trip, window upload, kernels, read-back, Welford"},
            "gpu_launches": int(st["blocks"]) * ngpu + (int(st["blocks"]) if dev_blocks else 0), "clocks": clocks,
            "roofline": {"bound": "fp64", "achieved": kernel_rate * ops / 1e12, "peak": peak / 1e12 if peak else None,
                         "unit": "T FP64-op/s per GPU", "frac": kernel_rate * ops / peak if peak else None, "traffic": None,
                         "ops_per_pair_eval": ops, "kernel_ms_per_step": float(np.mean(kers)) * 1e3,
                         "pair_evals_per_s_per_gpu": kernel_rate},
        }
        print(json.dumps(line), flush=True)
    finally:
        if not bundled:
            try:
                os.remove(path)
            except OSError:
                pass
    return 0


def load_pyanalisi(local_rank):
    """The pybind11 module (host C++ mirror of the reference API over the C ABI), bound to this rank's GPU."""
    os.environ.setdefault("ANALISI_DEVICES", str(local_rank))
    from analisi_b200 import build as b
    _, ext = b.build_host()
    d = os.path.dirname(ext)
    if d not in sys.path:
        sys.path.insert(0, d)
    import pyanalisi
    return pyanalisi


def jobs_of(nts, leff, skip, every):
    return ((leff + every - 1) // every) * ((nts + skip - 1) // skip)


def make_window(w, nframes):
    t0 = time.time()
    pos, box_lammps, types = synth.generate(w, nframes=nframes)
    box_internal = synth.lammps_rows_to_internal(box_lammps)
    log("[bench] generated %s: %d atoms x %d frames in %.1f s" % (w.name, w.natoms, nframes, time.time() - t0))
    return pos, box_lammps, box_internal, types


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.dev = device_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.3 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[5 + k].lower().startswith("active") for r in rows)]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------
# the reference's CPU implementation on a bounded sample
# ---------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample(w, pos, box_lammps, box_internal, types, target_s=12.0, threads=None):
    """Time the reference CPU path on the first frames of the same window.  Returns a dict."""
    import oracle
    threads = threads or host_threads()
    ref = oracle.load_ref()
    n2 = float(w.natoms) ** 2
    fmt_tri = w.triclinic

    def run(lags, origins):
        nfr = origins + lags
        p = np.ascontiguousarray(pos[:nfr])
        if ref is not None:
            fmt = ref.BoxFormat.LammpsTriclinic if fmt_tri else ref.BoxFormat.LammpsOrtho
            tr = ref.Trajectory(p, np.zeros_like(p), types, np.ascontiguousarray(box_lammps[:nfr]), fmt, True, False)
            g = ref.Gofrt(tr, w.rmin, w.rmax, w.nbin, lags, threads, 1, 1, False)
            g.reset(origins)
            t0 = time.perf_counter()
            g.calculate(0)
            dt = time.perf_counter() - t0
        else:
            pw = oracle.pbc_wrap(p, box_internal[:nfr])
            t0 = time.perf_counter()
            oracle.counts(pw, box_internal[:nfr], types, w.rmin, w.rmax, w.nbin, lags, origins, nthreads=threads,
                          ntypes=w.ntypes)
            dt = time.perf_counter() - t0
        return lags * origins * n2 / dt, dt

    # calibrate on one lag x two origins, then size the sample for ~target_s; with 1e10 pairs per (lag, origin)
    # job (C4) one job IS the sample
    if n2 >= 2.5e9:
        rate, dt = run(1, 1)
        return {
            "value": rate, "unit": UNIT, "cores": threads, "kind": "reference" if ref is not None else "port",
            "sample": "%s: frames 0-1, lag 0 x 1 origin = 1 (lag,origin) job of %d^2 pairs, %.1f s, %d threads"
                      % (w.name.split()[0], w.natoms, dt, threads),
        }
    rate, dt = run(1, 2)
    jobs = int(max(2, min(4096, target_s * rate / n2)))
    lags = max(1, min(4, jobs // 2, pos.shape[0] // 2))
    origins = max(1, min(jobs // lags, pos.shape[0] - lags))
    rate, dt = run(lags, origins)
    return {
        "value": rate, "unit": UNIT, "cores": threads,
        "kind": "reference" if ref is not None else "port",
        "sample": "%s: first %d frames, lags 0-%d x %d origins = %d (lag,origin) jobs of %d^2 pairs, %.1f s, %d threads"
                  % (w.name.split()[0], lags + origins, lags - 1, origins, lags * origins, w.natoms, dt, threads),
    }


# ---------------------------------------------------------------------------------------------
# DRAM traffic of the step's kernel, measured in this run: the same block once more under ncu
# ---------------------------------------------------------------------------------------------
# FP64 instructions the pair kernels ISSUE per pair evaluation (SASS of the shipped library,
# profiles/r2_sass_counts.txt): the 16 / 19 algorithmic operations of SURVEY.md section 8(d) minus the two range
# compares, which ride on the integer high word of d2.
FP64_ISSUED = {False: 14, True: 17}


def measure_traffic(args, name):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the step's pair kernel: bench.py re-runs itself
    (``--traffic-child``: same workload, one block, nothing timed) under ``ncu``.  None when ncu is not usable."""
    import shutil
    import tempfile
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    log_path = os.path.join(tempfile.gettempdir(), "agofrt_traffic_%d.csv" % os.getpid())
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "--print-units", "base",
           "--csv", "--log-file", log_path, "-k", "regex:pair_", sys.executable, os.path.abspath(__file__),
           "--traffic-child", "--workload", name, "--options", str(args.options)]
    if args.quick:
        cmd.append("--quick")
    if args.full:
        cmd.append("--full")
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        if r.returncode != 0:
            return None, "ncu run failed (rc %d): %s" % (r.returncode, (r.stderr or r.stdout)[-200:].replace("\n", " "))
        import csv
        rd = wr = 0.0
        launches = set()
        with open(log_path) as f:
            rows = [row for row in csv.reader(f) if len(row) > 5]
        hdr = next(k for k, row in enumerate(rows) if "Metric Name" in row)
        col = {n: k for k, n in enumerate(rows[hdr])}
        for row in rows[hdr + 1:]:
            v = float(row[col["Metric Value"]].replace(",", ""))
            launches.add(row[col["ID"]])
            if row[col["Metric Name"]] == "dram__bytes_read.sum":
                rd += v
            elif row[col["Metric Name"]] == "dram__bytes_write.sum":
                wr += v
        if not launches:
            return None, "no pair kernel in the ncu log"
        return {"dram_bytes_read": rd / len(launches), "dram_bytes_write": wr / len(launches), "launches": len(launches)}, "ncu"
    except Exception as e:
        return None, "ncu: %r" % (e,)
    finally:
        try:
            os.remove(log_path)
        except OSError:
            pass


def golden_sha(name, w, nts):
    """The committed checksum of the step's counts from the unmodified reference, when this exact step has one."""
    path = os.path.join(ROOT, "tests", "golden", "%s_subset_counts.json" % name.lower())
    try:
        g = json.load(open(path))
        if g["subset"] == {"ntimesteps": nts, "skip": w.skip, "every": w.every, "leff": min(nts, w.tmax)} and g["workload"] == w.name:
            return g["counts_sha256"]
    except Exception:
        pass
    return None


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None,
                    help="C4 (default: BASELINE.json configs[3], the north-star shape; a regular subset of its (lag, origin) "
                         "jobs per step unless --full), C2, C3 (one Gofrt block per step); C1, C5 (block-averaged chain from a file)")
    ap.add_argument("--full", action="store_true", help="the whole block of the workload instead of the bench subset (C4: 150 s/step)")
    ap.add_argument("--quick", action="store_true", help="tiny block (smoke/profiling), not a bench number")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the ncu re-run that measures roofline.traffic")
    ap.add_argument("--no-e2e", action="store_true", help="kernel A/B runs with a tuning build of the library: skip the end-to-end leg")
    ap.add_argument("--traffic-child", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--options", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=None, help="size of the CPU sample (default 12 s; 8 s per step for --impl reference)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    name = args.workload or "C4"
    if name in ("C1", "C5"):
        return run_c5(args, rank, world, name)
    w, nts, primo = block_spec(name, args.quick, args.full)
    leff = min(nts, w.tmax) if w.tmax else nts
    nframes = primo + (nts - 1) // w.skip * w.skip + (leff - 1) // w.every * w.every + 1
    njobs = jobs_of(nts, leff, w.skip, w.every)
    pairs_per_step = njobs * w.natoms * w.natoms
    config = {
        "workload": w.name, "natoms": w.natoms, "frames_in_window": nframes, "ntypes": w.ntypes,
        "triclinic": w.triclinic, "rmin": w.rmin, "rmax": w.rmax, "nbin": w.nbin, "lags": leff, "skip": w.skip,
        "every": w.every, "origins": (nts + w.skip - 1) // w.skip, "jobs_per_step": njobs,
        "pair_evals_per_step": pairs_per_step, "parallelism": "work units sharded over %d GPU(s), 1 NCCL all-reduce/step" % world,
        "l2": "window (%.0f MB) larger than L2; every step re-reads it from HBM" % (nframes * w.natoms * 24 / 1e6),
    }

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        pos, box_lammps, box_internal, types = make_window(w, min(nframes, 80 if w.natoms < 50000 else 4))
        vals = []
        res = None
        for k in range(args.warmup + args.steps):
            res = cpu_sample(w, pos, box_lammps, box_internal, types, target_s=args.cpu_seconds or 8.0)
            if k >= args.warmup:
                vals.append(res["value"])
        v = float(np.mean(vals))
        res["value"] = v
        line = {
            "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "impl": "reference", "cpu_baseline": res,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ native arm
    from analisi_b200 import cabi
    from analisi_b200 import dist as adist
    ranks = adist.Ranks()
    ctx = cabi.Context([local_rank])
    adist.join_communicator(ranks, ctx)

    # frames the loops touch (nframes) and frames the reference's length check wants in the trajectory
    # (lib/src/gofrt.cpp:81-83: leff + ntimesteps - 1; the e2e leg goes through that check)
    nframes_traj = max(nframes, primo + nts + leff - 1)
    pos_all, box_lammps_all, box_internal_all, types = make_window(w, nframes_traj)
    pos, box_lammps, box_internal = pos_all[:nframes], box_lammps_all[:nframes], box_internal_all[:nframes]
    # wrap=True, as the reference callers do (analisi/main.cpp:558): the wrap runs on the GPU
    pinned = cabi.PinnedArray(pos.shape)
    pinned.array[...] = pos
    ctx.pbc_wrap(pinned.array, box_internal)
    hpos = pinned.array

    tr = cabi.DeviceTrajectory(ctx, w.natoms, box_internal.shape[1], types, w.ntypes, nframes)
    plan = cabi.Plan(tr, w.rmin, w.rmax, w.nbin)

    if args.traffic_child:
        # under ncu (measure_traffic): the step's block once, nothing timed, nothing printed
        tr.upload(0, hpos, box_internal)
        plan.block(primo, nts, leff, w.skip, w.every, options=args.options)
        return 0

    barrier, maxrank = ranks.barrier, ranks.max_over_ranks

    # roofline denominator, measured in this run
    peak = ctx.fp64_peak(1.0)
    log("[bench] rank %d: FP64 issue rate %.3e DFMA/s" % (rank, peak))

    tr.upload(0, hpos, box_internal)
    counts0 = None
    for k in range(args.warmup):
        counts0, st = plan.block(primo, nts, leff, w.skip, w.every, options=args.options)
        log("[bench] warmup %d: %.1f ms (kernel %.1f ms), fast jobs %d/%d" % (k, st["total_ms"], st["kernel_ms"], st["jobs_fast"], st["jobs"]))

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # ---- device-resident timing ----
    barrier()
    t0 = time.time()
    dev_ms = 0.0
    ker_ms = 0.0
    launches = 0
    for k in range(args.steps):
        counts, st = plan.block(primo, nts, leff, w.skip, w.every, options=args.options)
        dev_ms += st["total_ms"]
        ker_ms += st["kernel_ms"]
        launches += st["launches"]
    barrier()
    t1 = time.time()
    wall_ms = (t1 - t0) * 1e3
    dev_ms = maxrank(dev_ms)
    ker_ms = maxrank(ker_ms)
    clocks = sampler.stop(t0, t1)
    if counts0 is not None and not np.array_equal(counts, counts0):
        raise SystemExit("non-deterministic counts between steps")

    # ---- end to end: the reference-facing call sequence (pyanalisi.Trajectory + Gofrt, reference
    # pyanalisi/src/pyanalisi.cpp:65-82, :487-504) on the caller's UNWRAPPED numpy arrays, every step: construct the
    # trajectory (box conversion, type compaction, upload from pageable memory -- the frames are dealt to the ranks and
    # exchanged GPU to GPU --, wrap on the GPUs), construct Gofrt, reset, calculate, read the result array back.
    if args.no_e2e:
        e2e_ms, e2e_h2d, e2e_d2h, e2e_parts = float('nan'), 0, 0, {}
    else:
        pa = load_pyanalisi(local_rank)
        if world > 1:
            uid = pa.comm_unique_id() if rank == 0 else b""
            pa.comm_join(ranks.broadcast_bytes(uid, cabi.COMM_ID_BYTES, 0), rank, world)
        fmt = pa.BoxFormat.LammpsTriclinic if w.triclinic else pa.BoxFormat.LammpsOrtho
        vel = np.zeros_like(pos_all)   # the interface wants velocities; g(r,t) never reads them
        raw_types = np.ascontiguousarray(types, dtype=np.int32)

        # the caller's arrays: page-locked (what the contract of this bench asks for: "the host->device copy of that
        # step's inputs from pinned host memory"), so the upload is one DMA per share; the same step from ordinary
        # pageable numpy arrays (staged through the library's pinned slots) is timed once and reported beside it
        pin_all = cabi.PinnedArray(pos_all.shape)
        pin_all.array[...] = pos_all
        e2e_parts, e2e_log = {}, []

        def e2e_step(src):
            t0 = time.perf_counter()
            tr_py = pa.Trajectory(src, vel, raw_types, box_lammps_all, fmt, True, False)
            t1 = time.perf_counter()
            g = pa.Gofrt(tr_py, w.rmin, w.rmax, w.nbin, w.tmax, 1, w.skip, w.every, False)
            g.reset(nts)
            t2 = time.perf_counter()
            g.calculate(primo)
            t3 = time.perf_counter()
            v = np.array(g, copy=True)
            st = g.last_stats()
            t3b = time.perf_counter()
            del g, tr_py
            t4 = time.perf_counter()
            e2e_log.append(round((t4 - t0) * 1e3, 1))
            log("[bench] rank %d e2e step: trajectory %.1f, Gofrt() + reset %.1f, calculate %.1f (device %.1f), result %.1f, "
                "teardown %.1f ms" % (rank, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, st["total_ms"],
                                      (t3b - t3) * 1e3, (t4 - t3b) * 1e3))
            e2e_parts.update(trajectory_ms=(t1 - t0) * 1e3, gofrt_ctor_reset_ms=(t2 - t1) * 1e3, calculate_ms=(t3 - t2) * 1e3,
                             calculate_device_ms=st["total_ms"], result_and_teardown_ms=(t4 - t3) * 1e3)
            return v

        v_e = e2e_step(pin_all.array)   # warm-up: module load, communicator, device allocations
        barrier()
        e0 = time.time()
        for k in range(args.steps):
            v_e = e2e_step(pin_all.array)
        barrier()
        e2e_ms = maxrank((time.time() - e0) * 1e3)
        e2e_parts_pinned = dict(e2e_parts)
        e2e_step(pos_all)            # pageable: first use of the staging slots
        barrier()
        e0 = time.time()
        v_p = e2e_step(pos_all)
        barrier()
        e2e_pageable_ms = maxrank((time.time() - e0) * 1e3)
        e2e_parts_pageable, e2e_parts = dict(e2e_parts), e2e_parts_pinned
        if not np.array_equal(v_p, v_e):
            raise SystemExit("e2e result from pageable arrays differs from the one from pinned arrays")
        pin_all.free()
        log("[bench] rank %d e2e steps (ms, the first is the warm-up): %s" % (rank, e2e_log))
        incr = cabi.gofrt_incr(nts, w.skip)
        if not np.array_equal(v_e, counts * incr):
            raise SystemExit("e2e result differs from the resident counts * incr")
        e2e_h2d = int(pos_all.nbytes // world + box_lammps_all.nbytes)   # every rank copies its share of the frames
        e2e_d2h = int(counts.nbytes)

    # ---- the counts themselves: one checksum that must not depend on the number of GPUs, and -- for the default
    # step -- must equal the checksum of the unmodified reference's counts (tests/golden/c4_subset_counts.json)
    import hashlib
    sha = hashlib.sha256(np.ascontiguousarray(counts).astype("<u8").tobytes()).hexdigest()
    gold = golden_sha(name, w, nts) if not (args.quick or args.full) else None
    # size-independent property: at lag 0 every atom meets itself at distance 0 once per origin (rmin = 0)
    P = w.ntypes * (w.ntypes + 1) // 2
    per_type = np.bincount(synth.lattice_types(w), minlength=w.ntypes)
    origins = (nts + w.skip - 1) // w.skip
    self_ok = True
    if w.rmin == 0.0:
        for a in range(w.ntypes):
            slot = P - (a + 1) * (a + 2) // 2 + a + P
            self_ok &= int(counts[0, slot, 0]) == int(per_type[a]) * origins and int(counts[0, slot, 1:].sum()) == 0
    if not self_ok:
        raise SystemExit("self row of lag 0 is wrong")
    if gold is not None and gold != sha:
        raise SystemExit("counts differ from the reference's (sha256 %s, reference %s)" % (sha, gold))

    value = args.steps * pairs_per_step / (dev_ms * 1e-3)
    e2e_value = args.steps * pairs_per_step / (e2e_ms * 1e-3)
    e2e_pageable = {}
    if not args.no_e2e:
        e2e_pageable = {"from_pageable_arrays": {"value": pairs_per_step / (e2e_pageable_ms * 1e-3), "ms_per_step": e2e_pageable_ms,
                                                  "breakdown_ms": {k: round(v, 2) for k, v in e2e_parts_pageable.items()},
                                                  "steps": 1}}
    ops = 19 if w.triclinic else 16
    issued = FP64_ISSUED[bool(w.triclinic)]
    kernel_rate = args.steps * pairs_per_step / (ker_ms * 1e-3) / world   # per GPU
    achieved = kernel_rate * ops
    in_range = float(counts.sum()) / float(pairs_per_step)
    traffic, traffic_extra = None, {}
    if rank == 0 and world == 1 and not args.no_traffic:
        # release the device window first: the child needs the same memory
        tj, how = measure_traffic(args, name)
        if tj is not None:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_extra = {"traffic_unit": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)",
                             "traffic_source": "ncu re-run of this step inside this bench run (%d launch(es))" % tj["launches"],
                             "algorithmic_bytes_per_launch": int(njobs) * 2 * int(w.natoms) * 24,
                             "hbm_gbs": traffic / (ker_ms / args.steps * 1e-3) / 1e9}
            try:
                mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                traffic_extra["hbm_peak_gbs"] = mp["hbm_gbs"]
                traffic_extra["hbm_frac"] = traffic_extra["hbm_gbs"] / mp["hbm_gbs"]
            except Exception:
                pass
        else:
            traffic_extra = {"traffic_source": "not measured: %s" % how}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": e2e_d2h,
                "ms_per_step": e2e_ms / args.steps, "last_step_breakdown_ms": {k: round(v, 2) for k, v in e2e_parts.items()},
                "path": "pyanalisi.Trajectory(pos, vel, types, box, fmt, wrap=True) + Gofrt(...).reset().calculate() + np.array(g), "
                        "from numpy arrays in page-locked host memory every step; H2D per rank = its share of the %d frames "
                        "(%d bytes in all), shares exchanged GPU to GPU" % (nframes_traj, pos_all.nbytes),
                **e2e_pageable},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "counts_sha256": sha,
        "counts_check": {"reference_sha256": gold, "equal_to_reference": (gold == sha) if gold else None,
                         "reference_source": "tests/golden/%s_subset_counts.json (unmodified reference, CPU)" % name.lower() if gold else None,
                         "self_row_lag0": bool(self_ok), "counts_sum": int(counts.sum())},
        "roofline": {
            "bound": "fp64", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "T FP64-op/s per GPU",
            "frac": achieved / peak, "traffic": traffic,
            "ops_per_pair_eval": ops, "fp64_instr_issued_per_pair_eval": issued,
            "fp64_pipe_issued_frac": kernel_rate * issued / peak,
            "kernel_ms_per_step": ker_ms / args.steps,
            "pair_evals_per_s_per_gpu": kernel_rate, "in_range_fraction": in_range,
            "peak_source": "DFMA-chain microbenchmark in this run (agofrt_fp64_peak); MEASURED_PEAKS.json has no FP64 entry",
            **traffic_extra,
        },
        "wall_ms_per_step": wall_ms / args.steps,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_sample(w, pos, box_lammps, box_internal, types, target_s=args.cpu_seconds or 12.0)
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "unavailable",
                                    "sample": "failed: %r" % (e,)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    plan.close()
    tr.close()
    ctx.close()
    ranks.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
