"""CPU, world size 2 on the gloo backend: the N>1 plumbing of bench.py / analisi_b200.dist and the shard
geometry of agofrt_block (agofrt_shard_range), without GPUs.

What runs on the GPUs in the product -- the pair kernels on each rank's share of the work units and the
NCCL all-reduce of the integer histograms -- is played here by the oracle on each rank's share of the
(lag, origin) jobs and a gloo all-reduce: the sum of the partial histograms must be the whole block,
bit for bit, which is the property that makes the multi-GPU result independent of the GPU count."""
import json
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def torchrun(nproc, script_args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(free_port())] + script_args
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), OMP_NUM_THREADS="2")
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env, cwd=ROOT)


WORKER = textwrap.dedent('''
    import json, sys
    import numpy as np
    import oracle
    from analisi_b200 import cabi, dist, synth

    ranks = dist.Ranks(backend="gloo")
    assert ranks.world == 2
    # 1. the NCCL-id broadcast path (128 opaque bytes made on rank 0)
    payload = bytes(range(128)) if ranks.rank == 0 else b""
    got = ranks.broadcast_bytes(payload, 128, 0)
    assert got == bytes(range(128))
    # 2. max over ranks of a device time
    assert ranks.max_over_ranks(10.0 + ranks.rank) == 11.0
    # 3. shard geometry: the (lag, origin) jobs of a block, cut like agofrt_block cuts its work units
    pos, box, types = synth.small_case(41, (4, 3, 3), 1.1, 2, True, 14)
    bi = synth.lammps_rows_to_internal(box)
    pos = oracle.pbc_wrap(pos, bi)
    rmin, rmax, nbin, lmax, nts, skip = 0.0, 2.4, 24, 4, 9, 2
    origins = list(range(0, nts, skip))
    jobs = [(t, o) for t in range(lmax) for o in origins]          # the reference's loop order
    b, e = cabi.shard_range(len(jobs), ranks.rank, ranks.world)
    part = np.zeros((lmax, 6, nbin), dtype=np.uint64)
    for (t, o) in jobs[b:e]:
        # one job = lag t at origin o: a block whose only origin is o (skip = ntimesteps), lags 0..t; keep row t
        c = oracle.counts(pos, bi, types, rmin, rmax, nbin, t + 1, t + 1, primo=o, skip=t + 1, ntypes=2, total_frames=100)
        part[t] += c[t]
    total = ranks.sum_counts(part)
    full = oracle.counts(pos, bi, types, rmin, rmax, nbin, lmax, nts, primo=0, skip=skip, ntypes=2)
    ranks.barrier()
    if ranks.rank == 0:
        print(json.dumps({"equal": bool(np.array_equal(total, full)), "sum": int(full.sum()),
                          "my_jobs": e - b, "jobs": len(jobs)}))
    ranks.close()
''')


def test_world2_shards_sum_to_the_whole_block(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = torchrun(2, [str(script)])
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["equal"] and line["sum"] > 0
    assert line["my_jobs"] * 2 in (line["jobs"], line["jobs"] + 1, line["jobs"] - 1)


def test_reference_arm_under_torchrun_prints_once():
    """bench.py --impl reference with two ranks: rank 0 alone times the CPU path and prints ONE JSON line,
    rank 1 exits 0 without work (contract of the reference arm)."""
    r = torchrun(2, ["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1",
                     "--workload", "C2"])   # (the default workload, C4, is one 1e10-pair job per step: minutes on these cores)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["unit"] == "pair_evals/s" and d["config"]["workload"].startswith("C2")
