"""GPU, two or more devices: the multi-GPU path of agofrt_block -- work units sharded over the GPUs, ONE NCCL
all-reduce (ncclUint64, sum) of the integer histograms -- must return the very same integers as one GPU.
Skipped on a one-GPU box (the driver's -m gpu run); run with `gpurun --gpus 2`."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import oracle
from analisi_b200 import cabi, synth
from conftest import GOLDEN, ROOT
from test_multirank_gloo import torchrun

pytestmark = pytest.mark.gpu

needs2 = pytest.mark.skipif(cabi.device_count() < 2, reason="needs two GPUs")


def case():
    pos, box, types = synth.small_case(51, (10, 9, 8), 1.05, 2, True, 8)
    bi = synth.lammps_rows_to_internal(box)
    return np.ascontiguousarray(pos), bi, types


def block(ctx, pos, bi, types, options=0):
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], bi.shape[1], types, 2, pos.shape[0])
    tr.upload(0, pos, bi)
    plan = cabi.Plan(tr, 0.0, 3.1, 60)
    c, st = plan.block(1, 5, 3, 2, 1, options=options)
    plan.close()
    tr.close()
    return c, st


@needs2
def test_single_process_all_devices_equals_one_device():
    """a context with several local devices is its own communicator (what the CLI uses on a multi-GPU box)"""
    pos, bi, types = case()
    one = cabi.Context([0])
    one.pbc_wrap(pos, bi)
    c1, st1 = block(one, pos, bi, types)
    one.close()
    allc = cabi.Context("all")
    assert allc.ndev >= 2
    cn, stn = block(allc, pos, bi, types)
    allc.close()
    assert np.array_equal(cn, c1)
    assert stn["world"] == stn["ndev_local"] >= 2 and stn["pair_evals_total"] == st1["pair_evals_total"]
    assert np.array_equal(c1, oracle.counts(pos, bi, types, 0.0, 3.1, 60, 3, 5, primo=1, skip=2, ntypes=2))


@needs2
@pytest.mark.parametrize("wrap", [False, True])
def test_shared_upload_exchanges_the_frames_between_devices(wrap):
    """agofrt_traj_upload_ex(AGOFRT_UP_SHARED): every device copies (and wraps) only its share of the frames, the
    shares are exchanged device to device; afterwards every device holds the whole window -- the block, whose work
    units are spread over all devices, counts what one device counts."""
    pos, bi, types = case()              # pageable numpy memory: staged through page-locked slots
    wrapped = oracle.pbc_wrap(pos, bi)
    allc = cabi.Context("all")
    tr = cabi.DeviceTrajectory(allc, pos.shape[1], bi.shape[1], types, 2, pos.shape[0])
    if wrap:
        back = np.empty_like(pos)
        tr.upload_ex(0, pos, bi, wrap=True, shared=True, out=back)
        assert np.array_equal(back, wrapped)     # every device returned its share
    else:
        tr.upload_ex(0, wrapped, bi, shared=True)
    assert np.array_equal(tr.download(0, pos.shape[0]), wrapped)
    plan = cabi.Plan(tr, 0.0, 3.1, 60)
    c, st = plan.block(1, 5, 3, 2, 1)
    assert st["world"] == allc.ndev >= 2
    assert np.array_equal(c, oracle.counts(wrapped, bi, types, 0.0, 3.1, 60, 3, 5, primo=1, skip=2, ntypes=2))
    plan.close()
    tr.close()
    allc.close()


@needs2
def test_whole_blocks_are_dealt_to_the_devices():
    """agofrt_blocks on a context with several devices: block b runs on device b mod n, every device receives every
    block, the batch equals the blocks one device computes one by one."""
    pos, box, types = synth.small_case(61, (4, 4, 4), 1.05, 2, True, 60)
    bi = synth.lammps_rows_to_internal(box)
    pos = np.ascontiguousarray(pos)
    allc = cabi.Context("all")
    allc.pbc_wrap(pos, bi)
    tr = cabi.DeviceTrajectory(allc, pos.shape[1], 9, types, 2, pos.shape[0])
    tr.upload_ex(0, pos, bi, shared=True)
    plan = cabi.Plan(tr, 0.0, 2.5, 32)
    n_b, s, leff = 7, 8, 3
    st = plan.blocks(0, s, n_b, s, leff, 2, 1)
    assert st["world"] == allc.ndev >= 2
    blocks = []
    for b in range(n_b):
        ref = oracle.counts(pos, bi, types, 0.0, 2.5, 32, 3, s, primo=b * s, skip=2, ntypes=2)
        assert np.array_equal(plan.block_counts(b, leff), ref)
        blocks.append(ref * 0.25)
    acc = cabi.BlockAverage(allc)
    acc.begin(leff * 6 * 32)
    acc.push_blocks(plan, 0.25)
    mean, var = acc.end(n_b)
    omean, ovar = oracle.mediavar(np.array(blocks))
    assert np.array_equal(mean, omean.ravel()) and np.array_equal(var, ovar.ravel())
    acc.close(); plan.close(); tr.close(); allc.close()


@needs2
def test_neighbour_histogram_on_all_devices():
    """agofrt_neighbour_hist shards (frame, atom tile) units over the devices and all-reduces the histogram"""
    pos, bi, types = case()
    one = cabi.Context([0])
    one.pbc_wrap(pos, bi)
    ref = oracle.neighbour_hist(pos, bi, types, 2.6, 1, 6, 2, ntypes=2)
    for c in (one, cabi.Context("all")):
        tr = cabi.DeviceTrajectory(c, pos.shape[1], 9, types, 2, pos.shape[0])
        tr.upload(0, pos, bi)
        h, st = tr.neighbour_hist(2.6, 1, 6, 2)
        assert np.array_equal(h, ref)
        assert st["world"] == c.ndev
        tr.close()
        c.close()


WORKER = textwrap.dedent('''
    import json, os
    import numpy as np
    import oracle
    from analisi_b200 import cabi, dist, synth
    ranks = dist.Ranks()
    assert ranks.backend == "nccl"
    ctx = cabi.Context([ranks.local_rank])
    dist.join_communicator(ranks, ctx)
    pos, box, types = synth.small_case(51, (10, 9, 8), 1.05, 2, True, 8)
    bi = synth.lammps_rows_to_internal(box)
    pos = np.ascontiguousarray(pos)
    ctx.pbc_wrap(pos, bi)
    tr = cabi.DeviceTrajectory(ctx, pos.shape[1], 9, types, 2, pos.shape[0])
    tr.upload(0, pos, bi)
    plan = cabi.Plan(tr, 0.0, 3.1, 60)
    ok = True
    # the shared upload as a collective over the processes: every rank copies and wraps its share of the UNWRAPPED
    # frames, the shares are exchanged over NCCL; the window every rank ends up with is the wrapped one
    raw, _, _ = synth.small_case(51, (10, 9, 8), 1.05, 2, True, 8)
    tr2 = cabi.DeviceTrajectory(ctx, pos.shape[1], 9, types, 2, pos.shape[0])
    tr2.upload_ex(0, np.ascontiguousarray(raw), bi, wrap=True, shared=True)
    ok = ok and bool(np.array_equal(tr2.download(0, pos.shape[0]), pos))
    plan2 = cabi.Plan(tr2, 0.0, 3.1, 60)
    c2, _ = plan2.block(1, 5, 3, 2, 1)
    ok = ok and bool(np.array_equal(c2, oracle.counts(pos, bi, types, 0.0, 3.1, 60, 3, 5, primo=1, skip=2, ntypes=2)))
    plan2.close(); tr2.close()
    for opt in (0, cabi.OPT_FORCE_GENERAL, cabi.OPT_NO_SAFE):
        c, st, e = plan.block(1, 5, 3, 2, 1, options=opt, edges=True) if opt == 0 else plan.block(1, 5, 3, 2, 1, options=opt) + (None,)
        ref, eref = oracle.counts(pos, bi, types, 0.0, 3.1, 60, 3, 5, primo=1, skip=2, ntypes=2, return_edges=True)
        ok = ok and bool(np.array_equal(c, ref)) and st["world"] == ranks.world and (e is None or e == eref)
        ok = ok and 0 < st["pair_evals"] < st["pair_evals_total"]
    ranks.barrier()
    import sys
    sys.stdout.write("AGOFRT_RESULT " + json.dumps({"rank": ranks.rank, "ok": ok, "world": ranks.world}) + "\\n")   # one write: ranks must not interleave
    sys.stdout.flush()
    plan.close(); tr.close(); ctx.close(); ranks.close()
''')


@needs2
def test_one_process_per_gpu_nccl(tmp_path):
    """torchrun, one rank per GPU: unique id broadcast, agofrt_comm_join, every rank ends up with the whole
    (all-reduced) histogram, equal to the oracle's -- also the edge-pair counter."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = torchrun(2, [str(script)])
    assert r.returncode == 0, r.stderr[-3000:]
    import re
    lines = [json.loads(m) for m in re.findall(r"AGOFRT_RESULT (\{[^{}]*\})", r.stdout)]
    assert len(lines) == 2 and all(l["ok"] and l["world"] == 2 for l in lines), (r.stdout[-1500:], r.stderr[-1500:])


@needs2
def test_cli_on_all_gpus_matches_reference_golden(tmp_path):
    """the CLI on every GPU of the box: same text as the reference's golden (counts are integers, the
    block order of MediaVar does not depend on the GPU count)"""
    from analisi_b200 import build as b
    cli, _ = b.build_host()
    path = os.path.join(ROOT, "tests", "_refdata", "lammps2020.bin")
    if not os.path.exists(path):
        pytest.skip("tests/_refdata/lammps2020.bin not present")
    r = subprocess.run([cli, "-i", path, "-g", "100", "-F", "0.0", "4.0", "-S", "10", "-s", "8"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, cwd=str(tmp_path), timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    gold = open(os.path.join(GOLDEN, "cli_pair_corr_t.txt")).read()
    assert r.stdout.rstrip("\n") == gold.rstrip("\n")


PY_WORKER = textwrap.dedent('''
    import json, os, sys
    import numpy as np
    os.environ["ANALISI_DEVICES"] = os.environ.get("LOCAL_RANK", "0")
    import oracle
    from analisi_b200 import build as b, dist, synth
    _, ext = b.build_host()
    sys.path.insert(0, os.path.dirname(ext))
    import pyanalisi as pa
    ranks = dist.Ranks()
    uid = pa.comm_unique_id() if ranks.rank == 0 else b""
    pa.comm_join(ranks.broadcast_bytes(uid, 128, 0), ranks.rank, ranks.world)
    pos, box, types = synth.small_case(52, (9, 9, 8), 1.05, 2, True, 9)
    pos = np.ascontiguousarray(pos)
    tr = pa.Trajectory(pos, np.zeros_like(pos), types.astype(np.int32), box, pa.BoxFormat.LammpsTriclinic, True, False)
    g = pa.Gofrt(tr, 0.0, 3.0, 50, 3, 1, 2, 1, False)
    g.reset(6)
    g.calculate(1)
    bi = synth.lammps_rows_to_internal(box)
    wrapped = oracle.pbc_wrap(pos, bi)
    ref = oracle.counts(wrapped, bi, types, 0.0, 3.0, 50, 3, 6, primo=1, skip=2, ntypes=2)
    ok = bool(np.array_equal(g.counts(), ref)) and g.last_stats()["world"] == ranks.world
    ok = ok and bool(np.array_equal(tr.get_positions_copy(), wrapped))   # the host copy arrives on demand, wrapped
    ranks.barrier()
    sys.stdout.write("AGOFRT_RESULT " + json.dumps({"rank": ranks.rank, "ok": ok, "world": ranks.world}) + "\\n")
    sys.stdout.flush()
    ranks.close()
''')


@needs2
def test_pyanalisi_one_process_per_gpu(tmp_path):
    """The reference-facing classes as an SPMD program (torchrun, one rank per GPU): pyanalisi.comm_join, then
    Trajectory(..., wrap=True) uploads each rank's share of the frames and wraps it on its GPU, the shares are
    exchanged, Gofrt.calculate shards the block and all-reduces: every rank holds the oracle's counts."""
    script = tmp_path / "worker_py.py"
    script.write_text(PY_WORKER)
    r = torchrun(2, [str(script)])
    assert r.returncode == 0, r.stderr[-3000:]
    import re
    lines = [json.loads(m) for m in re.findall(r"AGOFRT_RESULT (\{[^{}]*\})", r.stdout)]
    assert len(lines) == 2 and all(l["ok"] and l["world"] == 2 for l in lines), (r.stdout[-1500:], r.stderr[-1500:])
