"""CPU: the C-ABI library builds, loads, and exports every symbol include/agofrt.h declares.
No compute calls (there is no GPU here); the no-GPU behaviour must be a loud error, not a fallback."""
import os
import re

import numpy as np
import pytest

from analisi_b200 import cabi
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "agofrt.h")).read()
    return sorted(set(re.findall(r"AGOFRT_API\s+[\w\s\*]+?\b(agofrt_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(cabi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = cabi.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.agofrt_version()


def test_shard_range_partitions():
    for units in (0, 1, 7, 1000, 12864 * 196 * 3, 2 ** 40 + 12345):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                b, e = cabi.shard_range(units, r, world)
                assert b == prev and e >= b
                prev = e
            assert prev == units
            sizes = [cabi.shard_range(units, r, world) for r in range(world)]
            lens = [e - b for b, e in sizes]
            assert max(lens) - min(lens) <= 1
    with pytest.raises(cabi.AgofrtError):
        cabi.shard_range(10, 2, 2)


def test_block_share_deals_every_block_exactly_once():
    """The geometry of agofrt_blocks (no compute): over the ranks, the parts of every block tile [0, world) once and in rank
    order; every rank gets nblocks / world blocks' worth; whole blocks when the blocks divide among the ranks (the
    reference's dealing of blocks to MPI ranks, lib/include/blockaverage.h:146-186), contiguous runs always."""
    for nblocks in (1, 2, 3, 7, 8, 20, 33):
        for world in (1, 2, 3, 4, 8):
            per_rank = [0] * world
            for b in range(nblocks):
                edge = 0
                for r in range(world):
                    a, e = cabi.block_share(nblocks, r, world, b)
                    if e > a:
                        assert a == edge, (nblocks, world, b, r)
                        edge = e
                        per_rank[r] += e - a
                    else:
                        assert (a, e) == (0, 0)
                assert edge == world, (nblocks, world, b)
            assert per_rank == [nblocks] * world      # in units of 1/world block
            for r in range(world):
                mine = [b for b in range(nblocks) if cabi.block_share(nblocks, r, world, b)[1] > 0]
                assert mine == list(range(mine[0], mine[-1] + 1)) if mine else True
                if nblocks % world == 0:
                    assert all(cabi.block_share(nblocks, r, world, b) == (0, world) for b in mine)
                    assert mine == list(range(r * nblocks // world, (r + 1) * nblocks // world))
    assert cabi.block_share(20, 0, 8, 2) == (0, 4) and cabi.block_share(20, 1, 8, 2) == (4, 8)   # 2.5 blocks each
    for bad in ((4, 2, 2, 0), (4, -1, 2, 0), (4, 0, 0, 0), (4, 0, 2, 4)):
        with pytest.raises(cabi.AgofrtError):
            cabi.block_share(*bad)


def test_no_gpu_is_a_loud_error():
    if cabi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(cabi.AgofrtError) as e:
        cabi.Context()
    assert e.value.code == cabi.ERR_CUDA
    assert "no CPU path" in str(e.value)


def test_block_average_entry_points_check_their_arguments():
    """agofrt_blockavg_* / agofrt_plan_last_counts: NULL handles are an argument error (no device needed to say so),
    destroying nothing is fine -- the conventions of the rest of the ABI."""
    L = cabi.lib()
    assert L.agofrt_blockavg_create(None, None) == cabi.ERR_ARG
    assert L.agofrt_blockavg_begin(None, 10) == cabi.ERR_ARG
    assert L.agofrt_blockavg_push(None, None, 1.0) == cabi.ERR_ARG
    assert L.agofrt_blockavg_end(None, 2, None, None) == cabi.ERR_ARG
    assert L.agofrt_plan_last_counts(None, None, 0) == cabi.ERR_ARG
    assert b"NULL" in L.agofrt_last_error()
    assert L.agofrt_blockavg_destroy(None) == cabi.OK


def test_device_slots_helper():
    """every type group padded to 8 slots (kPadGroup): the sizes that pick the small-system kernel"""
    assert cabi.device_slots(np.zeros(56, np.int32)) == 56
    assert cabi.device_slots(np.array([0] * 32 + [1] * 24)) == 56
    assert cabi.device_slots(np.array([0] * 33 + [1] * 23)) == 64
    assert cabi.device_slots(np.arange(75) % 3) == 96
    assert cabi.SMALL_DEFAULT_SLOTS == 256 and cabi.SMALL_MAX_SLOTS == 512
    assert (cabi.OPT_NO_SMALL, cabi.OPT_ON_DEVICE, cabi.OPT_SMALL) == (256, 512, 1024)


def test_option_bits_match_the_header():
    text = open(os.path.join(ROOT, "include", "agofrt.h")).read()
    bits = {k: int(v) for k, v in re.findall(r"AGOFRT_OPT_(\w+)\s*=\s*(\d+)", text)}
    for name in ("EDGES", "FORCE_GENERAL", "NO_AGGREGATE", "AGGREGATE", "NO_SAFE", "DENSE", "SPARSE", "NO_UBOX",
                 "NO_SMALL", "ON_DEVICE", "SMALL"):
        assert getattr(cabi, "OPT_" + name) == bits[name], name


def test_host_side_gofrt_arithmetic_matches_oracle():
    import oracle
    for nts, lmax in ((700, 10), (5, 10), (9, 0), (0, 3)):
        assert cabi.gofrt_leff(nts, lmax) == oracle.leff(nts, lmax)
    for total, n_b, lmax in ((200, 20, 10), (200, 20, 1), (7958, 20, 0), (1000, 1, 201)):
        assert cabi.gofrt_nextra(total, n_b, lmax) == oracle.nextra(total, n_b, lmax)
    assert cabi.gofrt_incr(700, 10) == 1.0 / 70
    assert cabi.gofrt_incr(5, 10) == 1.0
    assert cabi.gofrt_incr(1899, 30) == 1.0 / 63
