#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python tools/ncu_summary.py full   gpurun_out/x.ncu-rep   > profiles/x_full.txt
    python tools/ncu_summary.py launch gpurun_out/launches.csv > profiles/x_launches.txt

`full` reads the raw page of one `ncu --set full` capture and prints the metrics the design
argues with (FP64 pipe, issue slots, shared atomics, DRAM traffic).  `launch` reads a
`--metrics gpu__time_duration.sum` launch list and prints per-kernel totals and shares.
"""
import csv
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max",
    "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.avg.per_cycle_active",
    "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s (ncu --set full --clock-control none; one launch, ~40 replays: cold-cache, not a bench number)" % path)
    for vals in rows[2:]:
        d = OrderedDict((h, (u, v)) for h, u, v in zip(hdr, units, vals))
        print("kernel: %s   grid %s block %s" % (d["Kernel Name"][1], d.get("Grid Size", ("", ""))[1], d.get("Block Size", ("", ""))[1]))
        for k in KEYS:
            if k in d:
                print("  %-82s %s %s" % (k, d[k][1], d[k][0]))
        try:
            fp64 = float(d["sm__inst_executed_pipe_fp64.sum"][1])
            tot = float(d["smsp__inst_executed.sum"][1])
            print("  derived: FP64-pipe share of warp instructions = %.3f" % (fp64 / tot))
            rd = d["dram__bytes_read.sum"]
            wr = d["dram__bytes_write.sum"]
            print("  derived: DRAM traffic per launch = %s %s read + %s %s written" % (rd[1], rd[0], wr[1], wr[0]))
        except (KeyError, ValueError, ZeroDivisionError):
            pass


def launch(path):
    tot = OrderedDict()
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        ns = float(r["Metric Value"])
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ns
    total = sum(v[1] for v in tot.values())
    step = sum(v[1] for k, v in tot.items() if "dfma_peak" not in k and "validate_safe" not in k)
    print("# %s (ncu --metrics gpu__time_duration.sum --clock-control none; serialised, cold-cache: compare SHARES)" % path)
    print("%-60s %8s %14s %8s %10s" % ("kernel", "launches", "total ms", "share", "share*"))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        star = "" if ("dfma_peak" in k or "validate_safe" in k) else "%9.2f%%" % (100 * v[1] / step)
        print("%-60s %8d %14.3f %7.2f%% %10s" % (k[:60], v[0], v[1] / 1e6, 100 * v[1] / total, star))
    print("share* = share of the bench steps proper (without the FP64-peak microbenchmark and the plan validation probe)")


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2])
