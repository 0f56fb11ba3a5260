#!/usr/bin/env python
"""SASS and ptxas evidence of the shipped library -> profiles/r2_sass_counts.txt

    python tools/sass_evidence.py > profiles/r2_sass_counts.txt

Per kernel of interest: registers / spills (ptxas -v, analisi_b200/csrc/ptxas.log), and the number of SASS instructions
of the kinds the design argues with (cuobjdump -sass of analisi_b200/libagofrt.so): UBLKCP (cp.async.bulk, the TMA
engine), SYNCS (mbarrier), FP64 pipe (DADD / DMUL / DFMA / DSETP), ATOMS.POPC.INC (hardware-merged shared increment),
MUFU / F2F (XU pipe), LOP3.  For the two hot variants also the FP64 instructions per pair evaluation of the inner loop
(the longest straight-line run of FP64 instructions / 16 pairs): the figure bench.py prints as
fp64_instr_issued_per_pair_eval."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "analisi_b200", "libagofrt.so")
LOG = os.path.join(ROOT, "analisi_b200", "csrc", "ptxas.log")

KERNELS = [
    ("pair_kernel<ortho, single-pass, safe-zone dense, ubox>  (C2)", "_ZN6agofrt11pair_kernelILb0ELb1ELi4ELb1EEEvNS_10PairParamsE"),
    ("pair_kernel<triclinic, single-pass, safe-zone, ubox>  (C3, C4: the default bench step)", "_ZN6agofrt11pair_kernelILb1ELb1ELi3ELb1EEEvNS_10PairParamsE"),
    ("pair_kernel<ortho, single-pass, two-floor, ubox>  (AGOFRT_OPT_SAFE2)", "_ZN6agofrt11pair_kernelILb0ELb1ELi5ELb1EEEvNS_10PairParamsE"),
    ("pair_kernel<triclinic, general minimum image, thresholds>", "_ZN6agofrt11pair_kernelILb1ELb0ELi0ELb0EEEvNS_10PairParamsE"),
    ("pair_small_kernel<ortho, single-pass, safe-zone dense, ubox>  (C1)", "_ZN6agofrt17pair_small_kernelILb0ELb1ELi4ELb1EEEvNS_10PairParamsE"),
    ("neighbour_kernel<triclinic, single-pass>", "_ZN6agofrt16neighbour_kernelILb1ELb1EEEvNS_15NeighbourParamsE"),
    ("msd_partial_kernel", "_ZN6agofrt18msd_partial_kernelENS_9MsdParamsE"),
    ("parse_records_kernel", "_ZN6agofrt20parse_records_kernelEPKdiiPKiiS3_PdPj"),
    ("neigh_list_kernel", "_ZN6agofrt17neigh_list_kernelENS_15NeighListParamsE"),
    ("sh_density_kernel", "_ZN6agofrt17sh_density_kernelENS_15ShDensityParamsE"),
]
KINDS = ["UBLKCP", "SYNCS", "DADD", "DMUL", "DFMA", "DSETP", "ATOMS.POPC.INC", "ATOMS", "MUFU", "F2F", "LOP3", "FFMA", "BAR", "LDS", "LDG", "STG"]


def ptxas_info():
    out = {}
    if not os.path.exists(LOG):
        return out
    cur = None
    for line in open(LOG):
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            cur = m.group(1)
            out[cur] = {}
        elif cur and "bytes stack frame" in line and "spill" not in out[cur]:
            out[cur]["spill"] = line.strip()   # (the first one: the entry function itself; those of its noinline callees follow)
        elif cur and "Used" in line and "registers" in line:
            out[cur]["regs"] = line.strip().replace("ptxas info    : ", "")
    return out


def sass(fun):
    r = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    ops = []
    for l in r.stdout.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            t = m.group(1).split()
            ops.append(t[1] if t[0].startswith("@") else t[0])
    return ops


def main():
    info = ptxas_info()
    print("# SASS / ptxas evidence of analisi_b200/libagofrt.so (tools/sass_evidence.py; nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)")
    for title, fun in KERNELS:
        ops = sass(fun)
        if not ops:
            print("\n%s\n  (not in the library)" % title)
            continue
        print("\n%s\n  %s" % (title, fun))
        pi = info.get(fun, {})
        if pi:
            print("  ptxas: %s; %s" % (pi.get("regs", "?"), pi.get("spill", "?")))
        counts = {}
        for k in KINDS:
            counts[k] = sum(1 for o in ops if o == k or o.startswith(k + "."))
        counts["ATOMS"] -= 0
        print("  instructions: %d   " % len(ops) + "  ".join("%s %d" % (k, counts[k]) for k in KINDS if counts[k]))
        # the longest run of instructions between two branches: the unrolled group of 16 pairs
        fp64 = ("DADD", "DMUL", "DFMA", "DSETP")
        best, run = [], []
        for o in ops:
            if o.split(".")[0] in ("BRA", "EXIT", "CALL", "RET", "BSYNC", "BSSY"):
                if sum(1 for x in run if x.split(".")[0] in fp64) > sum(1 for x in best if x.split(".")[0] in fp64):
                    best = run
                run = []
            else:
                run.append(o)
        nf = sum(1 for x in best if x.split(".")[0] in fp64)
        if "pair_" in fun and nf:
            print("  largest basic block: %d instructions, %d of them FP64 = %.2f FP64 instructions and %.2f others per pair evaluation (16 pairs per group)"
                  % (len(best), nf, nf / 16.0, (len(best) - nf) / 16.0))


if __name__ == "__main__":
    main()
