#!/usr/bin/env python
"""Per-source-line totals of an ncu source page: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv`,
then `python tools/ncu_lines.py X.csv [first_line last_line]` prints the share of warp-stall samples and of executed
instructions per file and for the hottest lines (optionally only lines in the given range of the main file)."""
import csv, sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 1 << 30)
out, hdr, fname = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split('/')[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        try:
            s, ie = int(r[6]), int(r[7])
        except ValueError:
            continue
        stalls = {h: int(v) for h, v in zip(hdr[31:48], r[31:48]) if v.isdigit()}
        out.append((fname, int(r[0]), s, ie, r[1].strip()[:100], stalls))
tot_s, tot_i = sum(o[2] for o in out), sum(o[3] for o in out)
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
bf = defaultdict(lambda: [0, 0])
for o in out:
    bf[o[0]][0] += o[2]
    bf[o[0]][1] += o[3]
for k, v in bf.items():
    print("  %-40s %5.1f%% samples %5.1f%% instructions" % (k, 100 * v[0] / tot_s, 100 * v[1] / tot_i))
main = out[0][0]
sel = [o for o in out if o[0] == main and lo <= o[1] <= hi]
print("lines %d-%d of %s: %.1f%% samples, %.1f%% instructions" % (lo, min(hi, 99999), main, 100 * sum(o[2] for o in sel) / tot_s,
                                                                     100 * sum(o[3] for o in sel) / tot_i))
for o in sorted(sel, key=lambda o: -o[2])[:40]:
    top = sorted(o[5].items(), key=lambda kv: -kv[1])[:2]
    print("%5d %6.2f%% smp %6.2f%% inst  %-22s| %s" % (o[1], 100 * o[2] / tot_s, 100 * o[3] / tot_i,
                                                         " ".join("%s=%d" % (k.replace("stall_", ""), v) for k, v in top if v), o[4]))
