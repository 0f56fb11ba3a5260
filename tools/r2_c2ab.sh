#!/bin/bash
# C2 A/B: two-floor kernel with / without the phase skew, and the r1 clamped kernel; plan info of the rmin > 0 case
out=gpurun_out; tag=${1:-r2e}; mkdir -p $out
AGOFRT_DEBUG=1 python - > $out/${tag}_planinfo.txt 2>&1 <<'PY'
import numpy as np
from analisi_b200 import cabi, synth
ctx = cabi.Context([0])
pos, box, types = synth.small_case(43, (12, 12, 12), 1.1, 1, False, 5, "parity")
bi = synth.lammps_rows_to_internal(box)
tr = cabi.DeviceTrajectory(ctx, pos.shape[1], 6, types, 1, 5)
for (rmin, rmax, nbin) in ((0.6, 6.6, 100), (0.0, 6.0, 96), (0.7, 3.5, 200), (0.5, 3.8, 100)):
    pl = cabi.Plan(tr, rmin, rmax, nbin)
    print(rmin, rmax, nbin, pl.info())
    pl.close()
PY
cat $out/${tag}_planinfo.txt
for opt in 0 4096 2048; do
  timeout 900 python bench.py --workload C2 --steps 2 --warmup 2 --no-cpu-baseline --no-traffic --options $opt > $out/${tag}_bench_C2_opt$opt.json 2> $out/${tag}_bench_C2_opt$opt.log
  python -c "
import json;d=json.load(open('$out/${tag}_bench_C2_opt$opt.json'))
print('opt $opt', '%.4g'%d['value'], 'frac %.4f'%d['roofline']['frac'], 'ms', '%.1f'%d['ms_per_step'], d['counts_sha256'][:12])"
done
