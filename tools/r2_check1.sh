#!/bin/bash
# 1-GPU check: default bench (e2e from page-locked and from pageable arrays), full ncu capture of the small-system kernel on C1
tag=${1:-r2v}; out=gpurun_out; mkdir -p $out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-traffic > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.log
echo "bench rc=$?"; python -c "
import json;d=json.loads([l for l in open('$out/${tag}_bench_default.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g ms %.1f e2e_ms %.1f pageable_ms %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['from_pageable_arrays']['ms_per_step']))
print(d['e2e']['last_step_breakdown_ms']); print(d['e2e']['from_pageable_arrays']['breakdown_ms'])"
tail -3 $out/${tag}_bench_default.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_small -s 3 -c 1 \
    -f -o $out/${tag}_small_c1 python bench.py --workload C1 --steps 1 --warmup 1 --no-cpu-baseline --no-traffic --no-e2e > $out/${tag}_ncu_c1.log 2>&1
echo "ncu c1 rc=$?"; tail -2 $out/${tag}_ncu_c1.log
ls -la $out/${tag}_*
