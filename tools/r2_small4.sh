#!/bin/bash
# short timeouts: a kernel that hangs must not eat the session
tag=${1:-r2x}; out=gpurun_out; mkdir -p $out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/${tag}_pytest_parity.log 2>&1
rc=$?; echo "parity rc=$rc"; tail -3 $out/${tag}_pytest_parity.log
[ $rc -ne 0 ] && exit 1
timeout 120 python bench.py --workload C1 --steps 5 --warmup 3 > $out/${tag}_C1.json 2> $out/${tag}_C1.log
echo "C1 rc=$?"; python -c "
import json;d=json.load(open('$out/${tag}_C1.json'))
print('C1: device ms %.2f e2e ms %.2f frac %.3f value %.4g'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['value']))"
timeout 180 ncu --set full --clock-control none --import-source on -k regex:pair_small -s 3 -c 1 \
    -f -o $out/${tag}_small_c1 python bench.py --workload C1 --steps 1 --warmup 1 --no-cpu-baseline --no-traffic --no-e2e > $out/${tag}_ncu_c1.log 2>&1
echo "ncu c1 rc=$?"
