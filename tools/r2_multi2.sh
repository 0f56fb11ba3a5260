#!/bin/bash
# 2-GPU check of everything the 8-GPU session will run
tag=${1:-r2m}; n=${2:-2}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $out/${tag}_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -4 $out/${tag}_pytest_multi.log
for g in 1 $n; do
  timeout 600 python bench.py --workload C1 --gpus $g --steps 3 --warmup 2 > $out/${tag}_C1_n$g.json 2> $out/${tag}_C1_n$g.log
  python -c "
import json;d=json.load(open('$out/${tag}_C1_n$g.json'))
print('C1 gpus $g: device ms %.2f e2e ms %.2f value %.4g n_gpus %d'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['n_gpus']))"
  grep "Time for 20\|batch of" $out/${tag}_C1_n$g.log | tail -2
done
for g in 1 $n; do
  timeout 900 python bench.py --workload C5 --quick --gpus $g --steps 2 --warmup 1 > $out/${tag}_C5q_n$g.json 2> $out/${tag}_C5q_n$g.log
  python -c "
import json;d=json.load(open('$out/${tag}_C5q_n$g.json'))
print('C5 quick gpus $g: device ms %.1f e2e ms %.1f value %.4g e2e %.4g'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['e2e']['value']))"
  grep "Reading time" $out/${tag}_C5q_n$g.log | tail -3
done
