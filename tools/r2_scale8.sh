#!/bin/bash
# N-GPU session: multi-GPU tests, the default bench at 1/2/4/N GPUs (torchrun), C1 on 1 and N GPUs
tag=${1:-r2s}; n=${2:-8}; list=${3:-"1 2 4 $n"}; out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $out/${tag}_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_multi.log
for g in $list; do
  [ $g -gt $n ] && continue
  if [ $g -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-traffic --no-cpu-baseline > $out/${tag}_bench_n$g.json 2> $out/${tag}_bench_n$g.log
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g \
      bench.py --gpus $g --steps 5 --warmup 3 --no-traffic --no-cpu-baseline > $out/${tag}_bench_n$g.json 2> $out/${tag}_bench_n$g.log
  fi
  python -c "
import json;d=json.loads([l for l in open('$out/${tag}_bench_n$g.json') if l.startswith('{')][-1])
print('N=$g value %.4g e2e %.4g ms %.1f e2e_ms %.1f pageable_ms %.1f sha %s eq_ref %s frac %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['from_pageable_arrays']['ms_per_step'], d['counts_sha256'][:12], d['counts_check']['equal_to_reference'], d['roofline']['frac']))"
  grep "e2e steps" $out/${tag}_bench_n$g.log | head -2
done
for g in $n; do
  timeout 200 python bench.py --workload C1 --gpus $g --steps 3 --warmup 2 > $out/${tag}_C1_n$g.json 2> $out/${tag}_C1_n$g.log
  python -c "
import json;d=json.load(open('$out/${tag}_C1_n$g.json'))
print('C1 gpus $g: device ms %.2f e2e ms %.2f value %.4g n_gpus %d'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['n_gpus']))"
done
