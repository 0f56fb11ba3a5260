#!/bin/bash
# 8-GPU session:  gpurun --gpus 8 --timeout 1500 -- 'bash tools/gpu_scale8.sh <tag>'
tag=${1:-s8}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
echo "== pytest multi-GPU"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $out/${tag}_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -4 $out/${tag}_pytest_multi.log
run() {  # name nproc args...
  name=$1; n=$2; shift; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $n "$@" > $out/${tag}_${name}_n$n.json 2> $out/${tag}_${name}_n$n.log
  echo "$name n=$n rc=$?"; grep '^{' $out/${tag}_${name}_n$n.json | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('  value %.4g  e2e %.4g  ms/step %.1f  frac %.4f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'] or 0))"
}
run c2 8 --steps 3 --warmup 3
run c2 4 --steps 3 --warmup 3
run c4 8 --workload C4 --steps 1 --warmup 1 --no-cpu-baseline
echo "== C5, one process driving 8 GPUs"
timeout 900 python bench.py --workload C5 --gpus 8 --steps 2 --warmup 1 > $out/${tag}_c5_n8.json 2> $out/${tag}_c5_n8.log
echo "c5 rc=$?"; cat $out/${tag}_c5_n8.json | cut -c1-400; grep "pass\|wrote" $out/${tag}_c5_n8.log
