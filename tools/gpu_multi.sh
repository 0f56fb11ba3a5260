#!/bin/bash
# Multi-GPU session:  gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_multi.sh <tag> <N> [bench args]'
tag=${1:-m}; n=${2:-2}; shift; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
nvidia-smi topo -m >> $out/${tag}_gpu.txt 2>&1
echo "== pytest multi-GPU"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $out/${tag}_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -15 $out/${tag}_pytest_multi.log
for g in 1 $n; do
  echo "== bench --gpus $g"
  if [ $g -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu-baseline "$@" > $out/${tag}_bench_n$g.json 2> $out/${tag}_bench_n$g.log
  else
    NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $g "$@" > $out/${tag}_bench_n$g.json 2> $out/${tag}_bench_n$g.log
  fi
  echo "rc=$?"; grep '^{' $out/${tag}_bench_n$g.json | tail -1; grep -i "NVLS\|error" $out/${tag}_bench_n$g.log | head -5
done
