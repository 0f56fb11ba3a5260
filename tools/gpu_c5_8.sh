#!/bin/bash
tag=${1:-c58}; n=${2:-8}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $out/${tag}_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -4 $out/${tag}_pytest_multi.log
timeout 900 python bench.py --workload C5 --gpus $n --steps 2 --warmup 1 > $out/${tag}_c5_n$n.json 2> $out/${tag}_c5_n$n.log
echo "c5 rc=$?"; cat $out/${tag}_c5_n$n.json | cut -c1-300; grep "pass\|wrote" $out/${tag}_c5_n$n.log
