#!/bin/bash
# gpurun [--gpus N] -- 'bash tools/gpu_c5b.sh <tag> <N>': host-layer tests + C5 full on N GPUs of one process
tag=${1:-c5b}; n=${2:-1}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_multi.py -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 $out/${tag}_pytest.log
timeout 900 python bench.py --workload C5 --gpus $n --steps 2 --warmup 1 > $out/${tag}_c5_n$n.json 2> $out/${tag}_c5_n$n.log
echo "c5 rc=$?"; cat $out/${tag}_c5_n$n.json | cut -c1-200; grep "pass\|wrote" $out/${tag}_c5_n$n.log; grep "Reading\|Time for block" $out/${tag}_c5_n$n.log | tail -8
