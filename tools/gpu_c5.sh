#!/bin/bash
# gpurun [--gpus N] --timeout 1800 -- 'bash tools/gpu_c5.sh <tag> <N> [quick|full]'
tag=${1:-c5}; n=${2:-1}; mode=${3:-quick}
out=gpurun_out; mkdir -p $out
echo "== pytest host layers"
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q > $out/${tag}_pytest_host.log 2>&1
echo "pytest rc=$?"; tail -8 $out/${tag}_pytest_host.log
df -h /tmp | tail -1; free -g | head -2
echo "== bench C5 quick ($n GPU)"
timeout 900 python bench.py --workload C5 --quick --gpus $n --steps 2 --warmup 1 > $out/${tag}_c5q_n$n.json 2> $out/${tag}_c5q_n$n.log
echo "rc=$?"; cat $out/${tag}_c5q_n$n.json; grep "pass\|wrote" $out/${tag}_c5q_n$n.log
if [ "$mode" = "full" ]; then
  echo "== bench C5 full ($n GPU)"
  timeout 1500 python bench.py --workload C5 --gpus $n --steps 2 --warmup 1 > $out/${tag}_c5_n$n.json 2> $out/${tag}_c5_n$n.log
  echo "rc=$?"; cat $out/${tag}_c5_n$n.json; grep "pass\|wrote\|Reading" $out/${tag}_c5_n$n.log | tail -30
fi
