#!/bin/bash
# gpurun --timeout 1500 -- 'bash tools/gpu_all.sh <tag>': every -m gpu test + smoke + C5 quick
tag=${1:-all}; out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -12 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
timeout 900 python bench.py --workload C5 --quick --steps 2 --warmup 1 > $out/${tag}_c5q.json 2> $out/${tag}_c5q.log
echo "c5q rc=$?"; grep "pass\|ahead" $out/${tag}_c5q.log | tail -12
