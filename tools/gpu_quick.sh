#!/bin/bash
# Short GPU-box session: parity tests (all layers), smoke, one bench line.
#   gpurun --timeout 1200 -- 'bash tools/gpu_quick.sh [tag] [bench args...]'
tag=${1:-q}
shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -25 $out/${tag}_pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
echo "== bench"
timeout 900 python bench.py "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.log
echo "bench rc=$?"; cat $out/${tag}_bench.json; tail -5 $out/${tag}_bench.log
