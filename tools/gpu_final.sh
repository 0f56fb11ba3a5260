#!/bin/bash
tag=${1:-fin}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -4 $out/${tag}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --workload C3 --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c3_full.json 2> $out/${tag}_bench_c3_full.log
echo "c3 rc=$?"; cat $out/${tag}_bench_c3_full.json | cut -c1-1600
