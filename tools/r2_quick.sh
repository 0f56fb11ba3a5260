#!/bin/bash
# short session: GPU parity suite + default bench (+ optional extra bench workloads given as arguments)
tag=${1:-r2q}; shift
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-traffic > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.log
echo "bench rc=$?"; cat $out/${tag}_bench_default.json
for wl in "$@"; do
  timeout 900 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-traffic > $out/${tag}_bench_${wl}.json 2> $out/${tag}_bench_${wl}.log
  echo "bench $wl rc=$?"; cat $out/${tag}_bench_${wl}.json
done
