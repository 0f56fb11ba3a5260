#!/bin/bash
tag=${1:-r2f}; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-traffic > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.log
echo "bench rc=$?"; cat $out/${tag}_bench_default.json | cut -c1-1500; tail -3 $out/${tag}_bench_default.log
bash tools/r2_variants.sh $tag libagofrt.so libagofrt_i4j4.so libagofrt_i4j2.so libagofrt_i3j4.so
