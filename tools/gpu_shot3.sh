#!/bin/bash
# Last short session of the round (about a minute of GPU budget left): the GPU suite without its two slowest
# tests, the small-system rates, the C1 bench line, smoke.
#   gpurun --timeout 62 -- 'bash tools/gpu_shot3.sh r1p'
tag=${1:-shot3}
out=gpurun_out
mkdir -p $out
timeout 45 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not c4_size and not 100k_atoms" > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 $out/${tag}_pytest_gpu.log
timeout 12 python tools/small_rate.py > $out/${tag}_small_rate.jsonl 2> $out/${tag}_small_rate.log
echo "rate rc=$?"; cut -c1-260 $out/${tag}_small_rate.jsonl
timeout 15 python bench.py --workload C1 --steps 3 --warmup 1 > $out/${tag}_bench_c1.json 2> $out/${tag}_bench_c1.log
echo "bench rc=$?"; cut -c1-200 $out/${tag}_bench_c1.json
timeout 12 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
