#!/bin/bash
# Round-2 session 1: GPU suite, the new default bench (C4 subset) + reference arm, C2 for comparison,
# launch lists and full captures of the pair kernel on the quick C4 / C2 blocks.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
echo "== bench default (C4 subset)"
timeout 1200 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.log
echo "bench rc=$?"; cat $out/${tag}_bench_default.json; tail -5 $out/${tag}_bench_default.log
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log
echo "ref rc=$?"; cat $out/${tag}_bench_ref.json
echo "== bench C2"
timeout 900 python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log
cat $out/${tag}_bench_c2.json
echo "== ncu launch list (default step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-traffic \
    > $out/${tag}_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
echo "== ncu full (quick C4, quick C2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 \
    -f -o $out/${tag}_pair_c4 python bench.py --workload C4 --quick --steps 1 --warmup 1 --no-cpu-baseline --no-traffic > $out/${tag}_ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 \
    -f -o $out/${tag}_pair_c2 python bench.py --workload C2 --quick --steps 1 --warmup 1 --no-cpu-baseline --no-traffic > $out/${tag}_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
ls -la $out | tail -20
