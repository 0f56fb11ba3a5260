#!/bin/bash
# One GPU-box session: parity tests, bench lines, the ncu launch list and one full capture of the
# pair kernel.  Run with:  gpurun --timeout 1500 -- 'bash tools/gpu_session.sh [tag]'
# Everything lands in gpurun_out/<tag>_*; summaries are copied to profiles/ by hand afterwards.
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1

echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log

echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log

echo "== bench C2 (default)"
timeout 900 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log
echo "bench rc=$?"; cat $out/${tag}_bench_c2.json

echo "== bench C4 quick / C3 quick (triclinic kernels)"
timeout 600 python bench.py --workload C4 --quick --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c4q.json 2> $out/${tag}_bench_c4q.log
cat $out/${tag}_bench_c4q.json
timeout 600 python bench.py --workload C3 --quick --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c3q.json 2> $out/${tag}_bench_c3q.log
cat $out/${tag}_bench_c3q.json

echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log
cat $out/${tag}_bench_ref.json

echo "== ncu launch list (quick C2 block)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_c2.csv python bench.py --quick --steps 2 --warmup 1 --no-cpu-baseline \
    > $out/${tag}_ncu_launch.log 2>&1
echo "ncu launches rc=$?"

echo "== ncu full capture of the pair kernel (quick C2, quick C4)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 \
    -f -o $out/${tag}_pair_c2 python bench.py --quick --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 \
    -f -o $out/${tag}_pair_c4 python bench.py --workload C4 --quick --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
ls -la $out
