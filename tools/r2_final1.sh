#!/bin/bash
# final 1-GPU evidence session of round 2 (every step under its own timeout)
tag=${1:-r2f1}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_gpu.log
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $out/${tag}_smoke.log
echo "== bench default"
timeout 400 python bench.py --steps 5 --warmup 3 > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.log
echo "bench rc=$?"; cut -c1-2500 $out/${tag}_bench_default.json
echo "== reference arm"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log
echo "ref rc=$?"; cut -c1-600 $out/${tag}_bench_ref.json
echo "== C2, C3, C1"
timeout 300 python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log
timeout 300 python bench.py --workload C3 --steps 2 --warmup 3 --no-cpu-baseline --no-traffic > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.log
timeout 200 python bench.py --workload C1 --steps 9 --warmup 3 > $out/${tag}_bench_c1.json 2> $out/${tag}_bench_c1.log
for w in c2 c3 c1; do python -c "
import json;d=json.loads([l for l in open('$out/${tag}_bench_$w.json') if l.startswith('{')][-1])
print('$w value %.4g e2e %.4g ms %.2f e2e_ms %.2f frac %.3f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac']))"; done
echo "== ncu launch list (default step)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-traffic \
    > $out/${tag}_ncu_launch.log 2>&1
echo "ncu launches rc=$?"
echo "== ncu full (quick C4, quick C2)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 \
    -f -o $out/${tag}_pair_c4 python bench.py --workload C4 --quick --steps 1 --warmup 1 --no-cpu-baseline --no-traffic --no-e2e > $out/${tag}_ncu_c4.log 2>&1
echo "ncu c4 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 \
    -f -o $out/${tag}_pair_c2 python bench.py --workload C2 --quick --steps 1 --warmup 1 --no-cpu-baseline --no-traffic --no-e2e > $out/${tag}_ncu_c2.log 2>&1
echo "ncu c2 rc=$?"
ls -la $out | grep ${tag} | tail -30
