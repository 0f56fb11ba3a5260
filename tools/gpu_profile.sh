#!/bin/bash
# Evidence session: full bench line, reference arm, ncu launch list, full captures, DRAM traffic of a full-size launch.
#   gpurun --timeout 1800 -- 'bash tools/gpu_profile.sh <tag>'
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
echo "== bench C2 (default)"
timeout 900 python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.log
echo "bench rc=$?"; cat $out/${tag}_bench_c2.json
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.log
cat $out/${tag}_bench_ref.json
echo "== bench C3 / C4 quick"
for wl in C3 C4; do
  timeout 600 python bench.py --workload $wl --quick --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_${wl}q.json 2> $out/${tag}_bench_${wl}q.log
  cat $out/${tag}_bench_${wl}q.json
done
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
echo "== DRAM traffic of one FULL-SIZE C2 launch"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:pair_kernel -c 1 --csv --log-file $out/${tag}_traffic_c2full.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $out/${tag}_ncu_traffic.log 2>&1
echo "ncu traffic rc=$?"; tail -3 $out/${tag}_traffic_c2full.csv
ls -la $out | tail -20
