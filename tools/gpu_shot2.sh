#!/bin/bash
# Second short session (tight GPU budget): the whole GPU suite, the C1 bench line, then the ncu evidence for the
# small-system kernel and the block-average kernel (full captures of one launch each, launch list of the CLI run).
#   gpurun --timeout 175 -- 'bash tools/gpu_shot2.sh r1n'
tag=${1:-shot2}
out=gpurun_out
mkdir -p $out
C1="bin/analisi -i tests/_refdata/lammps.bin -g 200 -F 0.7 3.5"
echo "== pytest -m gpu"
timeout 100 python -m pytest tests -m gpu -q --durations=3 -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 $out/${tag}_pytest_gpu.log
echo "== bench C1"
timeout 40 python bench.py --workload C1 --steps 3 --warmup 1 > $out/${tag}_bench_c1.json 2> $out/${tag}_bench_c1.log
echo "bench rc=$?"; cut -c1-300 $out/${tag}_bench_c1.json; tail -2 $out/${tag}_bench_c1.log
echo "== ncu full: one pair_small_kernel launch and one blockavg_push_kernel launch of the C1 CLI run"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:'pair_small_kernel|blockavg_push_kernel' -s 4 -c 2 \
    -f -o $out/${tag}_small_c1 $C1 > /dev/null 2> $out/${tag}_ncu_small.log
echo "ncu full rc=$?"; tail -3 $out/${tag}_ncu_small.log
echo "== ncu launch list of the C1 CLI run"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_c1.csv \
    $C1 > /dev/null 2> $out/${tag}_ncu_launch.log
echo "ncu launches rc=$?"; wc -l $out/${tag}_launches_c1.csv
