#!/bin/bash
# One short GPU-box session for the small-system kernel and the device block averages (tight GPU budget):
# smoke first (a hang or a wrong count stops the session), then the whole GPU suite, then the C1 bench line with
# and without the two new paths.
#   gpurun --timeout 240 -- 'bash tools/gpu_shot.sh r1m'
tag=${1:-shot}
out=gpurun_out
mkdir -p $out
echo "== smoke"
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
rc=$?
echo "smoke rc=$rc"; tail -4 $out/${tag}_smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
echo "== pytest -m gpu"
timeout 170 python -m pytest tests -m gpu -q --durations=5 -p no:cacheprovider > $out/${tag}_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 $out/${tag}_pytest_gpu.log
echo "== bench C1 (small-system kernel, MediaVarDevice)"
timeout 50 python bench.py --workload C1 --steps 3 --warmup 1 > $out/${tag}_bench_c1.json 2> $out/${tag}_bench_c1.log
echo "bench rc=$?"; cat $out/${tag}_bench_c1.json | cut -c1-600; tail -4 $out/${tag}_bench_c1.log
echo "== bench C1 (tile kernel, host MediaVar)"
ANALISI_KERNEL_OPTIONS=256 ANALISI_DEVICE_BLOCKS=0 timeout 50 python bench.py --workload C1 --steps 3 --warmup 1 > $out/${tag}_bench_c1_old.json 2> $out/${tag}_bench_c1_old.log
echo "bench rc=$?"; cat $out/${tag}_bench_c1_old.json | cut -c1-600; tail -4 $out/${tag}_bench_c1_old.log
