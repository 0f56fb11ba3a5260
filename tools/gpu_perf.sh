#!/bin/bash
# Kernel iteration session: parity tests of the C-ABI layer, then quick + full bench lines.
#   gpurun --timeout 1200 -- 'bash tools/gpu_perf.sh <tag> [full]'
tag=${1:-p}; full=${2:-}
out=gpurun_out
mkdir -p $out
echo "== pytest (C ABI parity)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 $out/${tag}_pytest.log
for wl in C2 C3 C4; do
  timeout 600 python bench.py --workload $wl --quick --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_${wl}q.json 2> $out/${tag}_${wl}q.log
  python - <<PY
import json
try:
    d=json.loads(open("$out/${tag}_${wl}q.json").read().strip().splitlines()[-1])
    r=d["roofline"]; print("$wl quick: %.4g pairs/s  frac %.4f  kernel_ms %.3f in_range %.3f" % (d["value"], r["frac"], r["kernel_ms_per_step"], r["in_range_fraction"]))
except Exception as e: print("$wl failed", e)
PY
done
if [ -n "$full" ]; then
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_C2full.json 2> $out/${tag}_C2full.log
  cat $out/${tag}_C2full.json
fi
