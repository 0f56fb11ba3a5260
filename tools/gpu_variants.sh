#!/bin/bash
# gpurun --timeout 1200 -- 'bash tools/gpu_variants.sh <tag> lib1.so lib2.so ...': parity subset + quick benches per tuning build
tag=$1; shift
out=gpurun_out; mkdir -p $out
for lib in "$@"; do
  export AGOFRT_LIB=$PWD/analisi_b200/$lib
  echo "== $lib"
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "live_reference or multi_tile or guarded" 2>&1 | tail -1
  for wl in C2 C4; do
    timeout 300 python bench.py --workload $wl --quick --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('   $wl quick: %.4g pairs/s  frac %.4f' % (d['value'], r['frac']))"
  done
done
