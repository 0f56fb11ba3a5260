#!/bin/bash
# tuning builds (AGOFRT_IPT x AGOFRT_JU): C2 with the two-floor and the clamped dense kernel, C4 subset
tag=$1; shift
out=gpurun_out; mkdir -p $out
for lib in "$@"; do
  export AGOFRT_LIB=$PWD/analisi_b200/$lib
  for spec in "C2 0" "C2 2048" "C4 0" "C3 0"; do
    set -- $spec
    timeout 600 python bench.py --workload $1 --options $2 --steps 2 --warmup 1 --no-cpu-baseline --no-traffic --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['roofline']
print('$lib $1 opt $2: %.4g pairs/s  frac %.4f  %.1f ms  %s' % (d['value'], r['frac'], d['ms_per_step'], d['counts_sha256'][:12]))" | tee -a $out/${tag}_variants.txt
  done
done
