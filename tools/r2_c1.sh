#!/bin/bash
# C1 (bundled 56-atom trajectory, 20 blocks) with the ingest / batch switches, timings on stderr
out=gpurun_out; tag=${1:-r2k}; mkdir -p $out
for env in "ANALISI_DEVICE_PARSE=1 ANALISI_BLOCK_BATCH=1" "ANALISI_DEVICE_PARSE=0 ANALISI_BLOCK_BATCH=1" "ANALISI_DEVICE_PARSE=1 ANALISI_BLOCK_BATCH=0" "ANALISI_DEVICE_PARSE=0 ANALISI_BLOCK_BATCH=0"; do
  name=$(echo $env | tr ' =' '__')
  env $env AGOFRT_DEBUG=1 timeout 600 python bench.py --workload C1 --steps 3 --warmup 3 > $out/${tag}_C1_$name.json 2> $out/${tag}_C1_$name.log
  python -c "
import json;d=json.load(open('$out/${tag}_C1_$name.json'))
print('$env', 'device ms %.2f'%d['ms_per_step'], 'e2e ms %.2f'%d['e2e']['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
  grep "agofrt\] upload\|Reading time\|Time for 20" $out/${tag}_C1_$name.log | tail -4
done
