#!/bin/bash
out=gpurun_out; tag=${1:-r2l}; mkdir -p $out
for env in "ANALISI_DEVICE_PARSE=1 ANALISI_BLOCK_BATCH=1" "ANALISI_DEVICE_PARSE=0 ANALISI_BLOCK_BATCH=1"; do
  name=$(echo $env | tr ' =' '__')
  env $env AGOFRT_DEBUG=1 timeout 600 python bench.py --workload C1 --steps 3 --warmup 3 > $out/${tag}_C1_$name.json 2> $out/${tag}_C1_$name.log
  python -c "
import json;d=json.load(open('$out/${tag}_C1_$name.json'))
print('$env', 'device ms %.2f'%d['ms_per_step'], 'e2e ms %.2f'%d['e2e']['ms_per_step'], 'frac %.3f'%d['roofline']['frac'])"
  grep "agofrt\]\|Reading time\|Time for 20\|\[blocks\]" $out/${tag}_C1_$name.log | tail -9
done
python tools/msd_rate.py > $out/${tag}_msd_rate.txt 2> $out/${tag}_msd_rate.jsonl; cat $out/${tag}_msd_rate.txt
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --print-units base --csv -k regex:msd_partial --log-file $out/${tag}_msd_c4full_dram.csv python tools/msd_rate.py C4full > /dev/null 2>&1
tail -4 $out/${tag}_msd_c4full_dram.csv
python tools/neighbour_rate.py > $out/${tag}_neighbour_rate.txt 2>&1; cat $out/${tag}_neighbour_rate.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "msd or MSD or neighbour or Neighbour" 2>&1 | tail -3
