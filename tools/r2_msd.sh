#!/bin/bash
tag=${1:-r2m}; out=gpurun_out; mkdir -p $out
timeout 200 python -m pytest tests -m gpu -x -q -k "msd or MSD" 2>&1 | tail -3
echo "== ring"; timeout 200 python tools/msd_rate.py > $out/${tag}_msd_rate_ring.txt 2>/dev/null; cat $out/${tag}_msd_rate_ring.txt
echo "== no ring"; AGOFRT_MSD_RING=0 timeout 200 python tools/msd_rate.py > $out/${tag}_msd_rate_noring.txt 2>/dev/null; cat $out/${tag}_msd_rate_noring.txt
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --print-units base --csv -k regex:msd_ring --log-file $out/${tag}_msd_c4full_ring.csv python tools/msd_rate.py C4full > /dev/null 2>&1
tail -3 $out/${tag}_msd_c4full_ring.csv
