#!/bin/bash
tag=${1:-c48}; out=gpurun_out; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 8 --workload C4 --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_c4_n8.json 2> $out/${tag}_c4_n8.log
echo "rc=$?"; grep '^{' $out/${tag}_c4_n8.json | tail -1 | cut -c1-2000
