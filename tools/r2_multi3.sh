#!/bin/bash
tag=${1:-r2t}; n=${2:-2}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > $out/${tag}_pytest_multi.log 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_pytest_multi.log
AGOFRT_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $n --steps 4 --warmup 2 --no-traffic --no-cpu-baseline > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.log
python -c "
import json;d=json.loads([l for l in open('$out/${tag}_bench_n$n.json') if l.startswith('{')][-1])
print('N=$n value %.4g e2e %.4g ms %.1f e2e_ms %.1f sha %s kernel_ms %.1f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['counts_sha256'][:12], d['roofline']['kernel_ms_per_step']), d['e2e']['last_step_breakdown_ms'])"
grep "agofrt\] upload" $out/${tag}_bench_n$n.log | tail -3
for g in 1 $n; do
  AGOFRT_DEBUG=1 timeout 600 python bench.py --workload C1 --gpus $g --steps 3 --warmup 2 > $out/${tag}_C1_n$g.json 2> $out/${tag}_C1_n$g.log
  python -c "
import json;d=json.load(open('$out/${tag}_C1_n$g.json'))
print('C1 gpus $g: device ms %.2f e2e ms %.2f value %.4g n_gpus %d'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['n_gpus']))"
  grep "batch of\|\[blocks\]" $out/${tag}_C1_n$g.log | tail -5
done
