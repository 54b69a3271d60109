#!/bin/bash
# N-GPU checks: parity of the sharded paths, then bench at N.  usage: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/mgpu_check.py 2>&1 | grep -v "^W\|warn" | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "bench exit $?"
grep '^{"metric"' gpurun_out/bench_n$N.log | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln); print('N', d['n_gpus'], 'cells/s', round(d['value']), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['stages_ms'])"
tail -3 gpurun_out/bench_n$N.log | cut -c1-300
