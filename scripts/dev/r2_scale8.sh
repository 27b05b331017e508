#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r2_bench_nccl_$N.err | tail -1 > gpurun_out/r2_bench_nccl_$N.json
python -c "
import json
l=json.loads(open('gpurun_out/r2_bench_nccl_$N.json').read().strip().splitlines()[-1])
print('N=$N value', round(l['value']), 'ms', round(l['ms_per_step'],3), 'e2e ms', round(l['e2e']['ms_per_step'],3), 'dp', l.get('dp_equals_single'), 'strong', (l['config'].get('strong') or {}))
"
tail -3 gpurun_out/r2_bench_nccl_$N.err
