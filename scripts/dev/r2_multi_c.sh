#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
export MMDGAN_NVLS_ADAM=1 MMDGAN_NVLS_ONE_GRAPH=1
run 29711 scripts/multi_gpu_check.py 2>&1 | grep -E "MULTI|Error|error" | head -5 | tee gpurun_out/r2_multic_$N.txt
run 29713 bench.py --gpus $N --steps 20 --warmup 5 --no-strong 2> gpurun_out/r2_bench_nvls1_$N.err | tail -1 > gpurun_out/r2_bench_nvls1_$N.json
python -c "
import json
l=json.loads(open('gpurun_out/r2_bench_nvls1_$N.json').read().strip().splitlines()[-1])
print('value', round(l['value']), 'ms', round(l['ms_per_step'],3), 'e2e ms', round(l['e2e']['ms_per_step'],3), 'dp', l.get('dp_equals_single'), l['config'].get('collectives'))
" | tee -a gpurun_out/r2_multic_$N.txt
tail -3 gpurun_out/r2_bench_nvls1_$N.err
