#!/bin/bash
# The full N-GPU bench line with the overlapped all-reduce on its own capped communicator (MMDGAN_AR_CTAS), twice, interruptible.
N=${1:-2}; C=${2:-4}
mkdir -p gpurun_out
for i in 1 2; do
MMDGAN_AR_CTAS=$C timeout -s INT -k 15 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2973$i bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/r2_bench_arctas_$N.err | tail -1 > gpurun_out/r2_bench_arctas_$N.json
echo "run $i rc=$?"
python -c "
import json
l=json.loads(open('gpurun_out/r2_bench_arctas_$N.json').read().strip().splitlines()[-1])
print('AR_CTAS=$C N=$N value', round(l['value']), 'ms', round(l['ms_per_step'],3), 'e2e ms', round(l['e2e']['ms_per_step'],3), 'dp', l.get('dp_equals_single'))" || tail -5 gpurun_out/r2_bench_arctas_$N.err
done
