#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
: > gpurun_out/r2_dp_ctas_$N.txt
for c in 0 4 8; do
echo "== MMDGAN_AR_CTAS=$c" | tee -a gpurun_out/r2_dp_ctas_$N.txt
MMDGAN_AR_CTAS=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2972$c scripts/dp_phase_times.py 2>&1 | grep -E "world=|Error|error" | tee -a gpurun_out/r2_dp_ctas_$N.txt
done
