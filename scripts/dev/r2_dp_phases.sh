#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 scripts/dp_phase_times.py 2>&1 | grep -E "world=|Error|error" | tee gpurun_out/r2_dp_phases_$N.txt
timeout 200 python scripts/phase_times.py 2>&1 | tail -2 | tee -a gpurun_out/r2_dp_phases_$N.txt
