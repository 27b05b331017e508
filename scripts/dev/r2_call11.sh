#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/prof_r2_* 
bash scripts/ncu_step.sh r2 > gpurun_out/ncu_step.log 2>&1
tail -20 gpurun_out/ncu_step.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dp-check --no-strong 2>/dev/null | tail -c 1500
