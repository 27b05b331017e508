#!/bin/bash
mkdir -p gpurun_out
MMDGAN_PROF=1 timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2_prof_roles2.txt 2>&1
grep PROF gpurun_out/r2_prof_roles2.txt | tail -27 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dp-check --no-strong 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['gemm_ms_per_step'])"
done
