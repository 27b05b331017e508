#!/bin/bash
mkdir -p gpurun_out
MMDGAN_DIRECT_CONV=0 timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2_events_nodirect.txt 2>&1
grep -E "^kind| 3 | gemm total" gpurun_out/r2_events_nodirect.txt | head -20
timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2_events_direct.txt 2>&1
grep -E "^kind| 3 | gemm total" gpurun_out/r2_events_direct.txt | head -20
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dp-check --no-strong 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['gemm_ms_per_step'])"
