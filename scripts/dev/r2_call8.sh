#!/bin/bash
mkdir -p gpurun_out
for v in "MMDGAN_WGRAD_CTAS=148" "MMDGAN_WGRAD_CTAS=296" "MMDGAN_WGRAD_CTAS=260" "MMDGAN_WGRAD_NO_PAIR=1" "MMDGAN_WGRAD_CTAS=185"; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dp-check --no-strong 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['gemm_ms_per_step'])"
done > gpurun_out/r2_wsweep.txt 2>&1
cat gpurun_out/r2_wsweep.txt
