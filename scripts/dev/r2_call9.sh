#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2h_pytest.txt
tail -4 gpurun_out/r2h_pytest.txt
for v in "MMDGAN_PDL=1" "MMDGAN_PDL=0" "MMDGAN_PDL=1" "MMDGAN_PDL=0"; do
  echo "== $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dp-check --no-strong 2>gpurun_out/r2i_bench.err | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['e2e']['ms_per_step'], l['roofline']['gemm_ms_per_step'])"
  tail -2 gpurun_out/r2i_bench.err
done > gpurun_out/r2_pdl.txt 2>&1
cat gpurun_out/r2_pdl.txt
