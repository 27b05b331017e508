#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/dev/dbg_linear.py > gpurun_out/dbg_linear.txt 2>&1; echo "dbg_linear rc=$?"
grep " 3 fwd" gpurun_out/dbg_linear.txt | cut -c1-110; tail -3 gpurun_out/dbg_linear.txt | cut -c1-300
for f in test_gpu_kernels test_gpu_golden_api test_gpu_step; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.txt 2>&1; echo "$f rc=$?"; tail -4 gpurun_out/$f.txt | cut -c1-300
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.txt 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench.txt').read().strip().splitlines()[-1]);print('ms/step',d['ms_per_step'],'img/s',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'gemm_ms',d['roofline']['gemm_ms_per_step'])" || tail -5 gpurun_out/bench.txt
