#!/bin/bash
# one gpurun call: per-op error table, the GPU test files one by one (a trapped kernel must not hide the others), a short bench
mkdir -p gpurun_out
timeout 300 python scripts/dev/dbg_linear.py > gpurun_out/dbg_linear.txt 2>&1; echo "dbg_linear rc=$?"
tail -25 gpurun_out/dbg_linear.txt
for f in test_gpu_kernels test_gpu_golden_api test_gpu_step test_gpu_multi; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.txt 2>&1; echo "$f rc=$?"; tail -6 gpurun_out/$f.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.txt 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.txt
