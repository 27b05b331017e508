#!/bin/bash
# Short bench lines back to back (SIGINT after 90 s so that a stuck run leaves a Python traceback) + the per-launch event table.
mkdir -p gpurun_out
for i in 1 2 3; do
timeout -s INT -k 20 90 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dp-check --no-strong > gpurun_out/r2_quick_bench.json 2> gpurun_out/r2_quick_bench.err
echo "bench rc=$?"
python -c "
import json
l=json.loads(open('gpurun_out/r2_quick_bench.json').read().strip().splitlines()[-1])
print('ms/step', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], 'gemm', l['roofline']['gemm_ms_per_step'], 'frac', l['roofline']['frac'], 'launches', l['gpu_launches_per_step'])" || tail -25 gpurun_out/r2_quick_bench.err
done
timeout 120 python scripts/profile_step.py cifar 256 3 2>&1 | grep -E "\*|gemm total"
