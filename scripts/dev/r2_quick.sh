#!/bin/bash
# Quick 1-GPU iteration loop: the -m gpu suite, then one bench line without the CPU baseline / N > 1 extras.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_quick_pytest.txt
tail -4 gpurun_out/r2_quick_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s INT -k 20 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dp-check --no-strong > gpurun_out/r2_quick_bench.json 2> gpurun_out/r2_quick_bench.err
python -c "
import json
l=json.loads(open('gpurun_out/r2_quick_bench.json').read().strip().splitlines()[-1])
print('ms/step', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], 'gemm', l['roofline']['gemm_ms_per_step'], 'frac', l['roofline']['frac'], 'launches', l['gpu_launches_per_step'])" || tail -3 gpurun_out/r2_quick_bench.err
