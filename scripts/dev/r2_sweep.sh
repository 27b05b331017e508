#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2g_sweep.txt
run() { echo "== $*" >> gpurun_out/r2g_sweep.txt; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['value'])" >> gpurun_out/r2g_sweep.txt 2>&1; }
run A=0
run MMDGAN_DIRECT_CONV=0
echo "== stl" >> gpurun_out/r2g_sweep.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload stl 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['value'], l['roofline']['frac'])" >> gpurun_out/r2g_sweep.txt 2>&1
echo "== celeba" >> gpurun_out/r2g_sweep.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload celeba 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['value'], l['roofline']['frac'])" >> gpurun_out/r2g_sweep.txt 2>&1
MMDGAN_DIRECT_CONV=0 timeout 200 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2g_events_nodirect.txt 2>&1
cat gpurun_out/r2g_sweep.txt; grep -E " 3 | 3  " gpurun_out/r2g_events_nodirect.txt | head
