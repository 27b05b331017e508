#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r2g_sweep.txt
run() { echo "== $*" >> gpurun_out/r2g_sweep.txt; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-roofline 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(l['ms_per_step'], l['value'])" >> gpurun_out/r2g_sweep.txt 2>&1; }
run A=0
run MMDGAN_SKIP_CONVERT=1
run MMDGAN_SKIP_WRED=1
run MMDGAN_SKIP_CONVERT=1 MMDGAN_SKIP_WRED=1
cat gpurun_out/r2g_sweep.txt
