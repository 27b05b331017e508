#!/bin/bash
# steady-state gap between the device-resident and the end-to-end step (100 and 300 timed steps)
mkdir -p gpurun_out
for k in 100 300; do
timeout -s INT -k 20 120 python bench.py --steps $k --warmup 5 --no-cpu-baseline --no-dp-check --no-strong 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps', l['steps'], 'dev', l['ms_per_step'], 'e2e', l['e2e']['ms_per_step'], l['clocks'])"
done
