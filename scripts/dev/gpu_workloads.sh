#!/bin/bash
mkdir -p gpurun_out
python scripts/mmd_sweep.py > gpurun_out/mmd_sweep.txt 2>&1; cat gpurun_out/mmd_sweep.txt
for w in stl celeba; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.txt 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_$w.txt').read().strip().splitlines()[-1]);print('$w','ms/step',d['ms_per_step'],'img/s',d['value'],'frac',d['roofline']['frac'])" || tail -3 gpurun_out/bench_$w.txt; done
