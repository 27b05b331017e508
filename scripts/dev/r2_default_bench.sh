#!/bin/bash
# the driver's own invocation: bench.py with no flags (N = 1), then the reference arm
mkdir -p gpurun_out
timeout -s INT -k 20 400 python bench.py > gpurun_out/r2_bench_cifar.json 2> gpurun_out/r2_bench_cifar.err; echo "rc=$?"
python -c "
import json
l=json.loads(open('gpurun_out/r2_bench_cifar.json').read().strip().splitlines()[-1])
print('steps', l['steps'], 'ms/step', l['ms_per_step'], 'value', l['value'], 'e2e', l['e2e']['value'], 'frac', l['roofline']['frac'], 'cpu', l['cpu_baseline']['value'], l['clocks'])" || tail -20 gpurun_out/r2_bench_cifar.err
