#!/bin/bash
# quick perf + correctness loop: bench (no CPU baseline) and the step / golden tests
TAG=${1:-q}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.txt 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_${TAG}.txt').read().strip().splitlines()[-1]);print('ms/step',d['ms_per_step'],'img/s',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'launches',d['gpu_launches_per_step'])" || tail -5 gpurun_out/bench_${TAG}.txt
for f in test_gpu_step test_gpu_golden_api; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.txt 2>&1; echo "$f rc=$?"; tail -3 gpurun_out/$f.txt
done
