#!/bin/bash
TAG=${1:-r1b}
mkdir -p gpurun_out
python scripts/profile_step.py cifar 256 3 > gpurun_out/events_${TAG}.txt 2>&1; grep -E "wgrad|total" gpurun_out/events_${TAG}.txt | awk '$10>0.03 || /total/'
MMDGAN_WGRAD_BN=256 python scripts/profile_step.py cifar 256 3 > gpurun_out/events_${TAG}_w256.txt 2>&1; grep -E "wgrad|total" gpurun_out/events_${TAG}_w256.txt | awk '$10>0.03 || /total/'
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.txt 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_${TAG}.txt').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['frac'])"
MMDGAN_WGRAD_BN=256 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_w256.txt 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_${TAG}_w256.txt').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['value'],d['roofline']['frac'])"
for f in test_gpu_kernels test_gpu_golden_api test_gpu_step; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.txt 2>&1; echo "$f rc=$?"; tail -3 gpurun_out/$f.txt
done
