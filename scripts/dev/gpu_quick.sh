#!/bin/bash
# quick perf + correctness loop: bench (no CPU baseline), kernel / step / golden tests
TAG=${1:-q}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.txt 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_${TAG}.txt').read().strip().splitlines()[-1]);print('ms/step',d['ms_per_step'],'img/s',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'gemm_ms',d['roofline']['gemm_ms_per_step'])" || tail -5 gpurun_out/bench_${TAG}.txt
for f in test_gpu_kernels test_gpu_step test_gpu_golden_api; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout 600 > gpurun_out/$f.txt 2>&1; echo "$f rc=$?"; tail -3 gpurun_out/$f.txt
done
python scripts/profile_step.py cifar 256 3 > gpurun_out/events_${TAG}.txt 2>&1; awk '$6>1 && $1=="fwd" {f+=$10} $6>1 && $1=="dgrad" {d+=$10} $6>1 && $1=="wgrad" {w+=$10} $6==1 {s+=$10} END {print "batch1",s,"fwd",f,"dgrad",d,"wgrad",w}' gpurun_out/events_${TAG}.txt
