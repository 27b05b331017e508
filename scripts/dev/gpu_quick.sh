#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}.txt 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/bench_${TAG}.txt').read().strip().splitlines()[-1]);print('ms/step',d['ms_per_step'],'img/s',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'gemm_ms',d['roofline']['gemm_ms_per_step'])" || tail -5 gpurun_out/bench_${TAG}.txt
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -3
