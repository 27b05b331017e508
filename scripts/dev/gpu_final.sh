#!/bin/bash
TAG=${1:-r1v3}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 900 > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/ncu_step.sh $TAG > /dev/null 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2>gpurun_out/bench_$TAG.err; tail -c 400 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_${TAG}_reference.json
du -sh gpurun_out
