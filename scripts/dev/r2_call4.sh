#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r2d_kernels.txt
tail -5 gpurun_out/r2d_kernels.txt
if grep -q "failed\|error\|Error" gpurun_out/r2d_kernels.txt; then exit 0; fi
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2d_pytest.txt
tail -5 gpurun_out/r2d_pytest.txt
timeout 200 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2d_events.txt 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
MMDGAN_PROF=1 timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2f_prof.txt 2>&1
tail -3 gpurun_out/r2d_events.txt; tail -c 300 gpurun_out/r2d_bench.json
