#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2h_pytest.txt
tail -4 gpurun_out/r2h_pytest.txt
timeout 300 python scripts/mmd_sweep.py > gpurun_out/r2_mmd_sweep.txt 2>&1
cat gpurun_out/r2_mmd_sweep.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -c 1200 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
