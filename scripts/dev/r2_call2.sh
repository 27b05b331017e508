#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.txt
python scripts/profile_step.py cifar 256 3 > gpurun_out/r2b_events.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -8 gpurun_out/r2b_pytest.txt; grep -E "\*|total" gpurun_out/r2b_events.txt; tail -c 600 gpurun_out/r2b_bench.json
