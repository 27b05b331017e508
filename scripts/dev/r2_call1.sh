#!/bin/bash
# round 2, first GPU call (1 GPU): full GPU test suite incl. the new baseline-operating-point parity tests, the shifted-descriptor probe,
# one bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2a_pytest.txt
bash scripts/dev/gpu_probe_desc_shift.sh > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest.txt; tail -30 gpurun_out/probe_desc_shift.txt; tail -c 1500 gpurun_out/r2a_bench.json
