#!/bin/bash
# ncu captures for profiles/ (run under gpurun, 1 GPU): launch list of the profiled steps + full sets of the dominant kernels
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python scripts/profile_step.py cifar 256 3 > gpurun_out/events_${TAG}.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_pair_kernel -s 2 -c 2 -o gpurun_out/prof_${TAG}_conv_fwd -f \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 30 -c 2 -o gpurun_out/prof_${TAG}_conv_dgrad -f \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm_kernel -s 12 -c 2 -o gpurun_out/prof_${TAG}_wgrad -f \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mmd_fused -s 1 -c 1 -o gpurun_out/prof_${TAG}_mmd -f \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
python scripts/profile_step.py cifar 256 3 > gpurun_out/events_${TAG}.txt 2>&1
ls -la gpurun_out/ | tail -12
