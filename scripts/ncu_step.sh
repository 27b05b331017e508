#!/bin/bash
# ncu captures for profiles/: launch list of one step + full sets of the dominant kernels (run under gpurun, 1 GPU)
set -x
mkdir -p gpurun_out
# launch list: skip the 3 warm-up eager steps? profile_step runs 3 steps then 1 instrumented; list everything of the last step
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv \
    python scripts/profile_step.py cifar 256 3 > gpurun_out/launches_r1.stdout 2>&1
# full capture: forward conv 128->128 k3 on 512 images (launch 23 of conv_gemm in step 1), stride-2 dgrad 64<-128 (launch 36)
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 23 -c 1 -o gpurun_out/prof_r1_conv_fwd -f \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 35 -c 2 -o gpurun_out/prof_r1_conv_dgrad -f \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm_kernel -s 12 -c 2 -o gpurun_out/prof_r1_wgrad -f \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mmd_fused -s 1 -c 1 -o gpurun_out/prof_r1_mmd -f \
    python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
ls -la gpurun_out/
