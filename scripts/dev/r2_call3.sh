#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv3x3" -s 4 -c 4 \
    -o gpurun_out/prof_r2c_direct -f python scripts/profile_step.py cifar 256 3 > /dev/null 2>&1
ls -la gpurun_out/prof_r2c_direct.ncu-rep
