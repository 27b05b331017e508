#!/bin/bash
mkdir -p gpurun_out
MMDGAN_PROF=1 timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2_prof_roles.txt 2>&1
grep PROF gpurun_out/r2_prof_roles.txt | tail -45
