#!/bin/bash
mkdir -p gpurun_out
MMDGAN_PROF=1 timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2f_prof.txt 2>&1
MMDGAN_PROF=1 MMDGAN_DEBUG=2 timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2f_prof_nomma.txt 2>&1
grep PROF gpurun_out/r2f_prof.txt | tail -28 | cut -c1-300
echo ======
grep PROF gpurun_out/r2f_prof_nomma.txt | tail -28 | cut -c1-300
