#!/bin/bash
mkdir -p gpurun_out
MMDGAN_PROF=1 MMDGAN_DIRECT_CONV=0 timeout 300 python scripts/profile_step.py cifar 256 3 > gpurun_out/r2f_prof_nd.txt 2>&1
grep PROF gpurun_out/r2f_prof_nd.txt | grep -E "bn=16|N=64 ksteps=2|N=8 " | tail -8 | cut -c1-330
